/*
 * mfs_oracle.c -- CPU restatement of the reference algorithm in plain C.  TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Only tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of bench.py may load this file's
 * shared object.  Nothing under mfs_b200/ links or calls it.
 *
 * It follows the reference's *dense* route step by step (not the structured route of the CUDA kernels):
 *   moment_quadrature        mfs/one_dim/quadtures.py:122-133   Hankel gather, Cholesky (lower), K = R^-1 H R^-T by two
 *                                                              triangular solves, symmetrise, dense symmetric
 *                                                              eigen-decomposition, w = V[0,:]^2, x = scale*lambda+mean
 *   moment_filter_rms        mfs/one_dim/filtering.py:73-86
 *   moment_filter_cms        mfs/one_dim/filtering.py:140-158
 *   moment_filter_scms       mfs/one_dim/filtering.py:218-237
 *   raw_moment_of_normal     mfs/one_dim/moments.py:70-74      (binomial sum, libm pow)
 *   TME transition moments   `tme.expectation` call sites mfs/one_dim/moments.py:151-171, evaluated per (node, order)
 *                            with libm pow like the reference's `u ** n`
 *   measurement pmfs/pdfs    mfs/one_dim/ss_models.py:43-47, :80-84; tests/test_filtering.py:41-42
 * The dense symmetric eigensolver is a Householder tridiagonalisation followed by implicit QL with accumulated
 * eigenvectors, i.e. the algorithm LAPACK's dsyev/dsyevd runs at these sizes (XLA:CPU's eigh lowers to syevd).
 *
 * Validated in tests/test_oracle_*.py against (i) the NumPy/SciPy restatement oracle/mfs_oracle.py (LAPACK) and
 * (ii) the golden vectors produced by running the reference's own source on the NumPy jax-shim.
 *
 * The argument struct is the product's mfs_filter1d_args (include/mfs_b200.h) with HOST pointers, so a test can hand
 * the same arguments to both implementations.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

#include "../include/mfs_b200.h"

#define MAXN MFS_MAX_N
#define MAXM (2 * MFS_MAX_N)

/* ---- dense symmetric eigensolver: a (n x n, row-major, symmetric) -> eigenvalues d, eigenvectors in columns of a ---- */
static void householder_tridiag(int n, double a[MAXN][MAXN], double d[MAXN], double e[MAXN]) {
  for (int i = n - 1; i > 0; --i) {
    int l = i - 1;
    double h = 0.0, scale = 0.0;
    if (l > 0) {
      for (int k = 0; k <= l; ++k) scale += fabs(a[i][k]);
      if (scale == 0.0) {
        e[i] = a[i][l];
      } else {
        for (int k = 0; k <= l; ++k) { a[i][k] /= scale; h += a[i][k] * a[i][k]; }
        double f = a[i][l];
        double g = (f >= 0.0) ? -sqrt(h) : sqrt(h);
        e[i] = scale * g;
        h -= f * g;
        a[i][l] = f - g;
        f = 0.0;
        for (int j = 0; j <= l; ++j) {
          a[j][i] = a[i][j] / h;
          g = 0.0;
          for (int k = 0; k <= j; ++k) g += a[j][k] * a[i][k];
          for (int k = j + 1; k <= l; ++k) g += a[k][j] * a[i][k];
          e[j] = g / h;
          f += e[j] * a[i][j];
        }
        const double hh = f / (h + h);
        for (int j = 0; j <= l; ++j) {
          f = a[i][j];
          e[j] = g = e[j] - hh * f;
          for (int k = 0; k <= j; ++k) a[j][k] -= (f * e[k] + g * a[i][k]);
        }
      }
    } else {
      e[i] = a[i][l];
    }
    d[i] = h;
  }
  d[0] = 0.0;
  e[0] = 0.0;
  for (int i = 0; i < n; ++i) {
    const int l = i - 1;
    if (d[i] != 0.0) {
      for (int j = 0; j <= l; ++j) {
        double g = 0.0;
        for (int k = 0; k <= l; ++k) g += a[i][k] * a[k][j];
        for (int k = 0; k <= l; ++k) a[k][j] -= g * a[k][i];
      }
    }
    d[i] = a[i][i];
    a[i][i] = 1.0;
    for (int j = 0; j <= l; ++j) a[j][i] = a[i][j] = 0.0;
  }
}

static int ql_implicit(int n, double d[MAXN], double e[MAXN], double z[MAXN][MAXN]) {
  for (int i = 1; i < n; ++i) e[i - 1] = e[i];
  e[n - 1] = 0.0;
  for (int l = 0; l < n; ++l) {
    int iter = 0, m;
    do {
      for (m = l; m < n - 1; ++m) {
        const double dd = fabs(d[m]) + fabs(d[m + 1]);
        if (fabs(e[m]) <= 2.220446049250313e-16 * dd) break;
      }
      if (m != l) {
        if (iter++ == 60) return -1;
        double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
        double r = hypot(g, 1.0);
        g = d[m] - d[l] + e[l] / (g + (g >= 0.0 ? fabs(r) : -fabs(r)));
        double s = 1.0, c = 1.0, p = 0.0;
        int i;
        for (i = m - 1; i >= l; --i) {
          double f = s * e[i];
          const double b = c * e[i];
          e[i + 1] = (r = hypot(f, g));
          if (r == 0.0) { d[i + 1] -= p; e[m] = 0.0; break; }
          s = f / r;
          c = g / r;
          g = d[i + 1] - p;
          r = (d[i] - g) * s + 2.0 * c * b;
          d[i + 1] = g + (p = s * r);
          g = c * r - b;
          for (int k = 0; k < n; ++k) {
            f = z[k][i + 1];
            z[k][i + 1] = s * z[k][i] + c * f;
            z[k][i] = c * z[k][i] - s * f;
          }
        }
        if (r == 0.0 && i >= l) continue;
        d[l] -= p;
        e[l] = g;
        e[m] = 0.0;
      }
    } while (m != l);
  }
  return 0;
}

/* mfs/utils.py:495-538 (ldl + ldl_chol) */
static void ldl_chol(int n, double G[MAXN][MAXN], double R[MAXN][MAXN]) {
  double l[MAXN][MAXN], d[MAXN], fro = 0.0;
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { l[i][j] = (i == j); fro += G[i][j] * G[i][j]; }
  const double eps = 1e-8 * sqrt(fro);
  for (int i = 1; i < n; ++i) l[i][0] = G[i][0] / G[0][0];
  for (int j = 0; j < n; ++j) d[j] = G[0][0];
  for (int j = 1; j < n; ++j) {
    double v[MAXN], dj = G[j][j];
    for (int k = 0; k < j; ++k) { v[k] = l[j][k] * d[k]; dj -= l[j][k] * v[k]; }
    d[j] = dj;
    for (int i = j + 1; i < n; ++i) {
      double s = G[i][j];
      for (int k = 0; k < j; ++k) s -= l[i][k] * v[k];
      l[i][j] = s / dj;
    }
  }
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) R[i][j] = l[i][j] * (d[j] < 0 ? eps : sqrt(d[j]));
}

/* mfs/one_dim/quadtures.py:122-133.  Returns 0 on success, -1 when the result is NaN (non-PD G, eig failure). */
static int moment_quadrature(int n, const double* ms, double mean, double scale, int use_ldl, double* w, double* x) {
  double G[MAXN][MAXN], H[MAXN][MAXN], R[MAXN][MAXN], Y[MAXN][MAXN], K[MAXN][MAXN];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) { G[i][j] = ms[i + j]; H[i][j] = ms[i + j + 1]; R[i][j] = 0.0; }
  if (use_ldl) {
    ldl_chol(n, G, R);
    for (int i = 0; i < n; ++i) for (int j = 0; j <= i; ++j) if (!isfinite(R[i][j])) return -1;
    for (int i = 0; i < n; ++i) if (R[i][i] == 0.0) return -1;
  } else {
    /* Cholesky--Crout, lower; LAPACK potrf fails on a pivot <= 0 or NaN and jax then returns NaN */
    for (int j = 0; j < n; ++j) {
      double s = G[j][j];
      for (int k = 0; k < j; ++k) s -= R[j][k] * R[j][k];
      if (!(s > 0.0) || !isfinite(s)) return -1;
      R[j][j] = sqrt(s);
      for (int i = j + 1; i < n; ++i) {
        double t = G[i][j];
        for (int k = 0; k < j; ++k) t -= R[i][k] * R[j][k];
        R[i][j] = t / R[j][j];
      }
    }
  }
  /* Y = R^-1 H (forward substitution, all columns), K = Y R^-T  <=>  K^T = R^-1 Y^T */
  for (int c = 0; c < n; ++c)
    for (int i = 0; i < n; ++i) {
      double s = H[i][c];
      for (int k = 0; k < i; ++k) s -= R[i][k] * Y[k][c];
      Y[i][c] = s / R[i][i];
    }
  for (int r = 0; r < n; ++r)
    for (int i = 0; i < n; ++i) {
      double s = Y[r][i];
      for (int k = 0; k < i; ++k) s -= R[i][k] * K[r][k];
      K[r][i] = s / R[i][i];
    }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < i; ++j) { const double v = 0.5 * (K[i][j] + K[j][i]); K[i][j] = K[j][i] = v; }
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) if (!isfinite(K[i][j])) return -1;
  double d[MAXN], e[MAXN];
  householder_tridiag(n, K, d, e);
  if (ql_implicit(n, d, e, K)) return -1;
  for (int i = 0; i < n; ++i) { w[i] = K[0][i] * K[0][i]; x[i] = scale * d[i] + mean; }
  return 0;
}

/* ---- model functors -------------------------------------------------------------------------------------------- */
typedef struct { double a0, a1, a2, a3, a4; } jet_t;

static jet_t drift_jet(int drift_id, double x, const double* prm) {
  jet_t j;
  if (drift_id == MFS_DRIFT_BENES) {
    const double t = tanh(x), u = 1.0 - t * t;
    j.a0 = t; j.a1 = u; j.a2 = -2.0 * t * u; j.a3 = -2.0 * u * u + 4.0 * t * t * u;
    j.a4 = 16.0 * t * u * u - 8.0 * t * t * t * u;
  } else if (drift_id == MFS_DRIFT_WELL) {
    const double th = prm[0];
    j.a0 = x * (1.0 - th * x * x); j.a1 = 1.0 - 3.0 * th * x * x; j.a2 = -6.0 * th * x; j.a3 = -6.0 * th; j.a4 = 0.0;
  } else {
    j.a0 = prm[0] * x; j.a1 = prm[0]; j.a2 = j.a3 = j.a4 = 0.0;
  }
  return j;
}

/* coefficients g^r_k of A^r phi = sum_k g^r_k phi^(k), r = 1..3, A phi = a phi' + c phi'' */
static void generator_powers(const jet_t* j, double c, double g[4][7]) {
  const double a = j->a0, a1 = j->a1, a2 = j->a2, a3 = j->a3, a4 = j->a4;
  memset(g, 0, sizeof(double) * 4 * 7);
  g[0][0] = 1.0;
  g[1][1] = a; g[1][2] = c;
  const double h1 = a * a1 + c * a2, h2 = a * a + 2 * c * a1, h3 = 2 * a * c, h4 = c * c;
  g[2][1] = h1; g[2][2] = h2; g[2][3] = h3; g[2][4] = h4;
  const double h1p = a1 * a1 + a * a2 + c * a3, h1pp = 3 * a1 * a2 + a * a3 + c * a4;
  const double h2p = 2 * a * a1 + 2 * c * a2, h2pp = 2 * a1 * a1 + 2 * a * a2 + 2 * c * a3;
  const double h3p = 2 * c * a1, h3pp = 2 * c * a2;
  g[3][1] = a * h1p + c * h1pp;
  g[3][2] = a * h1 + 2 * c * h1p + a * h2p + c * h2pp;
  g[3][3] = c * h1 + a * h2 + 2 * c * h2p + a * h3p + c * h3pp;
  g[3][4] = c * h2 + a * h3 + 2 * c * h3p;
  g[3][5] = c * h3 + a * h4;
  g[3][6] = c * h4;
}

/* E[((X_dt - m)/s)^p | x] by TME of the given order; pw[q] = ((x-m)/s)^q (0^0 = 1), sk[k] = s^-k, dtr[r] = dt^r/r! */
static double tme_moment(const double g[4][7], const double* dtr, int order, const double* pw, const double* sk, int p) {
  double out = 0.0;
  for (int r = 0; r <= order; ++r) {
    double ar = 0.0;
    for (int k = 0; k <= 2 * r && k <= p; ++k) {
      double ff = 1.0;
      for (int q = 0; q < k; ++q) ff *= (double)(p - q);
      ar += g[r][k] * ff * pw[p - k] * sk[k];
    }
    out += dtr[r] * ar;
  }
  return out;
}

static void power_table(double x, int n, double* pw) {
  pw[0] = 1.0;
  for (int q = 1; q < n; ++q) pw[q] = pw[q - 1] * x;
}

static void tme_mean_var(const jet_t* j, const double g[4][7], double dt, int order, double x, double c, double* mean,
                         double* var) {
  double mu = x, v = 0.0;
  { const double dtr[4] = {1.0, dt, dt * dt / 2.0, dt * dt * dt / 6.0};
    for (int r = 1; r <= order; ++r) mu += dtr[r] * g[r][1]; }
  /* Phi_r = A^r(x^2) - sum_s C(r,s) A^s x A^{r-s} x */
  const double A1x = g[1][1], A2x = g[2][1], A3x = g[3][1];
  const double phi1 = (2 * x * g[1][1] + 2 * g[1][2]) - 2 * x * A1x;
  const double phi2 = (2 * x * g[2][1] + 2 * g[2][2]) - (2 * x * A2x + 2 * A1x * A1x);
  const double phi3 = (2 * x * g[3][1] + 2 * g[3][2]) - (2 * x * A3x + 6 * A1x * A2x);
  if (order >= 1) v += dt * phi1;
  if (order >= 2) v += dt * dt / 2.0 * phi2;
  if (order >= 3) v += dt * dt * dt / 6.0 * phi3;
  (void)j; (void)c;
  *mean = mu; *var = v;
}

static double std_normal_raw_moment(int p) {
  if (p % 2) return 0.0;
  double f = 1.0;
  for (int k = 2; k <= p; ++k) f *= k;
  double h = 1.0;
  for (int k = 2; k <= p / 2; ++k) h *= k;
  return f / (pow(2.0, p / 2.0) * h);
}

static double binom(int n, int k) {
  double r = 1.0;
  for (int i = 1; i <= k; ++i) r = r * (n - k + i) / i;
  return r;
}

/* mfs/one_dim/moments.py:70-74: sum_m C(p,m) mean^m variance^((p-m)/2) E[Z^(p-m)].  mp[m] = mean^m, vh[k] = variance^(k/2)
 * (odd k through sqrt, so a negative variance gives NaN * 0 = NaN for every p >= 1, as in the reference). */
static double raw_moment_of_normal(const double* mp, const double* vh, int p) {
  double s = 0.0;
  for (int m = 0; m <= p; ++m) s += binom(p, m) * mp[m] * vh[p - m] * std_normal_raw_moment(p - m);
  return s;
}

static void normal_tables(double mean, double variance, int M, double* mp, double* vh) {
  const double sd = sqrt(variance);
  mp[0] = 1.0; vh[0] = 1.0;
  for (int k = 1; k < M; ++k) {
    mp[k] = mp[k - 1] * mean;
    vh[k] = (k % 2) ? vh[k - 1] * sd : vh[k - 2] * variance;
  }
}

static void cond_mean_var(const mfs_filter1d_args* a, const double* tprm, double x, double* mean, double* var) {
  const double c = 0.5 * a->dispersion * a->dispersion;
  if (a->trans_id == MFS_TRANS_NORMAL_AFFINE) { *mean = tprm[0] * x; *var = tprm[1]; return; }
  const jet_t j = drift_jet(a->drift_id, x, tprm);
  if (a->trans_id == MFS_TRANS_EULER) { *mean = x + j.a0 * a->dt; *var = a->dispersion * a->dispersion * a->dt; return; }
  double g[4][7];
  generator_powers(&j, c, g);
  tme_mean_var(&j, g, a->dt, a->tme_order, x, c, mean, var);
}

static double xlogy(double x, double y) { return x == 0.0 ? 0.0 : x * log(y); }

static double measurement_pdf(int meas_id, double y, double x, const double* prm) {
  if (meas_id == MFS_MEAS_BERNOULLI_LOGISTIC_CUBIC) {
    const double p = 1.0 / (1.0 + exp(-(x * x * x / prm[0] - prm[1])));
    /* jax.scipy.stats.bernoulli.pmf = exp(xlogy(y, p) + xlog1py(1 - y, -p)) */
    const double omy = 1.0 - y;
    return exp(xlogy(y, p) + (omy == 0.0 ? 0.0 : omy * log1p(-p)));
  } else if (meas_id == MFS_MEAS_POISSON_SOFTPLUS) {
    const double mu = log(1.0 + exp(prm[0] * x));
    return exp(xlogy(y, mu) - lgamma(y + 1.0) - mu);
  } else {
    const double z = (y - prm[0] * x) / prm[1];
    return exp(-0.5 * z * z - log(prm[1]) - 0.9189385332046727);
  }
}

static double load_y(const void* ys, int dtype, int64_t off) {
  if (dtype == MFS_YS_U8) return (double)((const unsigned char*)ys)[off];
  if (dtype == MFS_YS_I32) return (double)((const int32_t*)ys)[off];
  return ((const double*)ys)[off];
}

/* one filter, all T steps */
static void run_filter(const mfs_filter1d_args* a, int64_t b) {
  const int n = a->N, M = 2 * n, mode = a->mode;
  double ms[MAXM], w[MAXN], x[MAXN], lik[MAXN];
  memcpy(ms, a->ms0 + b * a->ms0_stride, sizeof(double) * M);
  double mean = mode != MFS_MODE_RAW ? a->mean0[b * a->mean0_stride] : 0.0;
  double scale = mode == MFS_MODE_SCALED ? a->scale0[b * a->scale0_stride] : 1.0;
  const double* tprm = a->trans_params + b * a->trans_param_stride;
  const double* mprm = a->meas_params + b * a->meas_param_stride;
  const double c = 0.5 * a->dispersion * a->dispersion;
  double nell = 0.0;
  int status = -1;
  int64_t t;
  for (t = 0; t < a->T; ++t) {
    const double y = load_y(a->ys, a->ys_dtype, b * a->ys_stride_b + t * a->ys_stride_t);
    /* prediction */
    if (moment_quadrature(n, ms, mean, scale, a->stable, w, x)) { status = (int)t; break; }
    if (mode != MFS_MODE_RAW) {
      double macc = 0.0, vacc = 0.0;
      for (int i = 0; i < n; ++i) {
        double mu, var;
        cond_mean_var(a, tprm, x[i], &mu, &var);
        if (a->trans_id == MFS_TRANS_TME || a->trans_id == MFS_TRANS_TME_NORMAL) {
          /* state_cond_mean = tme.expectation(identity): the same expansion */
        }
        macc += w[i] * mu; vacc += w[i] * var;
      }
      mean = macc;
      if (mode == MFS_MODE_SCALED) scale = sqrt(vacc);
    }
    for (int p = 0; p < M; ++p) ms[p] = 0.0;
    for (int i = 0; i < n; ++i) {
      if (a->trans_id == MFS_TRANS_TME) {
        const jet_t j = drift_jet(a->drift_id, x[i], tprm);
        double g[4][7], pw[MAXM], sk[7], dtr[4];
        generator_powers(&j, c, g);
        const double s_ = mode == MFS_MODE_SCALED ? scale : 1.0;
        power_table((x[i] - (mode == MFS_MODE_RAW ? 0.0 : mean)) / s_, M, pw);
        power_table(1.0 / s_, 7, sk);
        dtr[0] = 1.0; dtr[1] = a->dt; dtr[2] = a->dt * a->dt / 2.0; dtr[3] = a->dt * a->dt * a->dt / 6.0;
        for (int p = 0; p < M; ++p) ms[p] += w[i] * tme_moment(g, dtr, a->tme_order, pw, sk, p);
      } else {
        double mu, var, mp[MAXM], vh[MAXM];
        cond_mean_var(a, tprm, x[i], &mu, &var);
        normal_tables(mode == MFS_MODE_RAW ? mu : mu - mean, var, M, mp, vh);
        for (int p = 0; p < M; ++p) ms[p] += w[i] * raw_moment_of_normal(mp, vh, p);
      }
    }
    if (mode == MFS_MODE_SCALED && a->trans_id != MFS_TRANS_TME) {
      /* the reference's scaled Normal / Euler factories divide EVERY order by prod_k scale^k
       * (mfs/one_dim/moments.py:205, :243: `/ jnp.prod(scale ** jnp.arange(num_moments))`) -- restated as is */
      double s_all = 1.0;
      for (int k = 0; k < M; ++k) s_all *= pow(scale, (double)k);
      for (int p = 0; p < M; ++p) ms[p] /= s_all;
    }
    /* update */
    if (moment_quadrature(n, ms, mean, scale, a->stable, w, x)) { status = (int)t; break; }
    double pdf_y = 0.0;
    for (int i = 0; i < n; ++i) { lik[i] = measurement_pdf(a->meas_id, y, x[i], mprm); pdf_y += w[i] * lik[i]; }
    if (mode != MFS_MODE_RAW) {
      double acc = 0.0;
      for (int i = 0; i < n; ++i) acc += w[i] * x[i] * lik[i];
      mean = acc / pdf_y;
      if (mode == MFS_MODE_SCALED) {
        acc = 0.0;
        for (int i = 0; i < n; ++i) acc += w[i] * (x[i] - mean) * (x[i] - mean) * lik[i];
        scale = sqrt(acc / pdf_y);
      }
    }
    for (int p = 0; p < M; ++p) ms[p] = 0.0;
    for (int i = 0; i < n; ++i) {
      const double dlt = mode == MFS_MODE_RAW ? x[i] : mode == MFS_MODE_CENTRAL ? x[i] - mean : (x[i] - mean) / scale;
      double pw[MAXM];
      power_table(dlt, M, pw);
      for (int p = 0; p < M; ++p) ms[p] += w[i] * pw[p] * lik[i];
    }
    for (int p = 0; p < M; ++p) ms[p] /= pdf_y;
    nell -= log(pdf_y);
    if (a->out_mode == MFS_OUT_FULL) {
      memcpy(a->ms_out + b * a->ms_stride_b + t * a->ms_stride_t, ms, sizeof(double) * M);
      if (mode != MFS_MODE_RAW && a->mean_out) a->mean_out[b * a->aux_stride_b + t] = mean;
      if (mode == MFS_MODE_SCALED && a->scale_out) a->scale_out[b * a->aux_stride_b + t] = scale;
    }
  }
  if (status >= 0) {
    nell = NAN; mean = NAN; scale = NAN;
    for (int p = 0; p < M; ++p) ms[p] = NAN;
    if (a->out_mode == MFS_OUT_FULL)
      for (; t < a->T; ++t) {
        memcpy(a->ms_out + b * a->ms_stride_b + t * a->ms_stride_t, ms, sizeof(double) * M);
        if (mode != MFS_MODE_RAW && a->mean_out) a->mean_out[b * a->aux_stride_b + t] = NAN;
        if (mode == MFS_MODE_SCALED && a->scale_out) a->scale_out[b * a->aux_stride_b + t] = NAN;
      }
  }
  if (a->out_mode == MFS_OUT_LAST) {
    memcpy(a->ms_out + b * a->ms_stride_b, ms, sizeof(double) * M);
    if (mode != MFS_MODE_RAW && a->mean_out) a->mean_out[b * a->aux_stride_b] = mean;
    if (mode == MFS_MODE_SCALED && a->scale_out) a->scale_out[b * a->aux_stride_b] = scale;
  }
  a->nell_out[b] = nell;
  if (a->status_out) a->status_out[b] = status;
}

/* Entry point: same struct as the product, HOST pointers, filters handed out in blocks of 16 to `num_threads`
 * pthreads (0 = all online cores; no OpenMP runtime in this image).  Returns the number of threads used. */
typedef struct { const mfs_filter1d_args* a; atomic_llong next; } work_t;

static void* worker(void* arg) {
  work_t* wk = (work_t*)arg;
  for (;;) {
    const long long b0 = atomic_fetch_add(&wk->next, 16);
    if (b0 >= wk->a->B) break;
    const long long b1 = b0 + 16 < wk->a->B ? b0 + 16 : wk->a->B;
    for (long long b = b0; b < b1; ++b) run_filter(wk->a, b);
  }
  return NULL;
}

int mfs_oracle_max_threads(void) {
  const long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

int mfs_oracle_filter_1d(const mfs_filter1d_args* a, int num_threads) {
  if (!a || a->N < 1 || a->N > MAXN) return -1;
  if (num_threads <= 0) num_threads = mfs_oracle_max_threads();
  if (num_threads > 256) num_threads = 256;
  if ((long long)num_threads * 16 > a->B) num_threads = (int)((a->B + 15) / 16);
  if (num_threads < 1) num_threads = 1;
  work_t wk;
  wk.a = a;
  atomic_init(&wk.next, 0);
  if (num_threads == 1) { worker(&wk); return 1; }
  pthread_t th[256];
  int started = 0;
  for (int i = 0; i < num_threads - 1; ++i)
    if (pthread_create(&th[started], NULL, worker, &wk) == 0) ++started;
  worker(&wk);
  for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
  return started + 1;
}

int mfs_oracle_moment_quadrature_1d(int n, int64_t B, const double* ms, const double* mean, const double* scale,
                                    int use_ldl, double* weights, double* nodes) {
  if (n < 1 || n > MAXN) return -1;
  for (int64_t b = 0; b < B; ++b) {
    if (moment_quadrature(n, ms + b * 2 * n, mean ? mean[b] : 0.0, scale ? scale[b] : 1.0, use_ldl, weights + b * n,
                          nodes + b * n))
      for (int i = 0; i < n; ++i) weights[b * n + i] = nodes[b * n + i] = NAN;
  }
  return 0;
}
