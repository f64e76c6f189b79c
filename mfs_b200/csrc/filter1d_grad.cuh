// Negative log-likelihood AND its gradient w.r.t. functor parameters in one pass.
//
// The reference estimates parameters by differentiating nell through the whole filter scan with jax.grad and handing
// (value, gradient) to L-BFGS-B (dardel/parameter_estimation/mf.py:37-73: obj_func -> jaxopt.ScipyMinimize).  Here the
// same derivative is carried FORWARD through the filter: every quantity that depends on a parameter is a dual number
// (value + P tangents), so one kernel returns nell[b] and d nell[b] / d theta_p for B independent filters -- a theta
// grid, Monte-Carlo runs, or the B line-search points of a batched optimiser.  Forward mode is the natural choice on
// this path: there are 1-4 parameters, the scan is long (T = 1000) and nothing has to be stored for a reverse sweep.
//
// One filter per thread, time loop in the kernel.  Per step (mfs/one_dim/filtering.py:62-89 raw, :129-161 central):
//   prediction  ms^- = sum_i w_i T(x_i; theta)                       [dual]
//   quadrature  Hankel recurrence -> Jacobi matrix -> implicit QL     [dual; convergence tests on the values]
//   update      u_i = w_i p(y | x_i; theta), c = sum u_i, nell -= log c, posterior atoms (x_i, u_i / c)   [dual]
// The posterior of an update IS the N-atom measure {x_i, u_i/c}: its Gauss rule is itself, as a function of theta too
// (moments <-> atoms is a bijection), so the atoms are carried to the next prediction like in the value kernel; the
// Hankel pivots of the posterior moments are still checked on the values so that "not positive definite -> NaN" fires
// on the same quantity as in filter1d_kernel and in the reference's Cholesky.
// Differentiating the QL iteration: rotations are smooth functions of (d, e); deflation drops a coupling whose value
// -- and hence, by the cubic convergence of the shifted iteration, whose tangent -- is negligible.  Jacobi matrices of
// positive measures have simple eigenvalues, so the derivative of the eigen-decomposition exists.
#pragma once
#include "common.h"
#include "filter1d.cuh"

namespace mfs {

template <int P>
struct Dual {
  double v;
  double d[P];
};

template <int P> MFS_DEV Dual<P> make_dual(double v) {
  Dual<P> r;
  r.v = v;
#pragma unroll
  for (int k = 0; k < P; ++k) r.d[k] = 0.0;
  return r;
}
MFS_DEV double val(double a) { return a; }
template <int P> MFS_DEV double val(const Dual<P>& a) { return a.v; }

template <int P> MFS_DEV Dual<P> operator+(const Dual<P>& a, const Dual<P>& b) {
  Dual<P> r; r.v = a.v + b.v;
#pragma unroll
  for (int k = 0; k < P; ++k) r.d[k] = a.d[k] + b.d[k];
  return r;
}
template <int P> MFS_DEV Dual<P> operator-(const Dual<P>& a, const Dual<P>& b) {
  Dual<P> r; r.v = a.v - b.v;
#pragma unroll
  for (int k = 0; k < P; ++k) r.d[k] = a.d[k] - b.d[k];
  return r;
}
template <int P> MFS_DEV Dual<P> operator-(const Dual<P>& a) {
  Dual<P> r; r.v = -a.v;
#pragma unroll
  for (int k = 0; k < P; ++k) r.d[k] = -a.d[k];
  return r;
}
template <int P> MFS_DEV Dual<P> operator*(const Dual<P>& a, const Dual<P>& b) {
  Dual<P> r; r.v = a.v * b.v;
#pragma unroll
  for (int k = 0; k < P; ++k) r.d[k] = fma(a.v, b.d[k], a.d[k] * b.v);
  return r;
}
template <int P> MFS_DEV Dual<P> operator/(const Dual<P>& a, const Dual<P>& b) {
  Dual<P> r;
  const double inv = rcp_fast(b.v);     // <= 2 ulp, no denormal / inf branches (quadrature.cuh)
  r.v = a.v * inv;
#pragma unroll
  for (int k = 0; k < P; ++k) r.d[k] = fma(-r.v, b.d[k], a.d[k]) * inv;
  return r;
}
template <int P> MFS_DEV Dual<P> operator+(const Dual<P>& a, double b) { Dual<P> r = a; r.v += b; return r; }
template <int P> MFS_DEV Dual<P> operator+(double b, const Dual<P>& a) { Dual<P> r = a; r.v += b; return r; }
template <int P> MFS_DEV Dual<P> operator-(const Dual<P>& a, double b) { Dual<P> r = a; r.v -= b; return r; }
template <int P> MFS_DEV Dual<P> operator-(double b, const Dual<P>& a) { Dual<P> r = -a; r.v += b; return r; }
template <int P> MFS_DEV Dual<P> operator*(const Dual<P>& a, double b) {
  Dual<P> r; r.v = a.v * b;
#pragma unroll
  for (int k = 0; k < P; ++k) r.d[k] = a.d[k] * b;
  return r;
}
template <int P> MFS_DEV Dual<P> operator*(double b, const Dual<P>& a) { return a * b; }
template <int P> MFS_DEV Dual<P> operator/(const Dual<P>& a, double b) { return a * rcp_fast(b); }
template <int P> MFS_DEV Dual<P> operator/(double a, const Dual<P>& b) { return make_dual<P>(a) / b; }

// f(a) with derivative df: chain rule
template <int P> MFS_DEV Dual<P> chain(const Dual<P>& a, double f, double df) {
  Dual<P> r; r.v = f;
#pragma unroll
  for (int k = 0; k < P; ++k) r.d[k] = df * a.d[k];
  return r;
}
MFS_DEV double t_sqrt(double a) { return sqrt(a); }
// exp / log: the branch-free constant-bank versions of models.cuh (<= 4e-16 / 1e-15 relative against libm, measured on the
// device through mfs_math_selftest), libdevice outside their range
MFS_DEV double t_exp(double a) { return exp_any(a); }
MFS_DEV double t_log(double a) { return log_fast(a); }
MFS_DEV double t_tanh(double a) { return tanh(a); }
template <int P> MFS_DEV Dual<P> t_sqrt(const Dual<P>& a) { const double r = rsqrt_fast(a.v); return chain(a, a.v * r, 0.5 * r); }
template <int P> MFS_DEV Dual<P> t_exp(const Dual<P>& a) { const double e = exp_any(a.v); return chain(a, e, e); }
template <int P> MFS_DEV Dual<P> t_log(const Dual<P>& a) { return chain(a, log_fast(a.v), rcp_fast(a.v)); }
template <int P> MFS_DEV Dual<P> t_tanh(const Dual<P>& a) { const double t = tanh(a.v); return chain(a, t, fma(-t, t, 1.0)); }

template <class S> struct Scalar;
template <> struct Scalar<double> { static MFS_DEV double make(double v) { return v; } };
template <int P> struct Scalar<Dual<P>> { static MFS_DEV Dual<P> make(double v) { return make_dual<P>(v); } };

// dynamic-index read of a register array of S (select chain on compile-time indices)
template <class S, int N>
MFS_DEV S sel_t(const S (&a)[N], int idx) {
  S v = a[0];
#pragma unroll
  for (int i = 1; i < N; ++i)
    if (idx == i) v = a[i];
  return v;
}

// ---------------------------------------------------------------------------------------------------------------------
// Jacobi coefficients from moments by the Hankel-structured recurrence (same recurrence as jacobi_from_moments in
// quadrature.cuh; pivots d_k = L[k][k]^2 of the reference's Cholesky, mfs/one_dim/quadtures.py:122-127).
// ---------------------------------------------------------------------------------------------------------------------
template <class S, int N>
MFS_DEV bool jacobi_from_moments_t(const S (&m)[2 * N], S (&alpha)[N], S (&beta)[N]) {
  S ra[2 * N], rb[2 * N];
  bool ok = val(m[0]) > 0.0;
  S d_prev = m[0];
  S ratio_prev = m[1] / m[0];
  alpha[0] = ratio_prev;
#pragma unroll
  for (int l = 0; l < 2 * N; ++l) { ra[l] = m[l]; rb[l] = Scalar<S>::make(0.0); }
  S beta2_prev = Scalar<S>::make(0.0);
#pragma unroll
  for (int k = 1; k < N; ++k) {
    const S a = alpha[k - 1];
#pragma unroll
    for (int l = k; l < 2 * N - k; ++l) {
      if (k & 1) rb[l] = ra[l + 1] - a * ra[l] - beta2_prev * rb[l];
      else       ra[l] = rb[l + 1] - a * rb[l] - beta2_prev * ra[l];
    }
    const S dk = (k & 1) ? rb[k] : ra[k];
    const S sk1 = (k & 1) ? rb[k + 1] : ra[k + 1];
    ok = ok && (val(dk) > 0.0);
    const S ratio = sk1 / dk;
    alpha[k] = ratio - ratio_prev;
    beta2_prev = dk / d_prev;
    beta[k - 1] = t_sqrt(beta2_prev);
    d_prev = dk;
    ratio_prev = ratio;
  }
  beta[N - 1] = Scalar<S>::make(0.0);
  return ok;
}

// Implicit QL with Wilkinson shift on (d, e), carrying the first row z of the accumulated rotations
// (weights = z^2, mfs/one_dim/quadtures.py:133).  Deflation at the top index l, bottom of the block pinned at N-1;
// the chase is one unrolled run of rotations with compile-time indices.  Tests act on the values only.
template <class S, int N>
MFS_DEV bool tridiag_ql_t(S (&d)[N], S (&e)[N], S (&z)[N]) {
  z[0] = Scalar<S>::make(1.0);
#pragma unroll
  for (int i = 1; i < N; ++i) z[i] = Scalar<S>::make(0.0);
  if (N == 1) return true;
  int l = 0, iter = 0;
  while (true) {
    const S dl = sel_t<S, N>(d, l), dl1 = sel_t<S, N>(d, l + 1), el = sel_t<S, N>(e, l);
    if (!(fabs(val(el)) > kEps * (fabs(val(dl)) + fabs(val(dl1))))) {
      if (!(fabs(val(el)) >= 0.0)) return false;   // NaN
      ++l;
      iter = 0;
      if (l >= N - 1) return true;
      continue;
    }
    if (++iter > 40) return false;
    const S a = 0.5 * (dl1 - dl);
    const S e2 = el * el;
    const S hyp = t_sqrt(a * a + e2);
    S g = (d[N - 1] - dl) + e2 / (a + (val(a) >= 0.0 ? hyp : -hyp));
    S s = Scalar<S>::make(1.0), c = Scalar<S>::make(1.0), p = Scalar<S>::make(0.0);
    bool underflow = false;
#pragma unroll
    for (int i = N - 2; i >= 0; --i) {
      if (i >= l) {
        const S f = s * e[i];
        const S b = c * e[i];
        const S h = f * f + g * g;
        if (!(val(h) > 0.0)) underflow = true;
        const S r = t_sqrt(h);
        if (i + 1 < N - 1) e[i + 1] = r;
        s = f / r;
        c = g / r;
        g = d[i + 1] - p;
        const S rr = (d[i] - g) * s + 2.0 * (c * b);
        p = s * rr;
        d[i + 1] = g + p;
        g = c * rr - b;
        const S zi1 = z[i + 1];
        z[i + 1] = s * z[i] + c * zi1;
        z[i] = c * z[i] - s * zi1;
      } else if (i + 1 == l) {
        d[i + 1] = d[i + 1] - p;
        e[i + 1] = g;
      }
    }
    if (l == 0) {
      d[0] = d[0] - p;
      e[0] = g;
    }
    if (underflow) return false;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Model functors on S (same closed forms as models.cuh; precise libm calls instead of the MUFU-seeded fast paths)
// ---------------------------------------------------------------------------------------------------------------------
template <class S> struct JetT { S a0, a1, a2, a3, a4; };

template <class S>
MFS_DEV JetT<S> drift_jet_t(int drift_id, const S& x, const S* prm) {
  JetT<S> j;
  if (drift_id == MFS_DRIFT_BENES) {
    const S t = t_tanh(x);
    const S u = 1.0 - t * t;
    j.a0 = t;
    j.a1 = u;
    j.a2 = -2.0 * (t * u);
    j.a3 = u * (4.0 * (t * t) - 2.0 * u);
    j.a4 = 8.0 * (t * u) * (2.0 * u - t * t);
  } else if (drift_id == MFS_DRIFT_WELL) {
    const S th = prm[0];
    const S x2 = x * x;
    j.a0 = x * (1.0 - th * x2);
    j.a1 = 1.0 - 3.0 * (th * x2);
    j.a2 = -6.0 * (th * x);
    j.a3 = -6.0 * th;
    j.a4 = Scalar<S>::make(0.0);
  } else {
    j.a0 = prm[0] * x;
    j.a1 = prm[0];
    j.a2 = j.a3 = j.a4 = Scalar<S>::make(0.0);
  }
  return j;
}

// g_k = sum_r dt^r/r! g^r_k (see models.cuh: tme_coefficients); also yields tme.mean_and_cov:
//   mean = x + g_1,  var = 2 g_2 - 2 (sum_r ...) -- computed separately below to keep the reference's expansion order.
template <class S>
MFS_DEV void tme_coefficients_t(const JetT<S>& j, double c, double dt, int order, S (&g)[7]) {
  const S a = j.a0;
  g[0] = Scalar<S>::make(1.0);
  g[1] = dt * a;
  g[2] = Scalar<S>::make(dt * c);
  g[3] = g[4] = g[5] = g[6] = Scalar<S>::make(0.0);
  if (order >= 2) {
    const double w2 = 0.5 * dt * dt;
    const S h1 = a * j.a1 + c * j.a2;
    const S h2 = a * a + 2.0 * c * j.a1;
    const S h3 = 2.0 * c * a;
    const double h4 = c * c;
    g[1] = g[1] + w2 * h1;
    g[2] = g[2] + w2 * h2;
    g[3] = w2 * h3;
    g[4] = Scalar<S>::make(w2 * h4);
    if (order >= 3) {
      const double w3 = dt * dt * dt / 6.0;
      const S h1p = j.a1 * j.a1 + a * j.a2 + c * j.a3;
      const S h1pp = 3.0 * (j.a1 * j.a2) + a * j.a3 + c * j.a4;
      const S h2p = 2.0 * (a * j.a1 + c * j.a2);
      const S h2pp = 2.0 * (j.a1 * j.a1 + a * j.a2 + c * j.a3);
      const S h3p = 2.0 * c * j.a1;
      const S h3pp = 2.0 * c * j.a2;
      const S k1 = a * h1p + c * h1pp;
      const S k2 = a * h1 + 2.0 * c * h1p + a * h2p + c * h2pp;
      const S k3 = c * h1 + a * h2 + 2.0 * c * h2p + a * h3p + c * h3pp;
      const S k4 = c * h2 + a * h3 + 2.0 * c * h3p;
      const S k5 = c * h3 + h4 * a;
      const double k6 = c * h4;
      g[1] = g[1] + w3 * k1;
      g[2] = g[2] + w3 * k2;
      g[3] = g[3] + w3 * k3;
      g[4] = g[4] + w3 * k4;
      g[5] = w3 * k5;
      g[6] = Scalar<S>::make(w3 * k6);
    }
  }
}

template <class S>
MFS_DEV void tme_mean_var_t(const JetT<S>& j, const S& x, double c, double dt, int order, S& mean, S& var) {
  const S a = j.a0;
  mean = x + dt * a;
  var = Scalar<S>::make(2.0 * c * dt);
  if (order >= 2) {
    const double w2 = 0.5 * dt * dt;
    const S h1 = a * j.a1 + c * j.a2;
    mean = mean + w2 * h1;
    var = var + (w2 * 4.0 * c) * j.a1;
    if (order >= 3) {
      const double w3 = dt * dt * dt / 6.0;
      const S h1p = j.a1 * j.a1 + a * j.a2 + c * j.a3;
      const S h1pp = 3.0 * (j.a1 * j.a2) + a * j.a3 + c * j.a4;
      const S h2p = 2.0 * (a * j.a1 + c * j.a2);
      const S h2pp = 2.0 * (j.a1 * j.a1 + a * j.a2 + c * j.a3);
      const S k1 = a * h1p + c * h1pp;
      const S k2 = a * h1 + 2.0 * c * h1p + a * h2p + c * h2pp;
      mean = mean + w3 * k1;
      var = var + w3 * (2.0 * k2 - 6.0 * (a * h1));
    }
  }
}

template <class S>
MFS_DEV void normal_mean_var_t(int trans_id, int drift_id, int order, const S& x, double c, double dt, const S* prm,
                               S& mean, S& var) {
  if (trans_id == MFS_TRANS_NORMAL_AFFINE) {
    mean = prm[0] * x;
    var = prm[1];
  } else {
    const JetT<S> j = drift_jet_t<S>(drift_id, x, prm);
    if (trans_id == MFS_TRANS_EULER) {
      mean = x + dt * j.a0;
      var = Scalar<S>::make(2.0 * c * dt);
    } else {
      tme_mean_var_t<S>(j, x, c, dt, order, mean, var);
    }
  }
}

// p(y | x; theta): mfs/one_dim/ss_models.py:43-47, 80-84; tests/test_filtering.py:41-42.  lgam = lgamma(y + 1).
template <class S>
MFS_DEV S measurement_pdf_t(int meas_id, double y, double lgam, const S& x, const S* prm) {
  if (meas_id == MFS_MEAS_BERNOULLI_LOGISTIC_CUBIC) {
    const S zz = (x * x * x) / prm[0] - prm[1];
    const S p = 1.0 / (1.0 + t_exp(-zz));
    return (y != 0.0) ? p : 1.0 - p;
  }
  if (meas_id == MFS_MEAS_POISSON_SOFTPLUS) {
    const S mu = t_log(1.0 + t_exp(prm[0] * x));
    if (y == 0.0) return t_exp(-mu);
    return t_exp(y * t_log(mu) - lgam - mu);
  }
  const S zz = (y - prm[0] * x) / prm[1];
  return t_exp(-0.5 * (zz * zz)) / (prm[1] * 2.5066282746310002);
}

// ---------------------------------------------------------------------------------------------------------------------
// The kernel
// ---------------------------------------------------------------------------------------------------------------------
#ifndef MFS_GRAD_MIN_BLOCKS
#define MFS_GRAD_MIN_BLOCKS 2
#endif
// threads per CTA, and one CTA barrier per time step (the value kernel's step_barrier, filter1d.cuh): with the barrier the
// CTA's warps fetch the same code at the same time.  Measured (profiles/r2_ab_1d_step_barrier.log, N = 7): 64 threads x 4
// CTAs without the barrier 7.58e8 steps/s, with it 8.26e8, 128 threads x 2 CTAs with it 8.37e8 (+10 %).
#ifndef MFS_GRAD_BLOCK
#define MFS_GRAD_BLOCK 128
#endif
#ifndef MFS_GRAD_STEP_BARRIER
#define MFS_GRAD_STEP_BARRIER 1
#endif
#ifdef MFS_GRAD_UNROLL_NODES
#define MFS_NODE_LOOP _Pragma("unroll")
#else
#define MFS_NODE_LOOP _Pragma("unroll 1")
#endif

// Quadrature of a dual moment vector with the eigen-solve differentiated IMPLICITLY: the values go through the value
// kernel's register QL (tridiag_ql_first_row), the tangents follow from first-order perturbation theory of the
// symmetric tridiagonal eigenproblem J v_i = lambda_i v_i:
//     d lambda_i = v_i^T dJ v_i,      d z_i = sum_{j != i} (v_j^T dJ v_i) / (lambda_i - lambda_j) z_j,   z_i = v_i[0],
// with the eigenvectors rebuilt from (lambda_i, z_i) by the three-term recurrence of the orthonormal polynomials
// (v_i[k] = z_i p_k(lambda_i) / p_0).  Jacobi matrices of positive measures have simple eigenvalues (beta_k > 0), so
// the gaps never vanish.  Cost per tangent: N^3 + O(N^2) FMAs instead of pushing duals through ~2.3 N QL sweeps
// (a dual rotation is ~105 FP64 instructions against 22).  Falls back to the dual QL when the fast QL does not converge.
template <int N, int P>
MFS_DEV bool quadrature_implicit(const Dual<P> (&ms)[2 * N], const Dual<P>& mean, Dual<P> (&w)[N], Dual<P> (&x)[N]) {
  using S = Dual<P>;
  S a[N], b[N];
  if (!jacobi_from_moments_t<S, N>(ms, a, b)) return false;
  double d[N], e[N], z[N];
#pragma unroll
  for (int k = 0; k < N; ++k) { d[k] = a[k].v; e[k] = b[k].v; }
  if (!tridiag_ql_first_row<N>(d, e, z)) {
    S dd[N], ee[N], zz[N];
#pragma unroll
    for (int k = 0; k < N; ++k) { dd[k] = a[k]; ee[k] = b[k]; }
    const bool ok = tridiag_ql_t<S, N>(dd, ee, zz);
#pragma unroll
    for (int i = 0; i < N; ++i) { w[i] = zz[i] * zz[i]; x[i] = dd[i] + mean; }
    return ok;
  }
  // Eigenvectors are never stored: v_j[k] = z_j p_k(lambda_j) / p_0 is regenerated by the three-term recurrence on the
  // VALUES of (alpha, beta) while the dot products v_j . (dJ v_i) accumulate -- all N vectors v_j advance together
  // (N independent chains, compile-time j and k: registers), only the outer index i is a run-time loop.
  double lam_l[N], z_l[N];     // run-time indexed copies (local memory)
  double rb[N];
#pragma unroll
  for (int k = 0; k + 1 < N; ++k) rb[k] = rcp_fast(b[k].v);
#pragma unroll
  for (int i = 0; i < N; ++i) { lam_l[i] = d[i]; z_l[i] = z[i]; }
  MFS_NODE_LOOP
  for (int i = 0; i < N; ++i) {
    const double li = lam_l[i], zi = z_l[i];
    // v_i and u_p = dJ_p v_i
    double vi[N];
    vi[0] = zi;
#pragma unroll
    for (int k = 0; k + 1 < N; ++k) {
      double t = (li - a[k].v) * vi[k];
      if (k > 0) t = fma(-b[k > 0 ? k - 1 : 0].v, vi[k > 0 ? k - 1 : 0], t);
      vi[k + 1] = t * rb[k];
    }
    double u[P][N];
#pragma unroll
    for (int p = 0; p < P; ++p) {
#pragma unroll
      for (int k = 0; k < N; ++k) {
        double t = a[k].d[p] * vi[k];
        if (k > 0) t = fma(b[k > 0 ? k - 1 : 0].d[p], vi[k > 0 ? k - 1 : 0], t);
        if (k + 1 < N) t = fma(b[k].d[p], vi[k + 1 < N ? k + 1 : k], t);
        u[p][k] = t;
      }
    }
    double vm[N], vk[N], c[P][N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
      vm[j] = 0.0;
      vk[j] = z[j];
#pragma unroll
      for (int p = 0; p < P; ++p) c[p][j] = 0.0;
    }
#pragma unroll
    for (int k = 0; k < N; ++k) {
#pragma unroll
      for (int j = 0; j < N; ++j) {
#pragma unroll
        for (int p = 0; p < P; ++p) c[p][j] = fma(vk[j], u[p][k], c[p][j]);
      }
      if (k + 1 < N) {
#pragma unroll
        for (int j = 0; j < N; ++j) {
          double t = (d[j] - a[k].v) * vk[j];
          if (k > 0) t = fma(-b[k > 0 ? k - 1 : 0].v, vm[j], t);
          vm[j] = vk[j];
          vk[j] = t * rb[k];
        }
      }
    }
    double dl[P], dz[P];
#pragma unroll
    for (int p = 0; p < P; ++p) { dl[p] = 0.0; dz[p] = 0.0; }
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const bool self = (j == i);
      const double gz = self ? 0.0 : z[j] * rcp_fast(li - d[j]);
#pragma unroll
      for (int p = 0; p < P; ++p) {
        dl[p] = self ? c[p][j] : dl[p];
        dz[p] = fma(c[p][j], gz, dz[p]);
      }
    }
    S wi, xi;
    wi.v = zi * zi;
    xi.v = li + mean.v;
#pragma unroll
    for (int p = 0; p < P; ++p) {
      wi.d[p] = 2.0 * zi * dz[p];
      xi.d[p] = dl[p] + mean.d[p];
    }
    w[i] = wi;
    x[i] = xi;
  }
  return true;
}

template <class S, int N>
MFS_DEV bool quadrature_t(const S (&ms)[2 * N], const S& mean, S (&w)[N], S (&x)[N]) {
  S d[N], e[N], z[N];   // registers (compile-time indices); w / x are the caller's atom arrays (local memory)
  bool ok = jacobi_from_moments_t<S, N>(ms, d, e);
  if (!ok) return false;
  ok = tridiag_ql_t<S, N>(d, e, z);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    w[i] = z[i] * z[i];
    x[i] = d[i] + mean;
  }
  return ok;
}

// Default: duals through the QL iteration.  -DMFS_GRAD_IMPLICIT_EIG selects quadrature_implicit (same results, all
// gradient tests green), measured SLOWER at N = 7 (146 vs 103 ms, profiles/r1_grad_probe_v4_implicit.log and
// r1_ncu_filter1d_grad_N7_implicit.md; analysis in DESIGN.md section 5b).
#ifdef MFS_GRAD_IMPLICIT_EIG
#define MFS_GRAD_QUADRATURE quadrature_implicit<N, P>
#else
#define MFS_GRAD_QUADRATURE quadrature_t<S, N>
#endif

template <int N, int P>
__global__ void __launch_bounds__(MFS_GRAD_BLOCK, MFS_GRAD_MIN_BLOCKS) filter1d_grad_kernel(const mfs_filter1d_args A, const GradInfo G) {
  using S = Dual<P>;
  constexpr bool kBar = MFS_GRAD_STEP_BARRIER != 0;
  const int64_t b_own = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool present = true;
  if constexpr (kBar) {
    if ((int64_t)blockIdx.x * blockDim.x >= A.B) return;
    present = b_own < A.B;        // a thread beyond the batch runs the loop's barriers on the last filter's inputs, writes nothing
  } else {
    if (b_own >= A.B) return;
  }
  const int64_t b = present ? b_own : A.B - 1;
  const bool central = A.mode == MFS_MODE_CENTRAL;
  const bool normal_family = A.trans_id != MFS_TRANS_TME;
  const double c = 0.5 * A.dispersion * A.dispersion;
  const double dt = A.dt;

  S tprm[MFS_MAX_PARAMS], mprm[MFS_MAX_PARAMS];
#pragma unroll
  for (int i = 0; i < MFS_MAX_PARAMS; ++i) {
    tprm[i] = make_dual<P>(A.trans_params ? A.trans_params[b * A.trans_param_stride + i] : 0.0);
    mprm[i] = make_dual<P>(A.meas_params ? A.meas_params[b * A.meas_param_stride + i] : 0.0);
#pragma unroll
    for (int k = 0; k < P; ++k) {
      if (G.tangent_ids[k] == i) tprm[i].d[k] = 1.0;
      if (G.tangent_ids[k] == i + MFS_MAX_PARAMS) mprm[i].d[k] = 1.0;
    }
  }

  S ms[2 * N], w[N], x[N];
  S mean = make_dual<P>(central ? A.mean0[b * A.mean0_stride] : 0.0);
  S nell = make_dual<P>(0.0);
#pragma unroll
  for (int p = 0; p < 2 * N; ++p) ms[p] = make_dual<P>(A.ms0[b * A.ms0_stride + p]);
  int32_t status = -1;
  bool ok = MFS_GRAD_QUADRATURE(ms, mean, w, x);   // filtering.py:78 / :145 at k = 0
  if (!ok) status = 0;
  if (!present) ok = false;

  for (int64_t t = 0; t < A.T && (kBar || ok); ++t) {
    if constexpr (kBar) {
      __syncthreads();
      if (!ok) continue;
    }
    const double y = load_y(A.ys, A.ys_dtype, b * A.ys_stride_b + t * A.ys_stride_t);
    // ---- prediction (filtering.py:78-79 / :145-148)
    // The loops over the N atoms are ROLLED (MFS_NODE_LOOP): w, x, mu, var, g are indexed dynamically and live in
    // local memory, the 2N moment accumulators stay in registers.  Unrolled, a step is ~24k straight-line instructions
    // (385 KB at N = 7) and the kernel stalls on instruction fetch (ncu v1: stall_no_instruction 3.1 per issue).
    S mu[N], var[N];
    S g[N][7];
    if (normal_family) {
      MFS_NODE_LOOP
      for (int i = 0; i < N; ++i) normal_mean_var_t<S>(A.trans_id, A.drift_id, A.tme_order, x[i], c, dt, tprm, mu[i], var[i]);
    } else {
      MFS_NODE_LOOP
      for (int i = 0; i < N; ++i) {
        const JetT<S> j = drift_jet_t<S>(A.drift_id, x[i], tprm);
        S gi[7];
        tme_coefficients_t<S>(j, c, dt, A.tme_order, gi);
#pragma unroll
        for (int k = 0; k < 7; ++k) g[i][k] = gi[k];
        mu[i] = x[i] + gi[1];      // tme.expectation(identity) = x + sum_r dt^r/r! A^r x
      }
    }
    if (central) {
      S acc = make_dual<P>(0.0);
      MFS_NODE_LOOP
      for (int i = 0; i < N; ++i) acc = acc + w[i] * mu[i];
      mean = acc;
    }
#pragma unroll
    for (int p = 0; p < 2 * N; ++p) ms[p] = make_dual<P>(0.0);
    if (normal_family) {
      // moments of N(mu - mean, var): M_p = mu M_{p-1} + (p - 1) var M_{p-2}   (= the binomial sum of moments.py:70-74)
      MFS_NODE_LOOP
      for (int i = 0; i < N; ++i) {
        const S wi = w[i], vi = var[i];
        S m = central ? mu[i] - mean : mu[i];
        if (val(vi) < 0.0) m.v = nan("");
        S m2 = make_dual<P>(1.0), m1 = m;
        ms[0] = ms[0] + wi;
        ms[1] = ms[1] + wi * m1;
#pragma unroll
        for (int p = 2; p < 2 * N; ++p) {
          const S mp = m * m1 + ((double)(p - 1) * vi) * m2;
          ms[p] = ms[p] + wi * mp;
          m2 = m1;
          m1 = mp;
        }
      }
    } else {
      // T_p(x) = p! sum_k g_k(x) delta^(p-k)/(p-k)!,  delta = x - mean (central) or x (raw)      moments.py:141-179
      MFS_NODE_LOOP
      for (int i = 0; i < N; ++i) {
        const S wi = w[i];
        const S delta = central ? x[i] - mean : x[i];
        S gi[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) gi[k] = g[i][k];
        S pq[2 * N];
        pq[0] = make_dual<P>(1.0);
#pragma unroll
        for (int q = 1; q < 2 * N; ++q) pq[q] = (pq[q - 1] * delta) * (1.0 / q);
#pragma unroll
        for (int p = 0; p < 2 * N; ++p) {
          S acc = make_dual<P>(0.0);
#pragma unroll
          for (int k = 0; k < 7; ++k)
            if (k <= p) acc = acc + gi[k] * pq[p - k];
          ms[p] = ms[p] + wi * acc;
        }
      }
      double f = 1.0;
#pragma unroll
      for (int p = 2; p < 2 * N; ++p) { f *= (double)p; ms[p] = ms[p] * f; }
    }
    // ---- update (filtering.py:81-86 / :150-158)
    ok = MFS_GRAD_QUADRATURE(ms, central ? mean : make_dual<P>(0.0), w, x);
    if (!ok) {
      status = (int32_t)t;
      if constexpr (kBar) continue; else break;
    }
    const double lgam = (A.meas_id == MFS_MEAS_POISSON_SOFTPLUS) ? log_factorial(y) : 0.0;
    S cc = make_dual<P>(0.0);
    MFS_NODE_LOOP
    for (int i = 0; i < N; ++i) {
      const S u = w[i] * measurement_pdf_t<S>(A.meas_id, y, lgam, x[i], mprm);
      w[i] = u;
      cc = cc + u;
    }
    {
      const S cinv = 1.0 / cc;
      MFS_NODE_LOOP
      for (int i = 0; i < N; ++i) w[i] = w[i] * cinv;
    }
    nell = nell - t_log(cc);
    // posterior moments (values only): the pivots the next prediction's Cholesky would see
    {
      double pm[2 * N], pa[N], pb[N];
      double pmean = 0.0;
      if (central) {
        MFS_NODE_LOOP
        for (int i = 0; i < N; ++i) pmean = fma(w[i].v, x[i].v, pmean);
      }
#pragma unroll
      for (int p = 0; p < 2 * N; ++p) pm[p] = 0.0;
      MFS_NODE_LOOP
      for (int i = 0; i < N; ++i) {
        const double delta = x[i].v - pmean;
        double pw = w[i].v;
#pragma unroll
        for (int p = 0; p < 2 * N; ++p) { pm[p] += pw; pw *= delta; }
      }
      // (checked by the NEXT step's prediction in the reference scan: the failure is recorded at t + 1; the last
      // step's posterior is never factorised)
      if (t + 1 < A.T) {
        ok = jacobi_from_moments_t<double, N>(pm, pa, pb);
        if (!ok) status = (int32_t)(t + 1);
      }
    }
  }
  if (!present) return;
  const double qnan = nan("");
  A.nell_out[b] = ok ? nell.v : qnan;
#pragma unroll
  for (int k = 0; k < P; ++k) G.grad_out[b * P + k] = ok ? nell.d[k] : qnan;
  if (A.status_out) A.status_out[b] = status;
}

}  // namespace mfs
