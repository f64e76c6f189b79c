// Device functors: drift jets, TME generator coefficients, Normal-moment recursion, measurement likelihoods.
//
// TME (third-party `tme` package used at mfs/one_dim/moments.py:151-175, restated from its definition):
//   E[phi(X_dt) | x] ~= sum_{r<=M} dt^r/r! (A^r phi)(x),   A phi = a phi' + c phi'',  c = b^2/2  (constant dispersion b).
// For any smooth phi,  A^r phi = sum_k g^r_k(x) phi^(k)(x)  with g^r_k polynomials in the drift jet (a, a', ..., a'''').
// With phi(u) = ((u-m)/s)^p:  phi^(k)(x) = p!/(p-k)! ((x-m)/s)^(p-k) s^-k, hence
//   T_p(x) = p! * sum_{k=0}^{2M} G_k(x) * pq[p-k],   G_k = s^-k sum_r dt^r/r! g^r_k,   pq[q] = ((x-m)/s)^q / q!.
#pragma once
#include "quadrature.cuh"
#include "../../include/mfs_b200.h"

namespace mfs {

struct Jet { double a0, a1, a2, a3, a4; };  // a, a', a'', a''', a''''

// exp(x) for |x| <= 708 without libdevice's special-case branches and with the polynomial coefficients as constant-bank
// operands of the DFMAs (libdevice's exp materialises every coefficient with two UMOVs: 22 extra issue slots per call,
// ncu r1 v5: 70 % of the instructions of the logistic were not FP64).  k = rint(x log2 e) by the magic-number add,
// r = x - k ln2 (two-term Cody--Waite), degree-11 polynomial (Chebyshev interpolant on |r| <= ln2/2: approximation
// error 4e-18, tools/fit_exp_poly.py), 2^k added into the exponent field.  <= 1 ulp-class error like libm.
static __constant__ double kExpPoly[12] = {
    0x1.0000000000000p+0, 0x1.0000000000000p+0, 0x1.0000000000011p-1, 0x1.555555555555ap-3, 0x1.555555554f0bap-5,
    0x1.111111110f21ep-7, 0x1.6c16c1880029fp-10, 0x1.a01a01b1461c5p-13, 0x1.a01991a10d9aep-16, 0x1.71ddf56d8deb5p-19,
    0x1.28b4101c77212p-22, 0x1.af632a0f7e2cep-26};

MFS_DEV double exp_fast(double x) {   // caller guarantees |x| <= 708 (result stays a normal number)
  const double t = fma(x, 1.4426950408889634, 6755399441055744.0);
  const int k = __double2loint(t);
  const double kf = t - 6755399441055744.0;
  double r = fma(kf, -0x1.62e42fefa39efp-1, x);
  r = fma(kf, -0x1.abc9e3b39803fp-56, r);
  double p = kExpPoly[11];
#pragma unroll
  for (int j = 10; j >= 0; --j) p = fma(p, r, kExpPoly[j]);
  return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// tanh through one exp and one reciprocal: 1 - 2/(1 + e^{2x}).  Absolute error ~1e-16 (relative accuracy degrades
// only for |x| << 1, where tanh enters the transition moments multiplied by dt and added to O(1) terms).  The
// exponent is clamped so that e^{2x} stays a normal number; tanh is +-1 to the last bit there.
MFS_DEV double tanh_fast(double x) {
  return fma(-2.0, rcp_fast(1.0 + exp_fast(fmax(fmin(2.0 * x, 708.0), -708.0))), 1.0);
}

// mfs/one_dim/ss_models.py:37 (tanh), :71 (x(1-theta1 x^2)); tests/test_filtering.py:45 (-x/ell as a*x)
MFS_DEV Jet drift_jet(int drift_id, double x, const double* prm) {
  Jet j;
  if (drift_id == MFS_DRIFT_BENES) {
    const double t = tanh_fast(x);
    const double u = fma(-t, t, 1.0);
    j.a0 = t;
    j.a1 = u;
    j.a2 = -2.0 * t * u;
    j.a3 = u * fma(4.0 * t, t, -2.0 * u);
    j.a4 = 8.0 * t * u * fma(-t, t, 2.0 * u);
  } else if (drift_id == MFS_DRIFT_WELL) {
    const double th = prm[0];
    const double x2 = x * x;
    j.a0 = x * fma(-th, x2, 1.0);
    j.a1 = fma(-3.0 * th, x2, 1.0);
    j.a2 = -6.0 * th * x;
    j.a3 = -6.0 * th;
    j.a4 = 0.0;
  } else {  // MFS_DRIFT_LINEAR
    j.a0 = prm[0] * x;
    j.a1 = prm[0];
    j.a2 = j.a3 = j.a4 = 0.0;
  }
  return j;
}

// g_k = sum_{r=1}^{order} dt^r/r! g^r_k, k = 1..6 (g_0 = 1 implied).  Also used for tme.mean_and_cov.
struct TmeCoef { double g[7]; };

MFS_DEV TmeCoef tme_coefficients(const Jet& j, double c, double dt, int order) {
  TmeCoef o;
  const double a = j.a0;
  // r = 1
  o.g[0] = 1.0;
  o.g[1] = dt * a;
  o.g[2] = dt * c;
  o.g[3] = o.g[4] = o.g[5] = o.g[6] = 0.0;
  if (order >= 2) {
    const double w2 = 0.5 * dt * dt;
    const double h1 = fma(a, j.a1, c * j.a2);
    const double h2 = fma(a, a, 2.0 * c * j.a1);
    const double h3 = 2.0 * a * c;
    const double h4 = c * c;
    o.g[1] = fma(w2, h1, o.g[1]);
    o.g[2] = fma(w2, h2, o.g[2]);
    o.g[3] = w2 * h3;
    o.g[4] = w2 * h4;
    if (order >= 3) {
      const double w3 = dt * dt * dt / 6.0;
      const double h1p = fma(j.a1, j.a1, fma(a, j.a2, c * j.a3));
      const double h1pp = fma(3.0 * j.a1, j.a2, fma(a, j.a3, c * j.a4));
      const double h2p = 2.0 * fma(a, j.a1, c * j.a2);
      const double h2pp = 2.0 * fma(j.a1, j.a1, fma(a, j.a2, c * j.a3));
      const double h3p = 2.0 * c * j.a1;
      const double h3pp = 2.0 * c * j.a2;
      const double k1 = fma(a, h1p, c * h1pp);
      const double k2 = fma(a, h1, fma(2.0 * c, h1p, fma(a, h2p, c * h2pp)));
      const double k3 = fma(c, h1, fma(a, h2, fma(2.0 * c, h2p, fma(a, h3p, c * h3pp))));
      const double k4 = fma(c, h2, fma(a, h3, 2.0 * c * h3p));
      const double k5 = fma(c, h3, a * h4);
      const double k6 = c * h4;
      o.g[1] = fma(w3, k1, o.g[1]);
      o.g[2] = fma(w3, k2, o.g[2]);
      o.g[3] = fma(w3, k3, o.g[3]);
      o.g[4] = fma(w3, k4, o.g[4]);
      o.g[5] = w3 * k5;
      o.g[6] = w3 * k6;
    }
  }
  return o;
}

// tme.mean_and_cov (call sites mfs/one_dim/moments.py:175,190): mean = x + sum dt^r/r! A^r id,
// var = sum_{r>=1} dt^r/r! [A^r(x^2) - sum_s C(r,s) A^s x A^{r-s} x] = dt 2c + dt^2/2 4c a' + dt^3/6 (2 g^3_2 - 6 a h1).
MFS_DEV void tme_mean_var(const Jet& j, double x, double c, double dt, int order, double& mean, double& var) {
  const double a = j.a0;
  mean = fma(dt, a, x);
  var = 2.0 * c * dt;
  if (order >= 2) {
    const double w2 = 0.5 * dt * dt;
    const double h1 = fma(a, j.a1, c * j.a2);
    mean = fma(w2, h1, mean);
    var = fma(w2, 4.0 * c * j.a1, var);
    if (order >= 3) {
      const double w3 = dt * dt * dt / 6.0;
      const double h1p = fma(j.a1, j.a1, fma(a, j.a2, c * j.a3));
      const double h1pp = fma(3.0 * j.a1, j.a2, fma(a, j.a3, c * j.a4));
      const double h2p = 2.0 * fma(a, j.a1, c * j.a2);
      const double h2pp = 2.0 * fma(j.a1, j.a1, fma(a, j.a2, c * j.a3));
      const double k1 = fma(a, h1p, c * h1pp);
      const double k2 = fma(a, h1, fma(2.0 * c, h1p, fma(a, h2p, c * h2pp)));
      mean = fma(w3, k1, mean);
      var = fma(w3, fma(2.0, k2, -6.0 * a * h1), var);
    }
  }
}

// Conditional (mean, var) of the Normal transition families (moments.py:182-255; convergence_mf.py:86-107).
MFS_DEV void normal_mean_var(int trans_id, int drift_id, int order, double x, double c, double dt, const double* prm,
                             double& mean, double& var) {
  if (trans_id == MFS_TRANS_NORMAL_AFFINE) {
    mean = prm[0] * x;
    var = prm[1];
  } else {
    const Jet j = drift_jet(drift_id, x, prm);
    if (trans_id == MFS_TRANS_EULER) {
      mean = fma(j.a0, dt, x);
      var = 2.0 * c * dt;
    } else {
      tme_mean_var(j, x, c, dt, order, mean, var);
    }
  }
}

// p(y | x): mfs/one_dim/ss_models.py:43-47, :80-84; tests/test_filtering.py:41-42; mfs/multi_dims/ss_models.py:63-67
// Per-step, node-independent part (computed once per measurement).
struct MeasStep { double y; double c0; };

static __device__ __noinline__ double meas_step_constant(int meas_id, double y, double r) {
  if (meas_id == MFS_MEAS_POISSON_SOFTPLUS) return lgamma(y + 1.0);
  if (meas_id == MFS_MEAS_GAUSSIAN) return 1.0 / (r * 2.5066282746310002);
  return 0.0;
}

// The models that are not on the headline path live out of line: one copy of their exp/log code, not one per node.
static __device__ __noinline__ double measurement_pdf_generic(int meas_id, double y, double c0, double x, double p0,
                                                       double p1) {
  if (meas_id == MFS_MEAS_POISSON_SOFTPLUS) {
    const double mu = log(1.0 + exp(p0 * x));
    const double klogmu = (y == 0.0) ? 0.0 : y * log(mu);   // xlogy
    return exp(klogmu - c0 - mu);
  }
  // MFS_MEAS_GAUSSIAN
  const double zz = (y - p0 * x) / p1;
  return exp(-0.5 * zz * zz) * c0;
}

MFS_DEV double measurement_pdf(int meas_id, const MeasStep& st, double x, const double* prm) {
  if (meas_id == MFS_MEAS_BERNOULLI_LOGISTIC_CUBIC) {
    const double zz = fma(x * x, x * prm[2], -prm[1]);      // prm[2] = 1/c0, filled in by the kernel prologue
    const double p = rcp_fast(1.0 + exp_fast(fmax(fmin(-zz, 708.0), -708.0)));   // 1/(1+inf) = 0 in the reference; 3e-308 here
    return (st.y != 0.0) ? p : 1.0 - p;
  }
  return measurement_pdf_generic(meas_id, st.y, st.c0, x, prm[0], prm[1]);
}

// exp for any argument: the branch-free polynomial path inside |x| <= 708, libdevice outside (overflow / underflow /
// NaN semantics of the reference).
MFS_DEV double exp_any(double x) { return (fabs(x) <= 708.0) ? exp_fast(x) : exp(x); }

// Compile-time measurement kind (MEAS = MFS_MEAS_* or -1 = the run-time switch above): the likelihood of the headline
// models is inlined next to the node loop instead of sitting behind a run-time branch and an out-of-line call
// (ncu r1 N=7 Normal-family capture: 26 % of the issued instructions and the instruction-fetch stalls were there).
constexpr int kMeasRuntime = -1;

template <int MEAS>
MFS_DEV double measurement_pdf_ct(int meas_id, const MeasStep& st, double x, const double* prm) {
  if (MEAS == MFS_MEAS_BERNOULLI_LOGISTIC_CUBIC) {
    return measurement_pdf(MFS_MEAS_BERNOULLI_LOGISTIC_CUBIC, st, x, prm);
  } else if (MEAS == MFS_MEAS_POISSON_SOFTPLUS) {
    // poisson.pmf(y, log(1 + exp(theta2 x))) = exp(xlogy(y, mu) - lgamma(y + 1) - mu)   (ss_models.py:80-84)
    const double mu = log(1.0 + exp_any(prm[0] * x));
    const double klogmu = (st.y == 0.0) ? 0.0 : st.y * log(mu);
    return exp_any(klogmu - st.c0 - mu);
  } else {
    return measurement_pdf(meas_id, st, x, prm);
  }
}

template <int P>
struct Factorial { static constexpr double value = P * Factorial<P - 1>::value; };
template <>
struct Factorial<0> { static constexpr double value = 1.0; };

}  // namespace mfs
