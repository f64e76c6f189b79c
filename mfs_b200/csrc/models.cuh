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

// log(x) for positive normal x with constant-bank coefficients (same motivation as exp_fast): x = m 2^e with m in
// [sqrt(1/2), sqrt(2)), s = (m - 1)/(m + 1), log m = 2 atanh(s) = 2 s (1 + z P(z)), z = s^2 (degree-7 Chebyshev interpolant,
// approximation error 3e-20, tools/fit_log_poly.py).  Zero, denormal, negative, inf and NaN arguments take libdevice.
static __constant__ double kLogPoly[8] = {
    0x1.5555555555555p-2, 0x1.9999999999a3ap-3, 0x1.2492492476765p-3, 0x1.c71c7201f207cp-4,
    0x1.745cf8c4c7bcap-4, 0x1.3b1c42e3df02cp-4, 0x1.0fbd2bafa02d6p-4, 0x1.0c1263e1bac13p-4};

MFS_DEV double log_fast(double x) {
  if (!(x >= 2.2250738585072014e-308 && x <= 1.7976931348623157e308)) return log(x);
  int hi = __double2hiint(x);
  int e = (hi >> 20) - 1023;
  double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));     // [1, 2)
  if (m > 1.4142135623730951) { m *= 0.5; e += 1; }
  const double s = (m - 1.0) * rcp_fast(m + 1.0);
  const double z = s * s;
  double p = kLogPoly[7];
#pragma unroll
  for (int j = 6; j >= 0; --j) p = fma(p, z, kLogPoly[j]);
  const double two_s = s + s;
  const double ef = (double)e;
  // e ln2_hi + (2 s + 2 s z P + e ln2_lo)
  return fma(ef, 0x1.62e42fefa39efp-1, fma(two_s * z, p, fma(ef, 0x1.abc9e3b39803fp-56, two_s)));
}

// log(k!) for k = 0..170 (lgamma(k + 1) of the Poisson pmf; libdevice's lgamma is ~150 instructions per call)
static __device__ const double kLogFactorial[171] = {
    0x0.0p+0, 0x0.0p+0, 0x1.62e42fefa39efp-1, 0x1.cab0bfa2a2002p+0,
    0x1.96ca77c922cf9p+1, 0x1.326643c4479c9p+2, 0x1.a51273acf01cap+2, 0x1.10ce1f32dcc30p+3,
    0x1.5358e82fcb70dp+3, 0x1.99a8921a7f7cfp+3, 0x1.e357590954d15p+3, 0x1.180973f3a8d74p+4,
    0x1.3fcba16d50143p+4, 0x1.68d5a9c3b32cep+4, 0x1.930f3df162a42p+4, 0x1.be636a63fd346p+4,
    0x1.eabff061f1a84p+4, 0x1.0c0a63f2f353ap+5, 0x1.2329df2d5ee52p+5, 0x1.3ab8153363985p+5,
    0x1.52af57aed77bep+5, 0x1.6b0a8643472a9p+5, 0x1.83c4faba84f06p+5, 0x1.9cda78b856a45p+5,
    0x1.b6472034e8d14p+5, 0x1.d007622cd65e7p+5, 0x1.ea17f717c6794p+5, 0x1.023aeb67e4fefp+6,
    0x1.0f8f18d330240p+6, 0x1.1d07353917231p+6, 0x1.2aa208b59d0e5p+6, 0x1.385e6fd9e5a40p+6,
    0x1.463b59b942084p+6, 0x1.5437c633ace4ap+6, 0x1.6252c474896bap+6, 0x1.708b719e11658p+6,
    0x1.7ee0f79b26758p+6, 0x1.8d528c1243d96p+6, 0x1.9bdf6f75257a3p+6, 0x1.aa86ec2969812p+6,
    0x1.b94855c702ba2p+6, 0x1.c8230869ca105p+6, 0x1.d7166813e12eep+6, 0x1.e621e01eeba4fp+6,
    0x1.f544e2ba69cf1p+6, 0x1.023f743addd9fp+7, 0x1.09e7b7ea41ea9p+7, 0x1.119afe762626bp+7,
    0x1.19590c853a559p+7, 0x1.2121a930c6ec3p+7, 0x1.28f49ddeb1f31p+7, 0x1.30d1b61e86335p+7,
    0x1.38b8bf8931ddbp+7, 0x1.40a989a33a6cdp+7, 0x1.48a3e5c12af19p+7, 0x1.50a7a6ee08711p+7,
    0x1.58b4a1d39da73p+7, 0x1.60caaca474746p+7, 0x1.68e99f0757979p+7, 0x1.711152043b2c4p+7,
    0x1.79419ff26dc59p+7, 0x1.817a6467f6fb9p+7, 0x1.89bb7c2a0aea1p+7, 0x1.9204c51e7c761p+7,
    0x1.9a561e3e1a4bdp+7, 0x1.a2af6787e4609p+7, 0x1.ab1081f509726p+7, 0x1.b3794f6d9d7afp+7,
    0x1.bbe9b2bdfb621p+7, 0x1.c4618f8cc56f7p+7, 0x1.cce0ca5179100p+7, 0x1.d567484b8b7b6p+7,
    0x1.ddf4ef7a05a70p+7, 0x1.e689a69396befp+7, 0x1.ef2554ff15148p+7, 0x1.f7c7e2cc66183p+7,
    0x1.00389c56e3462p+8, 0x1.04909ff8b652bp+8, 0x1.08ebf13dbf263p+8, 0x1.0d4a85602b129p+8,
    0x1.11ac51df8932ap+8, 0x1.16114c7e34736p+8, 0x1.1a796b3ede1acp+8, 0x1.1ee4a46236d3ep+8,
    0x1.2352ee64b46d5p+8, 0x1.27c43ffc72962p+8, 0x1.2c3890172d057p+8, 0x1.30afd5d851956p+8,
    0x1.352a089728f1bp+8, 0x1.39a71fdd14947p+8, 0x1.3e271363e0df7p+8, 0x1.42a9db142a36ap+8,
    0x1.472f6f03d410cp+8, 0x1.4bb7c77491066p+8, 0x1.5042dcd27af64p+8, 0x1.54d0a7b2ba658p+8,
    0x1.596120d23c4ecp+8, 0x1.5df4411475a1cp+8, 0x1.628a018233bedp+8, 0x1.67225b4879462p+8,
    0x1.6bbd47b7669b6p+8, 0x1.705ac0412d89fp+8, 0x1.74fabe790f7bep+8, 0x1.799d3c1265c0ep+8,
    0x1.7e4232dfb367dp+8, 0x1.82e99cd1c0368p+8, 0x1.879373f6bc4fep+8, 0x1.8c3fb2796c21cp+8,
    0x1.90ee52a05c35fp+8, 0x1.959f4ecd1c8b3p+8, 0x1.9a52a17b831ccp+8, 0x1.9f084540f545ep+8,
    0x1.a3c034cbb7b2cp+8, 0x1.a87a6ae24493ap+8, 0x1.ad36e262a7cc0p+8, 0x1.b1f59641e0db5p+8,
    0x1.b6b6818b4a3ebp+8, 0x1.bb799f600610ap+8, 0x1.c03eeaf66facdp+8, 0x1.c5065f9992226p+8,
    0x1.c9cff8a8a340dp+8, 0x1.ce9bb196830eap+8, 0x1.d36985e93f7b8p+8, 0x1.d83971399c213p+8,
    0x1.dd0b6f329dea4p+8, 0x1.e1df7b911a74cp+8, 0x1.e6b592234b0c9p+8, 0x1.eb8daec863182p+8,
    0x1.f067cd7029d4dp+8, 0x1.f543ea1a97428p+8, 0x1.fa2200d7741ebp+8, 0x1.ff020dc5fcd0cp+8,
    0x1.01f2068a4395cp+9, 0x1.0463fd801573cp+9, 0x1.06d6e9ea365edp+9, 0x1.094ac9f576038p+9,
    0x1.0bbf9bd589663p+9, 0x1.0e355dc4e4164p+9, 0x1.10ac0e0492828p+9, 0x1.1323aadc1563ep+9,
    0x1.159c32993e34fp+9, 0x1.1815a3900cac1p+9, 0x1.1a8ffc1a8d2fep+9, 0x1.1d0b3a98b83c1p+9,
    0x1.1f875d7052afep+9, 0x1.2204630ccefc3p+9, 0x1.248249df2f2b1p+9, 0x1.2701105de7b8dp+9,
    0x1.2980b504c3372p+9, 0x1.2c013654c6b40p+9, 0x1.2e8292d416dddp+9, 0x1.3104c90dddddep+9,
    0x1.3387d79231e3dp+9, 0x1.360bbcf5fc5bfp+9, 0x1.389077d2e1cb2p+9, 0x1.3b1606c72a4a4p+9,
    0x1.3d9c6875aa9cfp+9, 0x1.40239b85adddfp+9, 0x1.42ab9ea2dfbd1p+9, 0x1.4534707d3748fp+9,
    0x1.47be0fc8e241ep+9, 0x1.4a487b3e30effp+9, 0x1.4cd3b19982794p+9, 0x1.4f5fb19b31b3fp+9,
    0x1.51ec7a0782708p+9, 0x1.547a09a68f387p+9, 0x1.57085f44377dfp+9, 0x1.599779b00e38ep+9,
    0x1.5c2757bd48ee8p+9, 0x1.5eb7f842af200p+9, 0x1.61495a1a8a1d5p+9};

MFS_DEV double log_factorial(double y) {
  return (y >= 0.0 && y <= 170.0 && y == floor(y)) ? __ldg(&kLogFactorial[(int)y]) : lgamma(y + 1.0);
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
    const double mu = log_fast(1.0 + exp_any(prm[0] * x));
    const double klogmu = (st.y == 0.0) ? 0.0 : st.y * log_fast(mu);
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
