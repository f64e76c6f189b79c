// The batched 1D moment filter: one filter per thread, the whole time loop inside the kernel.
// Restates the scan bodies of mfs/one_dim/filtering.py:73-86 (raw), :140-158 (central), :218-237 (scaled).
//
// Where the per-filter state lives (round 2): the moments and the eigen-solve (d, e, z) are in registers; the N atoms
// (w_i, x_i) of the current quadrature -- and, per transition kind, one or two values per atom that the prediction
// needs twice (tanh x_i, or the Normal family's (mu_i, var_i)) -- are PARKED IN SHARED MEMORY ([rows][kBlock] tile, one
// column per thread: a warp touches 32 consecutive doubles, conflict-free).  The atoms are dead while the Hankel
// recurrence of the next half-step runs and the prediction only streams them once per node, so keeping them in
// registers made the 128-register build spill ~1.3 KB per thread (ncu r1 v5: 70 local loads + 40 local stores per
// filter-step); parked, the hot loop of the 128-register build is down to ~40 spill instructions per step.
#pragma once
#include "models.cuh"

namespace mfs {

#ifndef MFS_BLOCK
#define MFS_BLOCK 128   // threads per CTA (64 / 96 / 256 measured no faster, profiles/r1_occupancy_sweep.md)
#endif
constexpr int kBlock = MFS_BLOCK;

// CTAs of 128 threads per SM that the register allocation must allow.  The step is a long dependent FP64 chain (QL
// rotations), so resident warps are what hides the DFMA latency.  Since the quadrature atoms live in the shared-memory
// tile (round 2), 128 registers are enough up to N = 14: 4 CTAs instead of 3 gives +6..8 % at N = 9..13, +2 % at N = 14,
// -3 % at N = 15 (profiles/r2_ab_1d_occupancy_large_N.log).  Re-measured with the final kernel (the QL with compile-time
// positions and the rolled atom loops lowered the pressure further; profiles/r2_ab_1d_occupancy_small_N.log): 6 CTAs (80
// registers) instead of 4 give +5 % at N = 8 -- +6 % on the headline case --, +4 % at N = 7; 5 CTAs +2 / +1 % at N = 9 / 10;
// one more CTA than that costs 2..50 % everywhere (spills).
#ifdef MFS_MIN_BLOCKS
template <int N> constexpr int min_blocks() { return MFS_MIN_BLOCKS; }
#else
template <int N> constexpr int min_blocks() { return N <= 8 ? 6 : N <= 10 ? 5 : N <= 14 ? 4 : 3; }
#endif

// Carried filter state: [0, 2N) moments, [2N, 4N) what the next prediction starts from, mean, scale, nell, flag.
// flag: 0 = the moments alone (first step, stable=True, literal recursion); 1 = [2N, 3N) atom weights, [3N, 4N) atom nodes
// of the posterior, which are the prediction quadrature of the next step; 2 = [2N, 4N) the tanh-weighted power sums of
// the posterior atoms (raw moments + Benes TME, see fuses_update_and_predict); -(1 + s) = the filter failed at (absolute)
// step s: everything downstream is NaN.
// Used (a) by the segmented execution below and (b) as the public time-chunked resume state of the C ABI
// (mfs_filter1d_args.carry_in / carry_out).
template <int N>
constexpr int seg_state_doubles() { return 4 * N + 4; }

// Segmented execution (long horizons).  A filter that loses positive definiteness idles for the rest of the scan, and
// a warp with one live lane costs as much as a full one; so the host cuts the time axis into segments and every
// segment runs only the filters that are still alive, re-packed densely: at the end of a segment a live filter parks
// its state in the workspace and appends its id to the next segment's list (warp-aggregated atomic; the order of the
// list is irrelevant, results are per filter).  No host round-trip: every launch is sized for the initial batch and
// threads beyond the device-side count exit at once.
struct SegInfo {
  const int32_t* idx_in;    // filter ids of this segment (nullptr: identity, first segment / unsegmented)
  const int32_t* count_in;  // how many ids (nullptr: P.B)
  int32_t* idx_out;         // ids alive at the end of the segment (nullptr: last segment / unsegmented)
  int32_t* count_out;
  const double* state_in;   // [B][4N + 4] state to resume from (nullptr: start from ms0 / mean0 / scale0)
  double* state_out;        // [B][4N + 4] state at the end of the segment (nullptr: not kept)
  int64_t t0, t1;           // time steps [t0, t1) of this launch, relative to P.ys / the output rows
  int32_t defer_nan_fill;   // 1: a failing filter only records its status; nan_fill_kernel writes the NaN tails
  int32_t last;             // 1: last segment of the call (writes nell / status / LAST outputs)
};

// Transition kinds the kernel is specialised on (compile time); drift / order / family are runtime switches inside.
enum { KIND_TME = 0, KIND_NORMAL = 1, KIND_BENES_TME = 2 };

// rows of the shared-memory tile: w, x, and what the prediction keeps per atom between its two passes
template <int N, int MODE, int KIND>
constexpr int smem_rows() {
#ifdef MFS_FUSE_RAW_BENES
  if (MODE == MFS_MODE_RAW && KIND == KIND_BENES_TME) return 4 * N;      // atoms + the parked power sums s1
#endif
  return MODE == MFS_MODE_RAW ? 2 * N : KIND == KIND_BENES_TME ? 3 * N : KIND == KIND_NORMAL ? 4 * N : 2 * N;
}

// Raw moments + Benes TME (the headline model): the prediction of step t+1 only needs the two power-sum families
// s0[q] = sum_i w_i x_i^q and s1[q] = sum_i w_i tanh(x_i) x_i^q over the posterior atoms of step t -- and s0 IS the
// posterior moment vector the update of step t has just formed.  The update therefore also accumulates s1 while it has
// the powers x_i^q in hand (one more FMA per node and order), and the next prediction is the 7-term TME combination of
// (ms, s1): no pass over the atoms, no power table, 2N x N FMAs + N x 2N multiplies fewer per step, and the atoms never
// leave the update.  (Raw mode only: in the central / scaled modes the powers are taken about the PREDICTED mean.)
//
// Measured on B200 (same-box A/B, profiles/r2_ab_1d_fusion.log): N = 8, T = 100: 3.071e9 vs 3.052e9 filter-steps/s (+0.6 %),
// T = 1000 full history: 4.154e9 vs 4.203e9 (-1.2 %) -- neutral: the step time is set by the dependent-issue latency of the
// QL chase, not by the number of FP64 instructions outside it.  The fusion also changes the rounding of the default path
// (sum_i u_i x^q scaled by 1/c instead of sum_i (u_i / c) x^q), so it is an opt-in build (-DMFS_FUSE_RAW_BENES); the
// GPU test-suite passes with it.
template <int MODE, int KIND>
constexpr bool fuses_update_and_predict() {
#ifdef MFS_FUSE_RAW_BENES
  return MODE == MFS_MODE_RAW && KIND == KIND_BENES_TME;
#else
  return false;
#endif
}

MFS_DEV double load_y(const void* ys, int dtype, int64_t off) {
  if (dtype == MFS_YS_U8) return (double)__ldg(reinterpret_cast<const unsigned char*>(ys) + off);
  if (dtype == MFS_YS_I32) return (double)__ldg(reinterpret_cast<const int*>(ys) + off);
  return __ldg(reinterpret_cast<const double*>(ys) + off);
}

// pq[q] = delta^q / q!  for q = 0..2N-1
template <int N>
MFS_DEV void scaled_powers(double delta, double (&pq)[2 * N]) {
  pq[0] = 1.0;
#pragma unroll
  for (int q = 1; q < 2 * N; ++q) pq[q] = pq[q - 1] * delta * (1.0 / q);
}

template <int N>
MFS_DEV void times_factorials(double (&ms)[2 * N]) {
  double f = 1.0;
#pragma unroll
  for (int p = 2; p < 2 * N; ++p) { f *= (double)p; ms[p] *= f; }
}

// ---------------------------------------------------------------------------------------------------------------------
// Prediction: ms <- sum_i w_i T(x_i), with (mean, scale) <- predicted mean / scale in CENTRAL / SCALED modes.
// The atoms are read from the shared-memory tile `sm` (w_i at row i, x_i at row N + i, scratch rows from 2N).
// ---------------------------------------------------------------------------------------------------------------------
// Loops over the N atoms that read the atoms from the shared-memory tile can stay ROLLED (the accumulators over the 2N
// orders are the unrolled dimension): one copy of the per-atom code (an inlined exp is ~35 instructions) instead of N --
// 1000 instructions (16 KB) less in the hot loop at N = 8.  Same operations in the same order, bit-identical results.
// Same-box A/B (profiles/r2_ab_1d_rolled_atoms.log): N = 8 Benes +0.5 %, central +1.7 %, literal recursion +3.4 %,
// N = 7 Normal family + Poisson +4.3 % (3.23e9 vs 3.09e9), N = 5 -4.8 % -> rolled from N = 7 on.
#ifndef MFS_ROLL_ATOMS_FROM
#define MFS_ROLL_ATOMS_FROM 7
#endif
template <int N>
constexpr bool roll_atoms() { return N >= MFS_ROLL_ATOMS_FROM; }

template <int N, int MODE, int KIND>
MFS_DEV void predict(const mfs_filter1d_args& P, const double* tprm, double* __restrict__ sm, double (&ms)[2 * N],
                     double& mean, double& scale) {
  const double c = 0.5 * P.dispersion * P.dispersion;
  const double dt = P.dt;
#define MFS_W(i) sm[(i) * kBlock]
#define MFS_X(i) sm[(N + (i)) * kBlock]
#define MFS_A(i) sm[(2 * N + (i)) * kBlock]
#define MFS_B(i) sm[(3 * N + (i)) * kBlock]

  // pass 1: predicted mean (and scale) -- filtering.py:146-147, :223-224
  if (MODE != MFS_MODE_RAW) {
    double m_acc = 0.0, v_acc = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const double wi = MFS_W(i), xi = MFS_X(i);
      double mu, var;
      if (KIND == KIND_BENES_TME) {
        // mean = x + dt tanh x (all orders), var = dt + dt^2 (1 - tanh^2 x) (order >= 2)
        const double tn = tanh_fast(xi);
        MFS_A(i) = tn;
        mu = fma(dt, tn, xi);
        var = (P.tme_order >= 2) ? fma(dt * dt, fma(-tn, tn, 1.0), dt) : dt;
      } else if (KIND == KIND_NORMAL) {
        // (state_cond_mean of the tme_normal factory is tme.expectation(identity): the same expansion as the mean)
        normal_mean_var(P.trans_id, P.drift_id, P.tme_order, xi, c, dt, tprm, mu, var);
        MFS_A(i) = mu;
        MFS_B(i) = var;
      } else {
        const Jet j = drift_jet(P.drift_id, xi, tprm);
        tme_mean_var(j, xi, c, dt, P.tme_order, mu, var);
      }
      m_acc = fma(wi, mu, m_acc);
      v_acc = fma(wi, var, v_acc);
    }
    mean = m_acc;
    if (MODE == MFS_MODE_SCALED) scale = sqrt(v_acc);
  }
  const double sinv = (MODE == MFS_MODE_SCALED) ? 1.0 / scale : 1.0;

  // pass 2: moments
  if (KIND == KIND_BENES_TME) {
    // a^2 + a' = 1 and a a' + a''/2 = 0 collapse the expansion (b = 1):  G_k = const_k (k even), const_k tanh(x) (k odd)
    double s0[2 * N], s1[2 * N];
#pragma unroll
    for (int q = 0; q < 2 * N; ++q) { s0[q] = 0.0; s1[q] = 0.0; }
#pragma unroll (roll_atoms<N>() ? 1 : N)
    for (int i = 0; i < N; ++i) {
      const double wi = MFS_W(i), xi = MFS_X(i);
      const double tn = (MODE != MFS_MODE_RAW) ? MFS_A(i) : tanh_fast(xi);
      const double delta = (MODE == MFS_MODE_RAW) ? xi : (xi - mean) * sinv;
      const double wt = wi * tn;
      double pw = 1.0;
#pragma unroll
      for (int q = 0; q < 2 * N; ++q) {
        s0[q] = fma(wi, pw, s0[q]);
        s1[q] = fma(wt, pw, s1[q]);
        pw *= delta;
      }
    }
    {  // s[q] /= q!  (the TME combination below works on delta^q / q!)
      double rf = 1.0;
#pragma unroll
      for (int q = 2; q < 2 * N; ++q) { rf *= 1.0 / q; s0[q] *= rf; s1[q] *= rf; }
    }
    const double dt2 = dt * dt, dt3 = dt2 * dt;
    const bool o2 = P.tme_order >= 2, o3 = P.tme_order >= 3;
    const double si2 = sinv * sinv, si3 = si2 * sinv, si4 = si2 * si2;
    const double g1 = dt * sinv;                                                   // * tanh
    const double g2 = (0.5 * dt + (o2 ? 0.5 * dt2 : 0.0)) * si2;
    const double g3 = ((o2 ? 0.5 * dt2 : 0.0) + (o3 ? dt3 / 6.0 : 0.0)) * si3;      // * tanh
    const double g4 = ((o2 ? 0.125 * dt2 : 0.0) + (o3 ? 0.25 * dt3 : 0.0)) * si4;
    const double g5 = (o3 ? 0.125 * dt3 : 0.0) * si4 * sinv;                        // * tanh
    const double g6 = (o3 ? dt3 / 48.0 : 0.0) * si4 * si2;
    // descending p: s0[p] / s1[p] are dead once ms[p] is formed, so the result grows into the registers they free
#pragma unroll
    for (int p = 2 * N - 1; p >= 0; --p) {
      double acc = s0[p];
      if (p >= 1) acc = fma(g1, s1[p - 1], acc);
      if (p >= 2) acc = fma(g2, s0[p - 2], acc);
      if (p >= 3) acc = fma(g3, s1[p - 3], acc);
      if (p >= 4) acc = fma(g4, s0[p - 4], acc);
      if (p >= 5) acc = fma(g5, s1[p - 5], acc);
      if (p >= 6) acc = fma(g6, s0[p - 6], acc);
      ms[p] = acc;
    }
    times_factorials<N>(ms);
  } else if (KIND == KIND_TME) {
#pragma unroll
    for (int p = 0; p < 2 * N; ++p) ms[p] = 0.0;
    double sk[7];
    sk[0] = 1.0;
#pragma unroll
    for (int k = 1; k < 7; ++k) sk[k] = sk[k - 1] * sinv;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const double wi = MFS_W(i), xi = MFS_X(i);
      const Jet j = drift_jet(P.drift_id, xi, tprm);
      const TmeCoef cf = tme_coefficients(j, c, dt, P.tme_order);
      double wg[7];
#pragma unroll
      for (int k = 0; k < 7; ++k) wg[k] = wi * cf.g[k] * sk[k];
      const double delta = (MODE == MFS_MODE_RAW) ? xi : (xi - mean) * sinv;
      double pq[2 * N];
      scaled_powers<N>(delta, pq);
#pragma unroll
      for (int p = 0; p < 2 * N; ++p) {
        double acc = ms[p];
#pragma unroll
        for (int k = 0; k < 7; ++k)
          if (k <= p) acc = fma(wg[k], pq[p - k], acc);
        ms[p] = acc;
      }
    }
    times_factorials<N>(ms);
  } else {  // KIND_NORMAL: moments of N(mu - mean, var) by M_p = mu M_{p-1} + (p-1) var M_{p-2}  (= moments.py:70-74)
#pragma unroll
    for (int p = 0; p < 2 * N; ++p) ms[p] = 0.0;
    // SCALED: the reference's Normal factories return raw_moment_of_normal(mu - mean, var, p) / prod_k scale^k for EVERY
    // order p (mfs/one_dim/moments.py:205, :243: `/ jnp.prod(scale ** orders)`, not `/ scale ** orders`), so the
    // zeroth "moment" is scale^{-N(2N-1)} rather than 1.  Reproduced literally (bar: identical to the reference).
    double div_all = 1.0;
    if (MODE == MFS_MODE_SCALED) {
      double pr = 1.0, pk = 1.0;
#pragma unroll
      for (int k = 1; k < 2 * N; ++k) { pk *= scale; pr *= pk; }
      div_all = 1.0 / pr;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const double wi = MFS_W(i);
      double mu, var;
      if (MODE == MFS_MODE_RAW) {
        normal_mean_var(P.trans_id, P.drift_id, P.tme_order, MFS_X(i), c, dt, tprm, mu, var);
      } else {   // computed by pass 1
        mu = MFS_A(i);
        var = MFS_B(i);
      }
      mu = (MODE == MFS_MODE_RAW) ? mu : mu - mean;
      if (var < 0.0) mu = nan("");  // variance**((p-m)/2) * 0. is NaN for every p >= 1 in the reference
      double m2 = 1.0, m1 = mu;
      ms[0] = fma(wi, 1.0, ms[0]);
      ms[1] = fma(wi, m1, ms[1]);
#pragma unroll
      for (int p = 2; p < 2 * N; ++p) {
        const double mp = fma(mu, m1, (double)(p - 1) * var * m2);
        ms[p] = fma(wi, mp, ms[p]);
        m2 = m1;
        m1 = mp;
      }
    }
    if (MODE == MFS_MODE_SCALED) {
#pragma unroll
      for (int p = 0; p < 2 * N; ++p) ms[p] *= div_all;
    }
  }
#undef MFS_W
#undef MFS_X
#undef MFS_A
#undef MFS_B
}

// ---------------------------------------------------------------------------------------------------------------------
// Update: ms <- sum_i w_i delta_i^p l_i / c ; returns c = sum_i w_i l_i   (filtering.py:82-85, :151-157, :228-236).
// (w, x) arrive in registers straight from the eigen-solve; the posterior atoms {x_i, w_i l_i / c} leave through the
// shared-memory tile (see MFS_FLAG_* in the header).
// ---------------------------------------------------------------------------------------------------------------------
template <int N, int MODE, int MEAS>
MFS_DEV double update(const mfs_filter1d_args& P, const double* mprm, double y, double (&w)[N],
                      const double (&x)[N], double* __restrict__ sm, double (&ms)[2 * N], double& mean, double& scale) {
  double cc = 0.0;
  MeasStep st;
  st.y = y;
  if (MEAS == MFS_MEAS_BERNOULLI_LOGISTIC_CUBIC) st.c0 = 0.0;
  else if (MEAS == MFS_MEAS_POISSON_SOFTPLUS) st.c0 = log_factorial(y);
  else st.c0 = (P.meas_id == MFS_MEAS_BERNOULLI_LOGISTIC_CUBIC) ? 0.0 : meas_step_constant(P.meas_id, y, mprm[1]);
  if constexpr (roll_atoms<N>()) {
  // atoms through the tile, rolled loops (same operations in the same order as the unrolled register version below)
#pragma unroll
  for (int i = 0; i < N; ++i) { sm[i * kBlock] = w[i]; sm[(N + i) * kBlock] = x[i]; }
#pragma unroll 1
  for (int i = 0; i < N; ++i) {
    const double u = sm[i * kBlock] * measurement_pdf_ct<MEAS>(P.meas_id, st, sm[(N + i) * kBlock], mprm);
    sm[i * kBlock] = u;
    cc += u;
  }
  const double cinv_r = 1.0 / cc;
  double sinv_r = 1.0;
  if (MODE != MFS_MODE_RAW) {
    double acc = 0.0;
#pragma unroll 1
    for (int i = 0; i < N; ++i) acc = fma(sm[i * kBlock], sm[(N + i) * kBlock], acc);
    mean = acc * cinv_r;
    if (MODE == MFS_MODE_SCALED) {
      double v = 0.0;
#pragma unroll 1
      for (int i = 0; i < N; ++i) { const double dlt = sm[(N + i) * kBlock] - mean; v = fma(sm[i * kBlock], dlt * dlt, v); }
      scale = sqrt(v * cinv_r);
      sinv_r = 1.0 / scale;
    }
  }
#pragma unroll
  for (int p = 0; p < 2 * N; ++p) ms[p] = 0.0;
#pragma unroll 1
  for (int i = 0; i < N; ++i) {
    const double xi = sm[(N + i) * kBlock];
    const double delta = (MODE == MFS_MODE_RAW) ? xi : (xi - mean) * sinv_r;
    double pw = sm[i * kBlock];
    sm[i * kBlock] = pw * cinv_r;
#pragma unroll
    for (int p = 0; p < 2 * N; ++p) {
      ms[p] += pw;
      pw *= delta;
    }
  }
#pragma unroll
  for (int p = 0; p < 2 * N; ++p) ms[p] *= cinv_r;
  return cc;
  } else {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    w[i] = w[i] * measurement_pdf_ct<MEAS>(P.meas_id, st, x[i], mprm);   // u_i = w_i l_i
    cc += w[i];
  }
  const double cinv = 1.0 / cc;
  double sinv = 1.0;
  if (MODE != MFS_MODE_RAW) {
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) acc = fma(w[i], x[i], acc);
    mean = acc * cinv;
    if (MODE == MFS_MODE_SCALED) {
      double v = 0.0;
#pragma unroll
      for (int i = 0; i < N; ++i) { const double dlt = x[i] - mean; v = fma(w[i], dlt * dlt, v); }
      scale = sqrt(v * cinv);
      sinv = 1.0 / scale;
    }
  }
#pragma unroll
  for (int p = 0; p < 2 * N; ++p) ms[p] = 0.0;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double delta = (MODE == MFS_MODE_RAW) ? x[i] : (x[i] - mean) * sinv;
    double pw = w[i];
    sm[i * kBlock] = pw * cinv;        // posterior weight; the node stays where the eigen-solve put it
#pragma unroll
    for (int p = 0; p < 2 * N; ++p) {
      ms[p] += pw;
      pw *= delta;
    }
  }
#pragma unroll
  for (int p = 0; p < 2 * N; ++p) ms[p] *= cinv;
  return cc;
  }
}

// Fused update (see fuses_update_and_predict): posterior raw moments AND the tanh-weighted power sums of the posterior
// atoms.  (u_i, x_i) are parked in rows [0, 2N) and streamed back, so that only the 4N accumulators live in registers;
// s1 leaves through rows [2N, 4N).  Returns c = sum_i w_i l_i.
template <int N, int MEAS>
MFS_DEV double update_fused_raw_benes(const mfs_filter1d_args& P, const double* mprm, double y, double (&w)[N],
                                      const double (&x)[N], double* __restrict__ sm, double (&ms)[2 * N]) {
  static_assert(MEAS == MFS_MEAS_BERNOULLI_LOGISTIC_CUBIC, "the Benes TME kind is instantiated with the Bernoulli likelihood");
  double cc = 0.0;
  MeasStep st;
  st.y = y;
  st.c0 = 0.0;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double u = w[i] * measurement_pdf_ct<MEAS>(P.meas_id, st, x[i], mprm);
    sm[i * kBlock] = u;
    sm[(N + i) * kBlock] = x[i];
    cc += u;
  }
  const double cinv = 1.0 / cc;
  double s1[2 * N];
#pragma unroll
  for (int p = 0; p < 2 * N; ++p) { ms[p] = 0.0; s1[p] = 0.0; }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double xi = sm[(N + i) * kBlock];
    const double tn = tanh_fast(xi);
    double pw = sm[i * kBlock];
#pragma unroll
    for (int p = 0; p < 2 * N; ++p) {
      ms[p] += pw;
      s1[p] = fma(pw, tn, s1[p]);
      pw *= xi;
    }
  }
#pragma unroll
  for (int p = 0; p < 2 * N; ++p) {
    ms[p] *= cinv;
    sm[(2 * N + p) * kBlock] = s1[p] * cinv;
  }
  return cc;
}

// Fused prediction: ms <- TME combination of the posterior moments (= s0) and the parked s1 (rows [2N, 4N)).
template <int N>
MFS_DEV void predict_fused_raw_benes(const mfs_filter1d_args& P, const double* __restrict__ sm, double (&ms)[2 * N]) {
  const double dt = P.dt;
  double s0[2 * N], s1[2 * N];
  {
    double rf = 1.0;
#pragma unroll
    for (int q = 0; q < 2 * N; ++q) {
      if (q >= 2) rf *= 1.0 / q;
      s0[q] = (q >= 2) ? ms[q] * rf : ms[q];
      const double v = sm[(2 * N + q) * kBlock];
      s1[q] = (q >= 2) ? v * rf : v;
    }
  }
  const double dt2 = dt * dt, dt3 = dt2 * dt;
  const bool o2 = P.tme_order >= 2, o3 = P.tme_order >= 3;
  const double g1 = dt;                                                        // * tanh
  const double g2 = 0.5 * dt + (o2 ? 0.5 * dt2 : 0.0);
  const double g3 = (o2 ? 0.5 * dt2 : 0.0) + (o3 ? dt3 / 6.0 : 0.0);           // * tanh
  const double g4 = (o2 ? 0.125 * dt2 : 0.0) + (o3 ? 0.25 * dt3 : 0.0);
  const double g5 = o3 ? 0.125 * dt3 : 0.0;                                    // * tanh
  const double g6 = o3 ? dt3 / 48.0 : 0.0;
#pragma unroll
  for (int p = 2 * N - 1; p >= 0; --p) {
    double acc = s0[p];
    if (p >= 1) acc = fma(g1, s1[p - 1], acc);
    if (p >= 2) acc = fma(g2, s0[p - 2], acc);
    if (p >= 3) acc = fma(g3, s1[p - 3], acc);
    if (p >= 4) acc = fma(g4, s0[p - 4], acc);
    if (p >= 5) acc = fma(g5, s1[p - 5], acc);
    if (p >= 6) acc = fma(g6, s0[p - 6], acc);
    ms[p] = acc;
  }
  times_factorials<N>(ms);
}

// The eigen-solve keeps (d, e, z) in registers.  -DMFS_Z_SMEM moves its first-row vector z to the (then dead) weight
// rows of the shared-memory tile (16 registers fewer at N = 8); measured slower on B200 at every occupancy
// (profiles/r2_ab_1d_kernel.md: N = 8, T = 100: 2.86e9 vs 3.01e9 filter-steps/s), kept as a build option.
#ifdef MFS_Z_SMEM
#define MFS_QUADRATURE(ms, mean, scale, w, x, ldl) moment_quadrature_zs<N, kBlock>(ms, mean, scale, w, x, ldl, sm)
#else
#define MFS_QUADRATURE(ms, mean, scale, w, x, ldl) moment_quadrature<N>(ms, mean, scale, w, x, ldl)
#endif

// (mean, variance) of the filtering distribution from the carried representation -- MFS_OUT_MEANVAR.
template <int N, int MODE>
MFS_DEV double2 mean_var(const double (&ms)[2 * N], double mean, double scale) {
  if (MODE == MFS_MODE_RAW) return make_double2(ms[1], fma(-ms[1], ms[1], ms[2]));
  if (MODE == MFS_MODE_CENTRAL) return make_double2(mean, ms[2]);
  return make_double2(mean, scale * scale * ms[2]);
}

// One CTA barrier per time step?  Measured on B200 (profiles/r2_ab_1d_step_barrier.log): it pays where the code of a step is
// large -- the Normal family (+6 % at N = 7, +6..13 % at N = 9, 12), Benes N = 9: +1 %, N = 10..14: +2..8 % at the shipped
// 4 CTAs per SM -- and costs 0.5-4 % for the small Benes instances (N = 2: -29 %).
template <int N, int KIND>
constexpr bool step_barrier() {
#ifdef MFS_1D_STEP_BARRIER
  return MFS_1D_STEP_BARRIER != 0;
#else
  return KIND == KIND_NORMAL || N >= 9;
#endif
}

// ---------------------------------------------------------------------------------------------------------------------
template <int N, int MODE, int KIND, int MEAS>
__global__ void __launch_bounds__(kBlock, min_blocks<N>()) filter1d_kernel(const mfs_filter1d_args P, const SegInfo G) {
  extern __shared__ double smem_tile[];                 // [smem_rows][kBlock], one column per thread
  double* const sm = smem_tile + threadIdx.x;
  const int64_t slot = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  const int64_t n_active = G.count_in ? (int64_t)__ldg(G.count_in) : P.B;
  // One CTA barrier per time step (step_barrier<N, KIND>()): keeps the CTA's four warps on the same code, so that one
  // instruction fetch serves all of them -- the 2-D kernel's lever (filter_nd.cuh), which pays here where the step's code
  // is large.  Every thread of a CTA that has at least one filter then runs the whole time loop; `alive` says whether it
  // still has work (a thread beyond the batch reads the last filter's inputs and writes nothing).
  constexpr bool kBar = step_barrier<N, KIND>();
  bool present = true;
  int64_t slot_rd = slot;
  if constexpr (kBar) {
    if ((int64_t)blockIdx.x * kBlock >= n_active) return;
    present = slot < n_active;
    slot_rd = present ? slot : n_active - 1;
  } else {
    if (slot >= n_active) return;
  }
  const int64_t b = G.idx_in ? (int64_t)__ldg(G.idx_in + slot_rd) : slot_rd;
  const double* park_in = G.state_in ? G.state_in + b * seg_state_doubles<N>() : nullptr;
  double* park_out = G.state_out ? G.state_out + b * seg_state_doubles<N>() : nullptr;

  // parameter grid over shared records: filter b = row * grid_records + rec
  const int64_t row = P.grid_records > 0 ? b / P.grid_records : b;
  const int64_t rec = P.grid_records > 0 ? b - row * P.grid_records : b;

  double ms[2 * N];
  double mean, scale;
  double nell = 0.0;
  int status = -1;
  // what the next prediction starts from: 0 = the moments (a full quadrature), 1 = the posterior atoms in rows [0, 2N),
  // 2 = the posterior moments + the tanh-weighted power sums in rows [2N, 4N) (fuses_update_and_predict)
  constexpr bool kFuse = fuses_update_and_predict<MODE, KIND>();
  int pred_from = 0;
  if (!park_in) {
    const double* ms0 = P.ms0 + row * P.ms0_stride;
#pragma unroll
    for (int p = 0; p < 2 * N; ++p) ms[p] = __ldg(ms0 + p);
    mean = (MODE != MFS_MODE_RAW) ? __ldg(P.mean0 + row * P.mean0_stride) : 0.0;
    scale = (MODE == MFS_MODE_SCALED) ? __ldg(P.scale0 + row * P.scale0_stride) : 1.0;
  } else {
#pragma unroll
    for (int p = 0; p < 2 * N; ++p) ms[p] = park_in[p];
    mean = (MODE != MFS_MODE_RAW) ? park_in[4 * N] : 0.0;
    scale = (MODE == MFS_MODE_SCALED) ? park_in[4 * N + 1] : 1.0;
    nell = park_in[4 * N + 2];
    const double flag = park_in[4 * N + 3];
    pred_from = flag == 1.0 ? 1 : (flag == 2.0 && kFuse) ? 2 : 0;
    const int row0 = pred_from == 2 ? 2 * N : 0;       // atoms (w rows, then x rows) or the power sums s1
#pragma unroll
    for (int i = 0; i < 2 * N; ++i) sm[(row0 + i) * kBlock] = park_in[2 * N + i];
    if (flag < 0.0) status = (int)(-flag) - 1;   // failed in an earlier chunk: NaN from the first step of this one
  }
  const double* tprm = P.trans_params + row * P.trans_param_stride;
  const double* mprm = P.meas_params + row * P.meas_param_stride;
  double tp[MFS_MAX_PARAMS], mp[MFS_MAX_PARAMS];
#pragma unroll
  for (int k = 0; k < MFS_MAX_PARAMS; ++k) { tp[k] = __ldg(tprm + k); mp[k] = __ldg(mprm + k); }
  if (MEAS == MFS_MEAS_BERNOULLI_LOGISTIC_CUBIC || (MEAS == kMeasRuntime && P.meas_id == MFS_MEAS_BERNOULLI_LOGISTIC_CUBIC))
    mp[2] = 1.0 / mp[0];

  const int64_t ys_off = rec * P.ys_stride_b;
  double* ms_out = P.ms_out ? P.ms_out + b * P.ms_stride_b : nullptr;
  double* mean_out = P.mean_out ? P.mean_out + b * P.aux_stride_b : nullptr;
  double* scale_out = P.scale_out ? P.scale_out + b * P.aux_stride_b : nullptr;

  int64_t t = G.t0;
  bool alive = present && status < 0;
  int64_t t_stop = G.t0;                 // kBar: where this thread's filter stopped (NaN from there on)
  if ((kBar || alive) && t < G.t1) {
    double y_next = alive ? load_y(P.ys, P.ys_dtype, ys_off + t * P.ys_stride_t) : 0.0;
    for (; t < G.t1; ++t) {
      if constexpr (kBar) {
        __syncthreads();
        if (!alive) continue;
      }
      const double y = y_next;
      if (t + 1 < G.t1) y_next = load_y(P.ys, P.ys_dtype, ys_off + (t + 1) * P.ys_stride_t);

      // Half-step 1, prediction (filtering.py:78-80).  From the second step on its quadrature is known: the posterior
      // of the previous update IS the N-atom measure parked in shared memory, so only the Hankel pivots of the
      // posterior moments are tested (the same quantity the reference's Cholesky fails on -> NaN).  The literal
      // recursion (first step, MFS_FLAG_RECOMPUTE_PREDICT_QUADRATURE, stable=True) derives the rule from the moments.
      bool ok;
      if (pred_from != 0) {
        double dj[N], ej[N];
        ok = jacobi_from_moments<N, false>(ms, dj, ej);          // pivot test only: ej is dead code here
      } else {
        double w[N], x[N];
        ok = MFS_QUADRATURE(ms, mean, scale, w, x, P.stable != 0);
#pragma unroll
        for (int i = 0; i < N; ++i) { sm[i * kBlock] = w[i]; sm[(N + i) * kBlock] = x[i]; }
      }
      if (ok) {
        bool predicted = false;
        if constexpr (kFuse) {
          if (pred_from == 2) { predict_fused_raw_benes<N>(P, sm, ms); predicted = true; }
        }
        if (!predicted) predict<N, MODE, KIND>(P, tp, sm, ms, mean, scale);
        // Half-step 2, update (filtering.py:82-86)
        double w[N], x[N];
        ok = MFS_QUADRATURE(ms, mean, scale, w, x, P.stable != 0);
        if (ok) {
          // stable=True always re-derives the prediction quadrature (its LDL completion is defined on the moments)
          const bool literal = (P.flags & MFS_FLAG_RECOMPUTE_PREDICT_QUADRATURE) || P.stable;
          double cc = 0.0;
          bool updated = false;
          if constexpr (kFuse) {
            if (!literal) {
              cc = update_fused_raw_benes<N, MEAS>(P, mp, y, w, x, sm, ms);
              pred_from = 2;
              updated = true;
            }
          }
          if (!updated) {
#pragma unroll
            for (int i = 0; i < N; ++i) sm[(N + i) * kBlock] = x[i];
            cc = update<N, MODE, MEAS>(P, mp, y, w, x, sm, ms, mean, scale);
            pred_from = literal ? 0 : 1;
          }
          nell -= log(cc);
        }
      }
      if (!ok) {
        status = (int)(P.t_offset + t);
        if constexpr (kBar) {
          alive = false;
          t_stop = t;
          continue;
        } else {
          break;
        }
      }

      if (P.out_mode == MFS_OUT_FULL) {
        double* o = ms_out + t * P.ms_stride_t;
#pragma unroll
        for (int p = 0; p < 2 * N; p += 2) *reinterpret_cast<double2*>(o + p) = make_double2(ms[p], ms[p + 1]);
        if (MODE != MFS_MODE_RAW && mean_out) mean_out[t] = mean;
        if (MODE == MFS_MODE_SCALED && scale_out) scale_out[t] = scale;
      } else if (P.out_mode == MFS_OUT_MEANVAR) {
        *reinterpret_cast<double2*>(ms_out + t * P.ms_stride_t) = mean_var<N, MODE>(ms, mean, scale);
      }
    }
  }

  if constexpr (kBar) {
    if (!present) return;
    if (status >= 0) t = t_stop;
  }
  if (status >= 0) {
    // JAX semantics: once the Cholesky fails everything downstream is NaN until the end of the scan.
    const double qnan = nan("");
    nell = qnan;
#pragma unroll
    for (int p = 0; p < 2 * N; ++p) ms[p] = qnan;
    mean = qnan;
    scale = qnan;
    if ((P.out_mode == MFS_OUT_FULL || P.out_mode == MFS_OUT_MEANVAR) && !G.defer_nan_fill) {
      for (; t < P.T; ++t) {
        double* o = ms_out + t * P.ms_stride_t;
        if (P.out_mode == MFS_OUT_MEANVAR) {
          *reinterpret_cast<double2*>(o) = make_double2(qnan, qnan);
          continue;
        }
#pragma unroll
        for (int p = 0; p < 2 * N; p += 2) *reinterpret_cast<double2*>(o + p) = make_double2(qnan, qnan);
        if (MODE != MFS_MODE_RAW && mean_out) mean_out[t] = qnan;
        if (MODE == MFS_MODE_SCALED && scale_out) scale_out[t] = qnan;
      }
    }
    if (P.carry_out) {   // a failed filter stays failed in the next chunk of a time-chunked run
      double* co = P.carry_out + b * seg_state_doubles<N>();
#pragma unroll
      for (int p = 0; p < 4 * N + 3; ++p) co[p] = qnan;
      co[4 * N + 3] = -(double)(status + 1);
    }
  } else if (park_out) {
    // alive at the end of a segment (or of a chunk: carry_out): park the state
#pragma unroll
    for (int p = 0; p < 2 * N; ++p) park_out[p] = ms[p];
    const int row0 = pred_from == 2 ? 2 * N : 0;
#pragma unroll
    for (int i = 0; i < 2 * N; ++i) park_out[2 * N + i] = sm[(row0 + i) * kBlock];
    park_out[4 * N] = mean;
    park_out[4 * N + 1] = scale;
    park_out[4 * N + 2] = nell;
    park_out[4 * N + 3] = (double)pred_from;
  }
  if (status < 0 && !G.last) {
    // alive at the end of a segment: enlist for the next one
    const unsigned peers = __activemask();           // whichever lanes arrive together share one atomic
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(peers) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(G.count_out, __popc(peers));
    base = __shfl_sync(peers, base, leader);
    G.idx_out[base + __popc(peers & ((1u << lane) - 1u))] = (int32_t)b;
    return;
  }
  if (P.out_mode == MFS_OUT_LAST) {
#pragma unroll
    for (int p = 0; p < 2 * N; p += 2) *reinterpret_cast<double2*>(ms_out + p) = make_double2(ms[p], ms[p + 1]);
    if (MODE != MFS_MODE_RAW && mean_out) mean_out[0] = mean;
    if (MODE == MFS_MODE_SCALED && scale_out) scale_out[0] = scale;
  }
  P.nell_out[b] = nell;
  if (P.status_out) P.status_out[b] = status;
}

// Host-side launcher for one (N, MODE, KIND, MEAS); defined in filter1d_inst.cu, one translation unit per N.
template <int N>
cudaError_t launch_filter1d(const mfs_filter1d_args& a, const SegInfo& g, int kind, int meas_ct, cudaStream_t stream);

}  // namespace mfs
