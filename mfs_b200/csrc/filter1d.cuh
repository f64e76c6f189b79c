// The batched 1D moment filter: one filter per thread, the whole time loop inside the kernel, all per-filter state in
// registers.  Restates the scan bodies of mfs/one_dim/filtering.py:73-86 (raw), :140-158 (central), :218-237 (scaled).
#pragma once
#include "models.cuh"

namespace mfs {

#ifndef MFS_BLOCK
#define MFS_BLOCK 128   // threads per CTA (64 / 96 / 256 measured no faster, profiles/r1_occupancy_sweep.md)
#endif
constexpr int kBlock = MFS_BLOCK;

// CTAs of 128 threads per SM that the register allocation must allow.  The step is a long dependent FP64 chain (QL
// rotations), so resident warps are what hides the DFMA latency; measured on B200 (profiles/r1_occupancy_sweep.md):
// N=5 fastest at 6 CTAs/SM (<=80 regs), N=8 at 4 (<=128 regs, spills beyond), N=12 at 3 (<=168 regs).
#ifdef MFS_MIN_BLOCKS
template <int N> constexpr int min_blocks() { return MFS_MIN_BLOCKS; }
#else
template <int N> constexpr int min_blocks() { return N <= 5 ? 6 : N <= 6 ? 5 : N <= 8 ? 4 : 3; }
#endif

// Segmented execution (long horizons).  A filter that loses positive definiteness idles for the rest of the scan, and
// a warp with one live lane costs as much as a full one; so the host cuts the time axis into segments and every
// segment runs only the filters that are still alive, re-packed densely: at the end of a segment a live filter parks
// its state (posterior moments, atoms, mean, scale, nell) in the workspace and appends its id to the next segment's
// list (warp-aggregated atomic; the order of the list is irrelevant, results are per filter).  No host round-trip:
// every launch is sized for the initial batch and threads beyond the device-side count exit at once.
struct SegInfo {
  const int32_t* idx_in;    // filter ids of this segment (nullptr: identity, first segment / unsegmented)
  const int32_t* count_in;  // how many ids (nullptr: P.B)
  int32_t* idx_out;         // ids alive at the end of the segment (nullptr: last segment / unsegmented)
  int32_t* count_out;
  double* state;            // [B][4N + 4] parked filter state
  int64_t t0, t1;           // time steps [t0, t1)
  int32_t defer_nan_fill;   // 1: a failing filter only records its status; nan_fill_kernel writes the NaN tails
};

template <int N>
constexpr int seg_state_doubles() { return 4 * N + 4; }

// Transition kinds the kernel is specialised on (compile time); drift / order / family are runtime switches inside.
enum { KIND_TME = 0, KIND_NORMAL = 1, KIND_BENES_TME = 2 };

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
// ---------------------------------------------------------------------------------------------------------------------
template <int N, int MODE, int KIND>
MFS_DEV void predict(const mfs_filter1d_args& P, const double* tprm, const double (&w)[N], const double (&x)[N],
                     double (&ms)[2 * N], double& mean, double& scale) {
  const double c = 0.5 * P.dispersion * P.dispersion;
  const double dt = P.dt;
  double tn[N];  // Benes: tanh at the nodes, kept between the two passes
  double mu_n[N], var_n[N];  // Normal families, central mode: (mean, var) at the nodes, kept between the two passes

  // pass 1: predicted mean (and scale) -- filtering.py:146-147, :223-224
  if (KIND == KIND_BENES_TME) {
#pragma unroll
    for (int i = 0; i < N; ++i) tn[i] = tanh_fast(x[i]);
  }
  if (MODE != MFS_MODE_RAW) {
    double m_acc = 0.0, v_acc = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double mu, var;
      if (KIND == KIND_BENES_TME) {
        // mean = x + dt tanh x (all orders), var = dt + dt^2 (1 - tanh^2 x) (order >= 2)
        mu = fma(dt, tn[i], x[i]);
        var = (P.tme_order >= 2) ? fma(dt * dt, fma(-tn[i], tn[i], 1.0), dt) : dt;
      } else if (KIND == KIND_NORMAL) {
        // (state_cond_mean of the tme_normal factory is tme.expectation(identity): the same expansion as the mean)
        normal_mean_var(P.trans_id, P.drift_id, P.tme_order, x[i], c, dt, tprm, mu, var);
        mu_n[i] = mu;
        var_n[i] = var;
      } else {
        const Jet j = drift_jet(P.drift_id, x[i], tprm);
        tme_mean_var(j, x[i], c, dt, P.tme_order, mu, var);
      }
      m_acc = fma(w[i], mu, m_acc);
      v_acc = fma(w[i], var, v_acc);
    }
    mean = m_acc;
    if (MODE == MFS_MODE_SCALED) scale = sqrt(v_acc);
  }
  const double sinv = (MODE == MFS_MODE_SCALED) ? 1.0 / scale : 1.0;

  // pass 2: moments
#pragma unroll
  for (int p = 0; p < 2 * N; ++p) ms[p] = 0.0;

  if (KIND == KIND_BENES_TME) {
    // a^2 + a' = 1 and a a' + a''/2 = 0 collapse the expansion (b = 1):  G_k = const_k (k even), const_k tanh(x) (k odd)
    double s0[2 * N], s1[2 * N];
#pragma unroll
    for (int q = 0; q < 2 * N; ++q) { s0[q] = 0.0; s1[q] = 0.0; }
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const double delta = (MODE == MFS_MODE_RAW) ? x[i] : (x[i] - mean) * sinv;
      const double wt = w[i] * tn[i];
      double pw = 1.0;
#pragma unroll
      for (int q = 0; q < 2 * N; ++q) {
        s0[q] = fma(w[i], pw, s0[q]);
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
#pragma unroll
    for (int p = 0; p < 2 * N; ++p) {
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
    double sk[7];
    sk[0] = 1.0;
#pragma unroll
    for (int k = 1; k < 7; ++k) sk[k] = sk[k - 1] * sinv;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const Jet j = drift_jet(P.drift_id, x[i], tprm);
      const TmeCoef cf = tme_coefficients(j, c, dt, P.tme_order);
      double wg[7];
#pragma unroll
      for (int k = 0; k < 7; ++k) wg[k] = w[i] * cf.g[k] * sk[k];
      const double delta = (MODE == MFS_MODE_RAW) ? x[i] : (x[i] - mean) * sinv;
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
    for (int i = 0; i < N; ++i) {
      double mu, var;
      if (MODE == MFS_MODE_RAW) {
        normal_mean_var(P.trans_id, P.drift_id, P.tme_order, x[i], c, dt, tprm, mu, var);
      } else {   // computed by pass 1
        mu = mu_n[i];
        var = var_n[i];
      }
      mu = (MODE == MFS_MODE_RAW) ? mu : mu - mean;
      if (var < 0.0) mu = nan("");  // variance**((p-m)/2) * 0. is NaN for every p >= 1 in the reference
      double m2 = 1.0, m1 = mu;
      ms[0] = fma(w[i], 1.0, ms[0]);
      ms[1] = fma(w[i], m1, ms[1]);
#pragma unroll
      for (int p = 2; p < 2 * N; ++p) {
        const double mp = fma(mu, m1, (double)(p - 1) * var * m2);
        ms[p] = fma(w[i], mp, ms[p]);
        m2 = m1;
        m1 = mp;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Update: ms <- sum_i w_i delta_i^p l_i / c ; returns c = sum_i w_i l_i   (filtering.py:82-85, :151-157, :228-236)
// ---------------------------------------------------------------------------------------------------------------------
template <int N, int MODE>
MFS_DEV double update(const mfs_filter1d_args& P, const double* mprm, double y, double (&w)[N],
                      const double (&x)[N], double (&ms)[2 * N], double& mean, double& scale) {
  double u[N];
  double cc = 0.0;
  MeasStep st;
  st.y = y;
  st.c0 = (P.meas_id == MFS_MEAS_BERNOULLI_LOGISTIC_CUBIC) ? 0.0 : meas_step_constant(P.meas_id, y, mprm[1]);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    u[i] = w[i] * measurement_pdf(P.meas_id, st, x[i], mprm);
    cc += u[i];
  }
  const double cinv = 1.0 / cc;
  double sinv = 1.0;
  if (MODE != MFS_MODE_RAW) {
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) acc = fma(u[i], x[i], acc);
    mean = acc * cinv;
    if (MODE == MFS_MODE_SCALED) {
      double v = 0.0;
#pragma unroll
      for (int i = 0; i < N; ++i) { const double dlt = x[i] - mean; v = fma(u[i], dlt * dlt, v); }
      scale = sqrt(v * cinv);
      sinv = 1.0 / scale;
    }
  }
#pragma unroll
  for (int p = 0; p < 2 * N; ++p) ms[p] = 0.0;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double delta = (MODE == MFS_MODE_RAW) ? x[i] : (x[i] - mean) * sinv;
    double pw = u[i];
#pragma unroll
    for (int p = 0; p < 2 * N; ++p) {
      ms[p] += pw;
      pw *= delta;
    }
  }
#pragma unroll
  for (int p = 0; p < 2 * N; ++p) ms[p] *= cinv;
  // the posterior is the N-atom measure {x_i, u_i / c}: its weights replace the prior ones (see MFS_FLAG_* in the header)
#pragma unroll
  for (int i = 0; i < N; ++i) w[i] = u[i] * cinv;
  return cc;
}

// ---------------------------------------------------------------------------------------------------------------------
template <int N, int MODE, int KIND>
__global__ void __launch_bounds__(kBlock, min_blocks<N>()) filter1d_kernel(const mfs_filter1d_args P, const SegInfo G) {
#ifdef MFS_QL_SMEM
  extern __shared__ double ql_smem[];                 // [3N][kBlock]: (d, e, z) of the eigen-solve, one column per thread
  double* const ql_tile = ql_smem + threadIdx.x;
#endif
  const int64_t slot = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  const int64_t n_active = G.count_in ? (int64_t)__ldg(G.count_in) : P.B;
  if (slot >= n_active) return;
  const int64_t b = G.idx_in ? (int64_t)__ldg(G.idx_in + slot) : slot;
  const bool resume = G.t0 > 0;
  double* park = G.state ? G.state + b * seg_state_doubles<N>() : nullptr;

  double ms[2 * N];
  double mean, scale;
  if (!resume) {
    const double* ms0 = P.ms0 + b * P.ms0_stride;
#pragma unroll
    for (int p = 0; p < 2 * N; ++p) ms[p] = __ldg(ms0 + p);
    mean = (MODE != MFS_MODE_RAW) ? __ldg(P.mean0 + b * P.mean0_stride) : 0.0;
    scale = (MODE == MFS_MODE_SCALED) ? __ldg(P.scale0 + b * P.scale0_stride) : 1.0;
  } else {
#pragma unroll
    for (int p = 0; p < 2 * N; ++p) ms[p] = park[p];
    mean = (MODE != MFS_MODE_RAW) ? park[4 * N] : 0.0;
    scale = (MODE == MFS_MODE_SCALED) ? park[4 * N + 1] : 1.0;
  }
  const double* tprm = P.trans_params + b * P.trans_param_stride;
  const double* mprm = P.meas_params + b * P.meas_param_stride;
  double tp[MFS_MAX_PARAMS], mp[MFS_MAX_PARAMS];
#pragma unroll
  for (int k = 0; k < MFS_MAX_PARAMS; ++k) { tp[k] = __ldg(tprm + k); mp[k] = __ldg(mprm + k); }
  if (P.meas_id == MFS_MEAS_BERNOULLI_LOGISTIC_CUBIC) mp[2] = 1.0 / mp[0];

  const int64_t ys_off = b * P.ys_stride_b;
  double* ms_out = P.ms_out ? P.ms_out + b * P.ms_stride_b : nullptr;
  double* mean_out = P.mean_out ? P.mean_out + b * P.aux_stride_b : nullptr;
  double* scale_out = P.scale_out ? P.scale_out + b * P.aux_stride_b : nullptr;

  double nell = 0.0;
  int status = -1;
  double w[N], x[N];          // quadrature of the current half-step; carried across steps as the posterior atoms
  bool have_atoms = false;
  if (resume) {
#pragma unroll
    for (int i = 0; i < N; ++i) { w[i] = park[2 * N + i]; x[i] = park[3 * N + i]; }
    nell = park[4 * N + 2];
    have_atoms = park[4 * N + 3] != 0.0;
  }
  int64_t t = G.t0;
  double y_next = load_y(P.ys, P.ys_dtype, ys_off + t * P.ys_stride_t);
  for (; t < G.t1; ++t) {
    const double y = y_next;
    if (t + 1 < G.t1) y_next = load_y(P.ys, P.ys_dtype, ys_off + (t + 1) * P.ys_stride_t);

    // two half-steps sharing ONE instance of the quadrature code: phase 0 = prediction, phase 1 = update.
    // From the second step on, phase 0 re-uses the atoms (x_i, w_i l_i / c) of the previous update as its quadrature.
    bool ok = true;
#pragma unroll 1
    for (int phase = 0; phase < 2; ++phase) {
      {
        // ONE instance of the moments -> Jacobi code; with carried atoms only its pivot test is used
        double dj[N], ej[N];
        ok = jacobi_from_moments<N, false>(ms, dj, ej);
        if (ok) {
#ifdef MFS_QL_SMEM
          if (!(phase == 0 && have_atoms)) ok = jacobi_to_rule_smem<N, kBlock>(dj, ej, mean, scale, w, x, ql_tile);
#else
          if (!(phase == 0 && have_atoms)) ok = jacobi_to_rule<N>(dj, ej, mean, scale, w, x);
#endif
        } else if (P.stable) {
          // stable=True: a non-positive pivot is not a failure but the LDL completion of mfs/utils.py:526-538
          ok = moment_quadrature_stable_fallback<N>(ms, mean, scale, w, x);
        }
      }
      if (!ok) break;
      if (phase == 0) {
        predict<N, MODE, KIND>(P, tp, w, x, ms, mean, scale);
      } else {
        const double cc = update<N, MODE>(P, mp, y, w, x, ms, mean, scale);
        nell -= log(cc);
        // stable=True always re-derives the prediction quadrature (its LDL completion is defined on the moments)
        have_atoms = !(P.flags & MFS_FLAG_RECOMPUTE_PREDICT_QUADRATURE) && !P.stable;
      }
    }
    if (!ok) { status = (int)t; break; }

    if (P.out_mode == MFS_OUT_FULL) {
      double* o = ms_out + t * P.ms_stride_t;
#pragma unroll
      for (int p = 0; p < 2 * N; p += 2) *reinterpret_cast<double2*>(o + p) = make_double2(ms[p], ms[p + 1]);
      if (MODE != MFS_MODE_RAW && mean_out) mean_out[t] = mean;
      if (MODE == MFS_MODE_SCALED && scale_out) scale_out[t] = scale;
    }
  }

  if (status >= 0) {
    // JAX semantics: once the Cholesky fails everything downstream is NaN until the end of the scan.
    const double qnan = nan("");
    nell = qnan;
#pragma unroll
    for (int p = 0; p < 2 * N; ++p) ms[p] = qnan;
    mean = qnan;
    scale = qnan;
    if (P.out_mode == MFS_OUT_FULL && !G.defer_nan_fill) {
      for (; t < P.T; ++t) {
        double* o = ms_out + t * P.ms_stride_t;
#pragma unroll
        for (int p = 0; p < 2 * N; p += 2) *reinterpret_cast<double2*>(o + p) = make_double2(qnan, qnan);
        if (MODE != MFS_MODE_RAW && mean_out) mean_out[t] = qnan;
        if (MODE == MFS_MODE_SCALED && scale_out) scale_out[t] = qnan;
      }
    }
  }
  if (status < 0 && G.t1 < P.T) {
    // alive at the end of a segment: park the state and enlist for the next segment
#pragma unroll
    for (int p = 0; p < 2 * N; ++p) park[p] = ms[p];
#pragma unroll
    for (int i = 0; i < N; ++i) { park[2 * N + i] = w[i]; park[3 * N + i] = x[i]; }
    park[4 * N] = mean;
    park[4 * N + 1] = scale;
    park[4 * N + 2] = nell;
    park[4 * N + 3] = have_atoms ? 1.0 : 0.0;
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

// Host-side launcher for one (N, MODE, KIND); defined in filter1d_inst.cu, one translation unit per N.
template <int N>
cudaError_t launch_filter1d(const mfs_filter1d_args& a, const SegInfo& g, int kind, cudaStream_t stream);

}  // namespace mfs
