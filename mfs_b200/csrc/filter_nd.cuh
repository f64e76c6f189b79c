// The batched 2-dimensional moment filter (prey--predator / Lotka--Volterra), one filter per WARP, matrices in shared
// memory.  Restates mfs/multi_dims/filtering.py:258-277 (central) / :325-341 (raw) and
// mfs/multi_dims/quadratures.py:149-178.
//
// Per half-step (quadrature):
//   G = ms[inds[0]] (s x s, s = C(N+1, 2)); R = chol(G) in place (lane-parallel right-looking);
//   for k = 0, 1:  K_k = R^-1 ms[inds[1+k]] R^-T  (forward substitutions, one lane per column / row), symmetrised;
//   both K_k are diagonalised TOGETHER by a cyclic Jacobi eigenvalue iteration with round-robin ordering (the K_k are
//   dense, not tridiagonal, in d >= 2): the up-to-(s/2) disjoint rotations of a round for both matrices are computed
//   by different lanes and applied to rows, then columns, then eigenvector columns in parallel;
//   nodes (lam1_i, lam2_j) + mean, weights <v1_i, v2_j> v1_i[0] v2_j[0]  (s^2 signed weights).
// Prediction uses the Normal transition families (Euler--Maruyama, TME + Normal of order 1/2) of the Lotka--Volterra
// SDE; the product moments of N(mu - m, C) come from the recursion E[x^{n+e_i}] = mu_i E[x^n] + sum_j C_ij n_j
// E[x^{n-e_j}], which equals the Kan--Magnus sum of mfs/multi_dims/moments.py:111-154 (checked in the tests).
// Eigenvector signs/order are arbitrary; every consumer is a sum over all s^2 nodes, so they cancel.
#pragma once
#include "models.cuh"

namespace mfs {

// warps (= filters) per CTA
#ifndef MFS_ND_WARPS_LARGE
#define MFS_ND_WARPS_LARGE 4    // N >= 6 (255 registers: 8 warps per SM): 4 = two CTAs, 8 = one
#endif
template <int N> constexpr int nd_warps() {
#ifdef MFS_ND_WARPS
  return N <= 5 ? MFS_ND_WARPS : MFS_ND_WARPS_LARGE;
#else
  return N == 5 ? 8 : N >= 6 ? MFS_ND_WARPS_LARGE : 4;      // N = 5: 2 CTAs x 8 warps at 128 registers (profiles/r2_ab_nd_barrier.log)
#endif
}
// deflation cascade of the register QL (quadrature.cuh): 0 = inlined per exit, 1 = shared guarded levels, 2 = shared switch
#ifndef MFS_ND_CASCADE_SMALL
#define MFS_ND_CASCADE_SMALL 0
#endif
#ifndef MFS_ND_CASCADE_LARGE
#define MFS_ND_CASCADE_LARGE 0
#endif
// re-associated rotation with the 7-deep dependency chain in the register QL (quadrature.cuh: ql_chase7) for S >= this.
// Measured (profiles/r2_ab_nd_batched_chain7.log): N = 7 (4 warps per SM, latency bound) +3 %; N = 3..6: -2 .. -8 %.
#ifndef MFS_ND_QL_CHAIN7_MIN_S
#define MFS_ND_QL_CHAIN7_MIN_S 28
#endif
// phase barriers per time step in filter_nd_kernel (0 = none; see the kernel).  Measured (profiles/r2_ab_nd_barrier.log):
// N = 2: none is best (-1.5 % with any); N = 3: +6.5 % with 2, no more with 4; N = 4, 5: +15 / +22 % with 4;
// N = 6, 7 (two / one CTA per SM): +-1 %.
template <int N> constexpr int nd_step_barriers() {
#ifdef MFS_ND_STEP_BARRIER
  return MFS_ND_STEP_BARRIER;
#else
  return N <= 2 ? 0 : N == 3 ? 2 : N <= 5 ? 4 : 2;
#endif
}
#ifndef MFS_ND_FAST_QL_MAX_S
#define MFS_ND_FAST_QL_MAX_S 28   // largest S = N(N+1)/2 whose stage D tries the register QL first (28: every supported N)
#endif

struct NdArgs {
  int32_t mode;        // MFS_MODE_RAW / MFS_MODE_CENTRAL
  int32_t trans_id;    // MFS_TRANS_EULER / MFS_TRANS_TME_NORMAL / MFS_TRANS_TME
  int32_t tme_order;   // 1 or 2
  int32_t meas_id;     // MFS_MEAS_BERNOULLI_LOGISTIC_CUBIC on x[obs_dim]
  int32_t obs_dim;
  int32_t out_mode;
  int32_t stable;      // 1: LDL completion of the Gram factor (mfs/utils.py:526-538) instead of NaN on a non-positive pivot
  int64_t B, T;
  double dt;
  const double* trans_params;  // {alpha, beta, delta, gamma, sigma}
  int64_t trans_param_stride;
  const double* meas_params;
  int64_t meas_param_stride;
  const double* ms0;           // [B|1][z]
  int64_t ms0_stride;
  const double* mean0;         // [B|1][2]
  int64_t mean0_stride;
  const unsigned char* ys;     // [B][T] uint8
  const int32_t* inds;         // [3][s][s] gather tables (gram_and_hankel_indices_graded_lexico)
  const int32_t* pos;          // [2N][2N] position of multi-index (a, b) in the moment vector, -1 if |n| > 2N-1
  double* ms_out;              // FULL [B][T][z]; LAST [B][z]
  double* mean_out;            // FULL [B][T][2]; LAST [B][2]
  double* nell_out;
  int32_t* status_out;
};

MFS_DEV double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Lotka--Volterra drift x1 (alpha - beta x2), x2 (delta x1 - gamma), dispersion diag(sigma x): conditional mean and
// covariance.  Euler: mfs/multi_dims/moments.py:291-293.  TME order 2: tme.mean_and_cov restated (closed forms
// verified in the tests against a symbolic application of the generator, and SURVEY.md Appendix B).
MFS_DEV void lv_mean_cov(int trans_id, int order, double x1, double x2, double dt, const double* p, double& m1,
                         double& m2, double& c11, double& c12, double& c22) {
  const double al = p[0], be = p[1], de = p[2], ga = p[3], sg2 = p[4] * p[4];
  const double a1 = x1 * (al - be * x2), a2 = x2 * (de * x1 - ga);
  m1 = fma(dt, a1, x1);
  m2 = fma(dt, a2, x2);
  c11 = sg2 * x1 * x1 * dt;
  c22 = sg2 * x2 * x2 * dt;
  c12 = 0.0;
  if (trans_id == MFS_TRANS_TME_NORMAL && order >= 2) {
    const double h = 0.5 * dt * dt;
    // A^2 x_i = a . grad(a_i) + 1/2 tr(bb^T Hess a_i);  Hess a_1 = [[0,-beta],[-beta,0]], bb^T diagonal -> trace term 0
    m1 = fma(h, a1 * (al - be * x2) - be * x1 * a2, m1);
    m2 = fma(h, de * x2 * a1 + a2 * (de * x1 - ga), m2);
    c11 = fma(dt * dt * sg2 * x1 * x1, 2.0 * al - 2.0 * be * x2 + 0.5 * sg2, c11);
    c22 = fma(dt * dt * sg2 * x2 * x2, 2.0 * de * x1 - 2.0 * ga + 0.5 * sg2, c22);
    c12 = 0.5 * dt * dt * sg2 * x1 * x2 * (de * x1 - be * x2);
  }
}

// TME of the Lotka--Volterra SDE WITHOUT the Normal approximation (sde_cond_moments_tme, mfs/multi_dims/moments.py:414-479;
// third-party tme.expectation restated from its definition).  With A = a . grad + 1/2 Gamma : Hess,
// Gamma = diag(sg2 x1^2, sg2 x2^2), any smooth phi satisfies
//     sum_{r <= order} dt^r / r! A^r phi = sum_{p+q <= 2 order} G[p][q](x) d1^p d2^q phi .
// A = sum over the four "operator atoms" (i, j, c): (1,0,a1) (0,1,a2) (2,0,Gamma11/2) (0,2,Gamma22/2), and the product
// rule of a second-order operator, A(c D) = c A D + (A c) D + Gamma11 d1c d1 D + Gamma22 d2c d2 D, gives A^2.
// Orders 1 and 2.  tools/derive_lv_tme.py prints the same G symbolically; the CPU checker applies A to every monomial.
MFS_DEV void lv_tme_operator(int order, double x1, double x2, double dt, const double* p, double (&G)[5][5]) {
  const double al = p[0], be = p[1], de = p[2], ga = p[3], sg2 = p[4] * p[4];
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int j = 0; j < 5; ++j) G[i][j] = 0.0;
  const double g11 = sg2 * x1 * x1, g22 = sg2 * x2 * x2;
  const double c[4] = {x1 * (al - be * x2), x2 * (de * x1 - ga), 0.5 * g11, 0.5 * g22};
  G[0][0] = 1.0;
  G[1][0] = dt * c[0];
  G[0][1] = dt * c[1];
  G[2][0] = dt * c[2];
  G[0][2] = dt * c[3];
  if (order >= 2) {
    const double h = 0.5 * dt * dt;
    constexpr int oi[4] = {1, 0, 2, 0}, oj[4] = {0, 1, 0, 2};
    const double d1c[4] = {al - be * x2, de * x2, sg2 * x1, 0.0};
    const double d2c[4] = {-be * x1, de * x1 - ga, 0.0, sg2 * x2};
    const double d11c[4] = {0.0, 0.0, sg2, 0.0}, d22c[4] = {0.0, 0.0, 0.0, sg2};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const double Ac = c[0] * d1c[u] + c[1] * d2c[u] + c[2] * d11c[u] + c[3] * d22c[u];
#pragma unroll
      for (int v = 0; v < 4; ++v) G[oi[u] + oi[v]][oj[u] + oj[v]] = fma(h * c[u], c[v], G[oi[u] + oi[v]][oj[u] + oj[v]]);
      G[oi[u]][oj[u]] = fma(h, Ac, G[oi[u]][oj[u]]);
      G[oi[u] + 1][oj[u]] = fma(h * g11, d1c[u], G[oi[u] + 1][oj[u]]);
      G[oi[u]][oj[u] + 1] = fma(h * g22, d2c[u], G[oi[u]][oj[u] + 1]);
    }
  }
}

template <int N>
struct NdDims {
  static constexpr int S = N * (N + 1) / 2;        // basis size
  static constexpr int Z = N * (2 * N + 1);        // number of moments, |n| <= 2N-1
  static constexpr int M = 2 * N;                  // orders 0..2N-1 per dimension
  static constexpr int SS = S * S;
  // Per-warp shared memory in doubles: ms | R | T[kNumT] | V[2] | (wts) | lam[2] | Householder v[2], w[2].
  // Round 2: the weights live in the R region (the Cholesky factor is dead once stage B is done, the weights are dead
  // before the next call's stage A writes the factor), and the two-pass sizes (S > 16: one matrix at a time) share one
  // transpose buffer.  N = 7: 39.8 -> 27.3 KB per filter, which with the byte-wide gather table makes room for a second
  // CTA per SM (profiles/r2_ab_nd_smem.log); -DMFS_ND_NO_SMEM_ALIAS restores the round-1 layout.
#ifdef MFS_ND_NO_SMEM_ALIAS
  static constexpr bool kAlias = false;
#else
  static constexpr bool kAlias = true;
#endif
  static constexpr int kNumT = (S <= 16 || !kAlias) ? 2 : 1;
  static constexpr int kOffR = Z;
  static constexpr int kOffT = kOffR + SS;
  static constexpr int kOffV = kOffT + kNumT * SS;
  static constexpr int kOffW = kAlias ? kOffR : kOffV + 2 * SS;
  static constexpr int kOffLam = kOffV + 2 * SS + (kAlias ? 0 : SS);
  static constexpr int kOffHv = kOffLam + 2 * S;
  static constexpr int kOffHw = kOffHv + 2 * S;
  static constexpr int kDoubles = kOffHw + 2 * S;
  static constexpr int kScratch = (kNumT + 2) * SS;        // T and V: idle between quadratures (warp_reduce_moments)
  static constexpr int kTabBytes = (3 * SS + 7) / 8 * 8;   // per CTA: the (3, S, S) gather table, one byte per entry (< Z <= 105)
};

// ---------------------------------------------------------------------------------------------------------------------
// Quadrature: ms (shared) -> lam[2][S], V[2][S][S] (eigenvectors in columns), wts[S][S].  0 ok, 1 pivot, 2 no convergence.
//
// Row-per-lane formulation, matrices in REGISTERS with compile-time indices (everything is unrolled over S):
//   A  Cholesky of G: lane r owns row r; column j of the factor travels by shuffles.  The factor goes to shared
//      memory once (it is only ever read as a warp-wide broadcast afterwards).
//   B  K_k = R^-1 H_k R^-T: lane c forward-substitutes column c of H_k (Y = R^-1 H_k), the columns are transposed
//      through shared memory, and a second substitution gives column c of K_k = R^-1 Y^T (H_k is symmetric);
//      one more transpose symmetrises (jax.lax.linalg.eigh symmetrises its input).
//   C  Householder reduction to tridiagonal form; lane r owns row r of K_k and row r of the accumulated V.
//   D  implicit-shift QL with splitting on the tridiagonal (d, e) -- every lane of a group runs the same scalar
//      recurrence on a private copy and rotates its own row of V.  The sweep loop is uniform over the warp, the chase
//      is unrolled over the index range with the rotation predicated per group, and the set of negligible couplings
//      is a bitmask maintained as the couplings are produced: the two matrices advance in lockstep.
//   For S <= 16 the two matrices live in the two half-warps (stages B-D run once), otherwise they are processed one
//   after the other with all 32 lanes.
// ---------------------------------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------------------------------
// CTA-batched stage D (MFS_ND_BATCHED_QL, S in [10, 16]): the scalar QL recurrence of a matrix is the same in all 16
// lanes of its group, i.e. 17 of the 21 FP64 operations of a rotation are executed 16 times for nothing.  Here the
// CTA's warps (phase-aligned anyway) hand their 2 x nd_warps matrices to warp 0, whose lanes run ONE recurrence each
// (per-lane control flow as in the 1-D kernel: ql_run_recorded) and record the rotations of up to kSweeps sweeps; then
// every warp applies the recorded sweeps to its rows of V (1 broadcast load + 4 FP64 per rotation) -- until all
// matrices are done.  Buffers: the warp's own R and T regions, which are idle during stage D:
//   R region: per matrix k = 0, 1: d[S], e[S], {l, iter, dl, dl1, el}; then ints {state[2], nsw[2]}
//   T[k]:     double2 rec[kSweeps][S], then ints rec_l[kSweeps]
// A warp with nothing to solve (failed / missing filter, Cholesky failure, a split matrix that needs the general loop)
// takes part with take = false: the barriers of the protocol are executed by every warp of the CTA.
template <int N>
struct NdBatch {
  using D = NdDims<N>;
  static constexpr int S = D::S, SS = D::SS;
  static constexpr int kMat = 2 * S + 5;
  static constexpr int kSweeps = (SS - 4) / (2 * S) > 7 ? 7 : (SS - 4) / (2 * S);
  static constexpr bool kEnabled = S >= 10 && S <= 16;
  enum { SKIP = 0, RUN = 1, DONE = 2, FAILED = 3 };
  static_assert(!kEnabled || 2 * kMat + 2 <= SS, "R region too small");
  static_assert(!kEnabled || 1 + 2 * kSweeps * S + (kSweeps + 1) / 2 <= SS, "T region too small");
};

// The scalar phase of nd_batched_ql (warp 0 only): lane m runs up to kSweeps recorded sweeps of matrix m = 2 w + k.  One
// out-of-line copy: the protocol is instantiated at several call sites (live, idle, Cholesky failure).
template <int N>
__device__ __noinline__ void nd_batched_scalar_phase(double* __restrict__ cta_base, const int lane) {
  using NB = NdBatch<N>;
  using D = NdDims<N>;
  constexpr int S = D::S, SS = D::SS, W = nd_warps<N>(), K = NB::kSweeps, kMat = NB::kMat;
  constexpr unsigned kFull = 0xffffffffu;
  auto region_R = [&](int w) { return cta_base + w * D::kDoubles + D::kOffR; };
  auto region_T = [&](int w, int kk) {       // 16-byte aligned (double2 records): Z + SS may be odd
    double* t = cta_base + w * D::kDoubles + D::kOffT + kk * SS;
    return t + ((reinterpret_cast<uintptr_t>(t) >> 3) & 1);
  };
  auto flags_of = [&](int w) { return reinterpret_cast<int*>(region_R(w) + 2 * kMat); };
  int* const all_done = reinterpret_cast<int*>(cta_base + D::kOffHv);      // warp 0's hv[0]
  int running = 0;
  if (lane < 2 * W) {
    int* fl = flags_of(lane >> 1);
    const int kk = lane & 1;
    if (fl[kk] == NB::RUN) {
      double* q = region_R(lane >> 1) + kk * kMat;
      double d[S], e[S];
#pragma unroll
      for (int i = 0; i < S; ++i) { d[i] = q[i]; e[i] = q[S + i]; }
      int l = (int)q[2 * S], iter = (int)q[2 * S + 1];
      double dl = q[2 * S + 2], dl1 = q[2 * S + 3], el = q[2 * S + 4];
      double2* rec = reinterpret_cast<double2*>(region_T(lane >> 1, kk));
      int* rec_l = reinterpret_cast<int*>(region_T(lane >> 1, kk) + 2 * K * S);
      int nsw = 0;
      const int st = ql_run_recorded<S>(d, e, l, iter, dl, dl1, el, K, rec, rec_l, nsw);
#pragma unroll
      for (int i = 0; i < S; ++i) { q[i] = d[i]; q[S + i] = e[i]; }
      q[2 * S] = (double)l; q[2 * S + 1] = (double)iter; q[2 * S + 2] = dl; q[2 * S + 3] = dl1; q[2 * S + 4] = el;
      fl[2 + kk] = nsw;
      fl[kk] = st == 0 ? NB::RUN : st == 1 ? NB::DONE : NB::FAILED;
      running = st == 0;
    } else {
      fl[2 + kk] = 0;
    }
  }
  const int any = __any_sync(kFull, running);
  if (lane == 0) *all_done = !any;
}

// LIVE = false: participant without matrices.  Returns (per half-warp) whether its matrix was solved; then fd holds the
// eigenvalues and fz the lane's row of V Q.
template <int N, bool LIVE>
MFS_DEV bool nd_batched_ql(double* __restrict__ cta_base, const int warp, const int lane, const bool take,
                           const double (&qd)[NdDims<N>::S], const double (&qe)[NdDims<N>::S],
                           double (&fz)[NdDims<N>::S], double (&fd)[NdDims<N>::S]) {
  using NB = NdBatch<N>;
  using D = NdDims<N>;
  constexpr int S = D::S, SS = D::SS, W = nd_warps<N>(), K = NB::kSweeps, kMat = NB::kMat;
  static_assert(2 * W <= 32, "one lane per matrix");
  const int k = lane >> 4, r = lane & 15;
  auto region_R = [&](int w) { return cta_base + w * D::kDoubles + D::kOffR; };
  auto region_T = [&](int w, int kk) {       // 16-byte aligned (double2 records): Z + SS may be odd
    double* t = cta_base + w * D::kDoubles + D::kOffT + kk * SS;
    return t + ((reinterpret_cast<uintptr_t>(t) >> 3) & 1);
  };
  auto flags_of = [&](int w) { return reinterpret_cast<int*>(region_R(w) + 2 * kMat); };
  int* const all_done = reinterpret_cast<int*>(cta_base + D::kOffHv);      // warp 0's hv[0]
  int* const my_flags = flags_of(warp);
  if (LIVE) {
    if (r == 0) {
      double* q = region_R(warp) + k * kMat;
#pragma unroll
      for (int i = 0; i < S; ++i) { q[i] = qd[i]; q[S + i] = qe[i]; }
      q[2 * S] = -1.0;                                   // l < 0: not started
      my_flags[k] = take ? NB::RUN : NB::SKIP;
      my_flags[2 + k] = 0;
    }
  } else {
    if (lane < 2) { my_flags[lane] = NB::SKIP; my_flags[2 + lane] = 0; }
  }
  __syncthreads();
  for (;;) {
    if (warp == 0) nd_batched_scalar_phase<N>(cta_base, lane);
    __syncthreads();
    if (LIVE) {
      const int nsw = my_flags[2 + k];
      const double2* rec = reinterpret_cast<const double2*>(region_T(warp, k));
      const int* rec_l = reinterpret_cast<const int*>(region_T(warp, k) + 2 * K * S);
      for (int sw = 0; sw < nsw; ++sw) {
        const int l = rec_l[sw];
#pragma unroll
        for (int I = S - 2; I >= 0; --I) {
          if (I >= l) {
            const double2 cs = rec[sw * S + I];
            const double zi1 = fz[I + 1];
            fz[I + 1] = fma(cs.y, fz[I], cs.x * zi1);
            fz[I] = fma(cs.x, fz[I], -cs.y * zi1);
          }
        }
      }
    }
    const int done = *all_done;
    __syncthreads();
    if (done) break;
  }
  if (LIVE) {
    if (my_flags[k] == NB::DONE) {
      const double* q = region_R(warp) + k * kMat;
#pragma unroll
      for (int i = 0; i < S; ++i) fd[i] = q[i];
      return true;
    }
  }
  return false;
}

template <int N>
MFS_DEV void nd_batched_ql_idle(double* __restrict__ cta_base, const int warp, const int lane) {
  double u0[NdDims<N>::S], u1[NdDims<N>::S], u2[NdDims<N>::S], u3[NdDims<N>::S];      // never touched (LIVE = false)
  nd_batched_ql<N, false>(cta_base, warp, lane, false, u0, u1, u2, u3);
}

// BATCH: stage D through the CTA-batched protocol (nd_batched_ql): EVERY warp of the CTA must then make the same calls
// (a warp without a live filter calls nd_batched_ql_idle instead); cta_base = start of the CTA's shared memory.
template <int N, bool BATCH = false>
__device__ __noinline__ int quadrature_nd(double* sm, const unsigned char* __restrict__ tab, int lane, int stable,
                                          double* cta_base = nullptr, int warp = 0) {   // one copy, two call sites
  using D = NdDims<N>;
  constexpr int S = D::S, SS = D::SS;
  constexpr unsigned kFull = 0xffffffffu;
  double* ms = sm;
  double* R = sm + D::kOffR;      // [S][S] Cholesky factor
  double* T = sm + D::kOffT;      // [kNumT][S][S] transpose buffers
  double* V = sm + D::kOffV;      // [2][S][S]
  double* wts = sm + D::kOffW;    // [S][S] (the R region: see NdDims)
  double* lam = sm + D::kOffLam;  // [2][S]
  double* hv = sm + D::kOffHv;    // [2][S] Householder v
  double* hw = sm + D::kOffHw;    // [2][S] Householder w

  // ---- A: Cholesky (lower).  Failure <=> a pivot is not > 0 (jax: NaN-filled factor).
  double rdi[S];               // 1 / R[j][j], identical in every lane
  {
    const int rr = (lane < S) ? lane : 0;
    double g[S];
#pragma unroll
    for (int c = 0; c < S; ++c) g[c] = ms[tab[rr * S + c]];
    if (!stable) {
#pragma unroll
      for (int j = 0; j < S; ++j) {
        const double piv = __shfl_sync(kFull, g[j], j);
        if (!(piv > 0.0)) {
          if constexpr (BATCH) nd_batched_ql_idle<N>(cta_base, warp, lane);
          return 1;
        }
        const double rinv = rsqrt_fast(piv);
        rdi[j] = rinv;
        g[j] *= rinv;                                   // L[r][j] for r >= j
#pragma unroll
        for (int c = j + 1; c < S; ++c) {
          const double lcj = __shfl_sync(kFull, g[j], c);
          g[c] = fma(-g[j], lcj, g[c]);                 // only the entries c <= r are ever used
        }
      }
    } else {
      // `stable=True` (quadratures.py:152 with ldl=True): LDL^T without pivoting and the factor
      // L diag(where(d < 0, 1e-8 ||G||_F, sqrt(d)))  (mfs/utils.py:515-538); d == 0 gives inf/NaN like the reference.
      double fro = 0.0;
#pragma unroll
      for (int c = 0; c < S; ++c) fro = fma(g[c], g[c], fro);
      fro = warp_sum((lane < S) ? fro : 0.0);
      const double eps = 1e-8 * sqrt(fro);
#pragma unroll
      for (int j = 0; j < S; ++j) {
        const double piv = __shfl_sync(kFull, g[j], j);                  // d_j
        if (!(piv == piv)) {
          if constexpr (BATCH) nd_batched_ql_idle<N>(cta_base, warp, lane);
          return 1;
        }
        const double sj = (piv < 0.0) ? eps : sqrt(piv);
        rdi[j] = 1.0 / sj;
        const double gj = g[j];                                          // l_rj d_j
        const double lj = gj / piv;                                      // l_rj (1 on the diagonal)
#pragma unroll
        for (int c = j + 1; c < S; ++c) {
          const double gcj = __shfl_sync(kFull, gj, c);                  // l_cj d_j
          g[c] = fma(-lj, gcj, g[c]);
        }
        g[j] = lj * sj;                                                  // R[r][j]
      }
    }
    if (lane < S) {
#pragma unroll
      for (int c = 0; c < S; ++c) R[lane * S + c] = g[c];
    }
  }
  __syncwarp();

  constexpr int LPM = (S <= 16) ? 16 : 32;       // lanes per matrix
  constexpr int NPASS = (S <= 16) ? 1 : 2;
  constexpr unsigned kBits = (1u << S) - 1u;
  int fail = 0;
#pragma unroll 1
  for (int ps = 0; ps < NPASS; ++ps) {
    const int k = (LPM == 16) ? (lane >> 4) : ps;
    const int r = lane & (LPM - 1);
    const int gbase = lane & ~(LPM - 1);
    const bool act = r < S;
    const int cc = act ? r : 0;
    double* Tk = T + (D::kNumT == 2 ? k : 0) * SS;
    double* v = hv + k * S;
    double* w = hw + k * S;
    auto group_sum = [&](double x) {
#pragma unroll
      for (int o = LPM / 2; o > 0; o >>= 1) x += __shfl_xor_sync(kFull, x, o);
      return x;
    };
    auto solve = [&](double (&y)[S]) {               // y <- R^-1 y  (R broadcast from shared memory)
#pragma unroll
      for (int i = 0; i < S; ++i) {
        double acc = y[i];
#pragma unroll
        for (int q = 0; q < i; ++q) acc = fma(-R[i * S + q], y[q], acc);
        y[i] = acc * rdi[i];
      }
    };
    // ---- B
    double a[S];
    {
      double y[S];
#pragma unroll
      for (int i = 0; i < S; ++i) y[i] = ms[tab[(1 + k) * SS + i * S + cc]];      // column cc of H_k
      solve(y);
      if (act) {
#pragma unroll
        for (int i = 0; i < S; ++i) Tk[i * S + r] = y[i];
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < S; ++i) y[i] = Tk[cc * S + i];                          // row cc of Y
      __syncwarp();
      solve(y);                                                                   // column cc of K_k
      if (act) {
#pragma unroll
        for (int i = 0; i < S; ++i) Tk[i * S + r] = y[i];
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < S; ++i) a[i] = 0.5 * (y[i] + Tk[cc * S + i]);           // row cc of the symmetrised K_k
    }
    // ---- C
    double z[S];
    double qd[S], qe[S];
#pragma unroll
    for (int c = 0; c < S; ++c) z[c] = (c == r) ? 1.0 : 0.0;
#pragma unroll
    for (int j = 0; j + 2 < S; ++j) {
      const bool below = act && r > j;
      const double xr = below ? a[j] : 0.0;
      const double sigma = group_sum(xr * xr);
      const double x0 = __shfl_sync(kFull, a[j], gbase + j + 1);
      const double tail = sigma - x0 * x0;
      const bool skip = !(tail > 0.0) || !(sigma > 0.0);   // nothing below the sub-diagonal: H_j = I
      double alpha = x0, tau = 0.0;
      if (!skip) {
        alpha = -copysign(sqrt(sigma), x0);
        tau = rcp_fast(alpha * (alpha - x0));
      }
      const double vr = below ? ((r == j + 1) ? xr - alpha : xr) : 0.0;
      if (act) v[r] = vr;
      __syncwarp();
      double vc[S];
      double pr = 0.0, sq = 0.0;
#pragma unroll
      for (int c = j + 1; c < S; ++c) {
        vc[c] = v[c];
        pr = fma(a[c], vc[c], pr);
        sq = fma(z[c], vc[c], sq);
      }
      pr = below ? pr * tau : 0.0;
      sq *= tau;
      const double kappa = 0.5 * tau * group_sum(vr * pr);
      const double wr = fma(-kappa, vr, pr);
      if (act) w[r] = wr;
      __syncwarp();
#pragma unroll
      for (int c = j + 1; c < S; ++c) {
        a[c] -= fma(vr, w[c], wr * vc[c]);            // rows r <= j have vr = wr = 0
        z[c] = fma(-sq, vc[c], z[c]);                 // V <- V H_j
      }
      qd[j] = __shfl_sync(kFull, a[j], gbase + j);
      qe[j] = alpha;
    }
    if (S >= 2) {
      qd[S - 2] = __shfl_sync(kFull, a[S - 2], gbase + S - 2);
      qe[S - 2] = __shfl_sync(kFull, a[S - 2], gbase + S - 1);
    }
    qd[S - 1] = __shfl_sync(kFull, a[S - 1], gbase + S - 1);
    qe[S - 1] = 0.0;
    // ---- D
    // Large S: the V rows go to shared memory and the chase below is a ROLLED loop (run-time index) of ~100
    // instructions that stays in the instruction cache, instead of S-1 unrolled bodies (22 KB at S = 15) re-fetched
    // every sweep (measured: +8 % at N = 5, -4 % at N = 3, 4 -- profiles/r1_nd_occupancy.md).
    constexpr bool kRolled = S >= 15;
    double* Vr = V + k * SS + cc * S;
    unsigned neg = 1u << (S - 1);            // bit i <=> coupling e[i] negligible; bit S-1 is a sentinel
    double bound = 0.0;                      // Gershgorin bound of the spectrum: every |d| of the iteration stays below it
    auto scan_couplings = [&]() {
      neg = 1u << (S - 1);
      bound = 0.0;
#pragma unroll
      for (int i = 0; i + 1 < S; ++i) {
        if (fabs(qe[i]) <= kEps * (fabs(qd[i]) + fabs(qd[i + 1]))) neg |= 1u << i;
        bound = fmax(bound, fabs(qd[i]) + fabs(qe[i]) + (i > 0 ? fabs(qe[i - 1]) : 0.0));
      }
      bound = fmax(bound, fabs(qd[S - 1]) + (S > 1 ? fabs(qe[S - 2]) : 0.0));
    };
    scan_couplings();
    bool done = false, bad = false;
    if (kRolled && act) {
#pragma unroll
      for (int c = 0; c < S; ++c) Vr[c] = z[c];
    }
#ifndef MFS_ND_NO_FAST_QL
    // Round 2 fast path: when no coupling of the tridiagonal form is negligible at entry -- every step but the first
    // few, where the initial product-like measure makes K_1 / K_2 degenerate -- the eigen-solve is the 1-D kernel's
    // register chase (quadrature.cuh: ql_chase, compile-time indices, no split mask, no dynamic index, 26 instructions
    // per rotation instead of 76), run by every lane of the group on its own row z of V.  It never splits the interior:
    // a chase through a small coupling is still an exact orthogonal similarity.  If it does not converge within its
    // sweep budget (or meets an exactly decoupled block) the general loop below starts over from (qd, qe, V).
    if constexpr (BATCH) {
      // every warp of the CTA is here (or in nd_batched_ql_idle) at the same time
      const bool take = neg == (1u << (S - 1));
      double fd[S], fz[S];
#pragma unroll
      for (int i = 0; i < S; ++i) fz[i] = z[i];
      bool ok = nd_batched_ql<N, true>(cta_base, warp, lane, take, qd, qe, fz, fd);
      double chk = 0.0;
#pragma unroll
      for (int i = 0; i < S; ++i) chk += fabs(ok ? fd[i] : 0.0);
      ok = ok && (chk <= 1.79e308);
      if (ok) {
#pragma unroll
        for (int i = 0; i < S; ++i) qd[i] = fd[i];
        if constexpr (kRolled) {
          if (act) {
#pragma unroll
            for (int c = 0; c < S; ++c) Vr[c] = fz[c];
          }
        } else {
#pragma unroll
          for (int i = 0; i < S; ++i) z[i] = fz[i];
        }
        done = true;
      }
    } else if constexpr (S <= MFS_ND_FAST_QL_MAX_S) {
      if (neg == (1u << (S - 1))) {
        // on copies: a rotation whose f and g both vanish (an exactly decoupled block, which the general loop handles by
        // splitting) would turn the state into NaN; the result is committed only if it converged with finite eigenvalues
        double fd[S], fe[S], fz[S];
#pragma unroll
        for (int i = 0; i < S; ++i) { fd[i] = qd[i]; fe[i] = qe[i]; fz[i] = z[i]; }
        bool ok = tridiag_ql_first_row<S, false, (S <= 16 ? MFS_ND_CASCADE_SMALL : MFS_ND_CASCADE_LARGE), (S >= MFS_ND_QL_CHAIN7_MIN_S)>(fd, fe, fz);
        double chk = 0.0;
#pragma unroll
        for (int i = 0; i < S; ++i) chk += fabs(fd[i]);          // identical in every lane of the group
        ok = ok && (chk <= 1.79e308);
        if (ok) {
#pragma unroll
          for (int i = 0; i < S; ++i) qd[i] = fd[i];
          if constexpr (kRolled) {
            if (act) {
#pragma unroll
              for (int c = 0; c < S; ++c) Vr[c] = fz[c];
            }
          } else {
#pragma unroll
            for (int i = 0; i < S; ++i) z[i] = fz[i];
          }
          done = true;
        }
      }
    }
#endif
    const double tiny = 2.0 * kEps * bound;
    int l = 0, iter = 0;
    for (;;) {
      int m = 0;
      if (!done) {
        const unsigned active = ~neg & kBits & (~0u << l);
        if (active == 0u) {
          done = true;
        } else {
          const int l_new = __ffs(active) - 1;      // first unreduced block starts here ...
          if (l_new != l) { l = l_new; iter = 0; }
          m = __ffs(neg & (~0u << l)) - 1;          // ... and ends at m > l
          if (++iter > 40) { bad = true; done = true; }
        }
      }
      if (__all_sync(kFull, done)) break;
      int hi = done ? 0 : m, lo = done ? S : l;
      if (LPM == 16) {
        hi = max(hi, __shfl_xor_sync(kFull, hi, 16));
        lo = min(lo, __shfl_xor_sync(kFull, lo, 16));
      }
      double g = 0.0, sn = 1.0, cs = 1.0, pp = 0.0;
      if (!done) {   // Wilkinson shift (dynamic reads: qd / qe live in local memory)
        const double el = qe[l], dl = qd[l];
        const double ah = (qd[l + 1] - dl) * rcp_fast(2.0 * el);
        g = qd[m] - dl + el * rcp_fast(ah + copysign(sqrt_fast(fma(ah, ah, 1.0)), ah));
      }
      bool stop = done;
      double d_below = 0.0;                  // d[i+2] as left by the previous rotation
      auto rotate = [&](const int i, const double e_i, const double d_i, const double d_up) {
        const double f = sn * e_i, b = cs * e_i;
        const double h2 = fma(f, f, g * g);
        if (h2 == 0.0) {                     // underflow: the matrix splits here
          qd[i + 1] = d_up - pp;
          qe[i + 1] = 0.0;
          neg |= 1u << (i + 1);
          stop = true;
        } else {
          const double rinv = rsqrt_fast(h2);
          const double e_new = h2 * rinv;
          qe[i + 1] = e_new;
          sn = f * rinv;
          cs = g * rinv;
          g = d_up - pp;
          const double rr = fma(d_i - g, sn, (cs + cs) * b);
          pp = sn * rr;
          const double d_new = g + pp;
          qd[i + 1] = d_new;
          g = fma(cs, rr, -b);
          if constexpr (kRolled) {
            if (act) {
              const double z1 = Vr[i + 1], z0 = Vr[i];
              Vr[i + 1] = fma(sn, z0, cs * z1);
              Vr[i] = fma(cs, z0, -sn * z1);
            }
          } else {
            const double z1 = z[i + 1];
            z[i + 1] = fma(sn, z[i], cs * z1);
            z[i] = fma(cs, z[i], -sn * z1);
          }
          // e[i+1] and d[i+1], d[i+2] are final for this sweep.  Inside a block every bit is clear, so bits are only ever
          // SET; `tiny` (2 eps x a bound of the spectrum) is a necessary condition that keeps the exact test off the
          // common path.
          if (e_new <= tiny) {
            if (e_new <= kEps * (fabs(d_new) + fabs(d_below))) neg |= 1u << (i + 1);
          }
          d_below = d_new;
        }
      };
      if constexpr (kRolled) {
        // qd / qe live in local memory (run-time indices): rotation i reads (e[i], d[i]) one iteration ahead -- it does
        // not modify them -- and takes d[i+1] from the previous iteration's d[i], so no load sits on the recurrence.
        // (measured, profiles/r1_nd_prefetch_ab.log: +33 % at N = 6 (S = 21), -4 % at N = 5 (S = 15) -> S >= 21 only)
        if constexpr (S >= 21) {
          double e_nx = 0.0, d_nx = 0.0, d_carry = 0.0;
          if (hi - 1 >= lo) { e_nx = qe[hi - 1]; d_nx = qd[hi - 1]; d_carry = qd[hi]; }
#pragma unroll 1
          for (int i = hi - 1; i >= lo; --i) {
            const double e_i = e_nx, d_i = d_nx, d_up = d_carry;
            if (i > lo) { e_nx = qe[i - 1]; d_nx = qd[i - 1]; }
            d_carry = d_i;
            if (!stop && i < m && i >= l) rotate(i, e_i, d_i, d_up);
          }
        } else {
#pragma unroll 1
          for (int i = hi - 1; i >= lo; --i)
            if (!stop && i < m && i >= l) rotate(i, qe[i], qd[i], qd[i + 1]);
        }
      } else {
#pragma unroll
        for (int i = S - 2; i >= 0; --i) {
          if (i < lo) break;
          if (i < hi && !stop && i < m && i >= l) rotate(i, qe[i], qd[i], qd[i + 1]);
        }
      }
      if (!done) {
        qe[m] = 0.0;
        neg |= 1u << m;
        if (!stop) {
          const double dl = qd[l] - pp;
          qd[l] = dl;
          qe[l] = g;
          if (fabs(g) <= tiny && fabs(g) <= kEps * (fabs(dl) + fabs(qd[l + 1]))) neg |= 1u << l;
        }
      }
    }
    if (bad) fail = 1;
    if (!kRolled && act) {
#pragma unroll
      for (int c = 0; c < S; ++c) V[k * SS + r * S + c] = z[c];
    }
    if (r == 0) {
#pragma unroll
      for (int i = 0; i < S; ++i) lam[k * S + i] = qd[i];
    }
    __syncwarp();
  }
  if (__any_sync(kFull, fail)) return 2;
  // weights: <v1_i, v2_j> v1_i[0] v2_j[0]   (quadratures.py:169-170)
  for (int e = lane; e < SS; e += 32) {
    const int i = e / S, j = e % S;
    double dot = 0.0;
    for (int r = 0; r < S; ++r) dot = fma(V[r * S + i], V[SS + r * S + j], dot);
    wts[e] = dot * V[i] * V[SS + j];
  }
  __syncwarp();
  return 0;
}

// accumulate acc[pos(a, b)] += wgt * M[a][b] for the Gaussian product moments of N((mu1, mu2), C) -- registers only
template <int N>
MFS_DEV void accumulate_gaussian_moments(double wgt, double mu1, double mu2, double c11, double c12, double c22,
                                         double (&acc)[NdDims<N>::Z]) {
  constexpr int M = 2 * N;
  // rows indexed by b (power of x2); within a row a = 0..M-1-b.  M[a+1][b] = mu1 M[a][b] + c11 a M[a-1][b] + c12 b M[a][b-1]
  double prev[M], cur[M], prev2[M];
#pragma unroll
  for (int a = 0; a < M; ++a) { prev[a] = 0.0; prev2[a] = 0.0; cur[a] = 0.0; }
#pragma unroll
  for (int b = 0; b < M; ++b) {
    // first entry of the row: M[0][b] = mu2 M[0][b-1] + c22 (b-1) M[0][b-2]
    if (b == 0) cur[0] = 1.0;
    else cur[0] = fma(mu2, prev[0], (b >= 2) ? c22 * (double)(b - 1) * prev2[0] : 0.0);
#pragma unroll
    for (int a = 0; a + 1 < M - b; ++a) {
      double v = mu1 * cur[a];
      if (a >= 1) v = fma(c11 * (double)a, cur[a - 1], v);
      if (b >= 1) v = fma(c12 * (double)b, prev[a], v);
      cur[a + 1] = v;
    }
#pragma unroll
    for (int a = 0; a < M - b; ++a) {
      constexpr int dummy = 0;
      (void)dummy;
      acc[(a + b) * (a + b + 1) / 2 + a] = fma(wgt, cur[a], acc[(a + b) * (a + b + 1) / 2 + a]);
    }
#pragma unroll
    for (int a = 0; a < M; ++a) { prev2[a] = prev[a]; prev[a] = cur[a]; }
  }
}

// acc[pos(a, b)] += wgt * sum_{p,q} G[p][q] d1^p d2^q [dl1^a dl2^b]
//                  = wgt * a! b! * sum_q (sum_p G[p][q] dl1^{a-p}/(a-p)!) dl2^{b-q}/(b-q)!      -- registers only
template <int N>
MFS_DEV void accumulate_tme_moments(double wgt, double dl1, double dl2, const double (&G)[5][5],
                                    double (&acc)[NdDims<N>::Z]) {
  constexpr int M = 2 * N;
  double P1[M], P2[M];
  P1[0] = 1.0;
  P2[0] = 1.0;
#pragma unroll
  for (int a = 1; a < M; ++a) { P1[a] = P1[a - 1] * dl1 * (1.0 / a); P2[a] = P2[a - 1] * dl2 * (1.0 / a); }
  double fa = wgt;   // wgt * a!
#pragma unroll
  for (int a = 0; a < M; ++a) {
    if (a > 0) fa *= (double)a;
    double u[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      double v = 0.0;
#pragma unroll
      for (int pp = 0; pp + q < 5; ++pp)
        if (pp <= a) v = fma(G[pp][q], P1[a - pp], v);
      u[q] = v * fa;
    }
    double fb = 1.0;   // b!
#pragma unroll
    for (int b = 0; b < M - a; ++b) {
      if (b > 0) fb *= (double)b;
      double t = 0.0;
#pragma unroll
      for (int q = 0; q < 5; ++q)
        if (q <= b) t = fma(u[q], P2[b - q], t);
      acc[(a + b) * (a + b + 1) / 2 + a] = fma(t, fb, acc[(a + b) * (a + b + 1) / 2 + a]);
    }
  }
}

// ms[p] = scale * (sum over the warp's lanes of acc[p]) for all Z moments, through a shared-memory scratch of 4 S^2
// doubles (the transpose / eigenvector buffers, idle outside the quadrature): each lane drops its partial sums in a
// skewed [moment][lane] tile (conflict-free in both directions) and lane q adds up moment q.  One rolled summation loop
// instead of Z unrolled butterfly reductions (the latter were 60 % of the kernel's code size).
template <int N>
MFS_DEV void warp_reduce_moments(const double (&acc)[NdDims<N>::Z], double* scratch, double* ms, double scale, int lane) {
  constexpr int Z = NdDims<N>::Z;
  constexpr int kScr = NdDims<N>::kScratch;
  constexpr int BS = (kScr / 32 > 32) ? 32 : (kScr / 32 < 1 ? 1 : kScr / 32);   // moments per batch
#pragma unroll
  for (int p0 = 0; p0 < Z; p0 += BS) {
#pragma unroll
    for (int q = 0; q < BS; ++q)
      if (p0 + q < Z) scratch[q * 32 + ((lane + q) & 31)] = acc[p0 + q];
    __syncwarp();
    if (lane < BS && p0 + lane < Z) {
      double sum = 0.0;
#pragma unroll 4
      for (int j = 0; j < 32; ++j) sum += scratch[lane * 32 + ((j + lane) & 31)];
      ms[p0 + lane] = sum * scale;
    }
    __syncwarp();
  }
}

// CTAs per SM the register allocation must allow (measured on B200, profiles/r1_nd_occupancy.md): the eigen-solver is a
// latency-bound scalar recurrence, so resident warps pay until the moment accumulators start to spill.
#ifdef MFS_ND_MIN_BLOCKS
template <int N> constexpr int nd_min_blocks() { return MFS_ND_MIN_BLOCKS; }
#else
template <int N> constexpr int nd_min_blocks() { return N <= 4 ? 4 : N == 5 ? 2 : 8 / nd_warps<N>(); }
#endif
template <int N>
__global__ void __launch_bounds__(nd_warps<N>() * 32, nd_min_blocks<N>()) filter_nd_kernel(const NdArgs P) {
  using D = NdDims<N>;
  constexpr int kNdWarps = nd_warps<N>();
  constexpr int S = D::S, SS = D::SS, Z = D::Z, M = D::M;
  extern __shared__ double smem_all[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * kNdWarps + warp;
  // the (3, S, S) gather table, shared by the CTA's warps
  unsigned char* tab = reinterpret_cast<unsigned char*>(smem_all + kNdWarps * D::kDoubles);
  for (int e = threadIdx.x; e < 3 * SS; e += kNdWarps * 32) tab[e] = (unsigned char)__ldg(P.inds + e);
  __syncthreads();
  // Phase barriers (nd_step_barriers<N>() per time step): the CTA's warps are independent filters, but left alone they
  // drift apart (data-dependent sweep counts) and each streams the kernel's 250+ KB of unrolled code through the
  // instruction caches on its own; re-aligning them a few times per step lets one fetch serve all warps of the CTA
  // (profiles/r2_ab_nd_barrier.log: +10..20 %).  A step is the fixed sequence  B1 Q B3 B2 Q B4  (B = barrier if enabled,
  // Q = a quadrature, which in the CTA-batched build contains barriers of its own); a warp without a filter, or whose
  // filter failed, runs the rest of that sequence as a participant without work (zombie_rest) before it leaves.
  constexpr int kBarriers = nd_step_barriers<N>();
#ifdef MFS_ND_BATCHED_QL
  constexpr bool kBatch = NdBatch<N>::kEnabled;
#else
  constexpr bool kBatch = false;
#endif
#define MFS_ND_PHASE_BARRIER(level) \
  if constexpr (kBarriers >= (level)) __syncthreads();
  auto zombie_rest = [&](const int64_t t0, int stage) {   // 0: whole steps from t0; 1 / 2: step t0 after its first / second quadrature
    for (int64_t tt = t0; tt < P.T; ++tt) {
      if (stage == 0) {
        MFS_ND_PHASE_BARRIER(1)
        if constexpr (kBatch) nd_batched_ql_idle<N>(smem_all, warp, lane);
      }
      if (stage <= 1) {
        MFS_ND_PHASE_BARRIER(3)
        MFS_ND_PHASE_BARRIER(2)
        if constexpr (kBatch) nd_batched_ql_idle<N>(smem_all, warp, lane);
      }
      MFS_ND_PHASE_BARRIER(4)
      stage = 0;
    }
  };
  if (b >= P.B) {
    zombie_rest(0, 0);
    return;
  }
  double* sm = smem_all + warp * D::kDoubles;
  double* ms = sm;
  double* scratch = sm + D::kOffT;    // T and V of the quadrature (NdDims::kScratch doubles), idle between quadratures
  double* wts = sm + D::kOffW;
  double* lam = sm + D::kOffLam;

  for (int e = lane; e < Z; e += 32) ms[e] = __ldg(P.ms0 + b * P.ms0_stride + e);
  double mean1 = 0.0, mean2 = 0.0;
  if (P.mode == MFS_MODE_CENTRAL) {
    mean1 = __ldg(P.mean0 + b * P.mean0_stride);
    mean2 = __ldg(P.mean0 + b * P.mean0_stride + 1);
  }
  double tp[5], mp[3];
#pragma unroll
  for (int k = 0; k < 5; ++k) tp[k] = __ldg(P.trans_params + b * P.trans_param_stride + k);
  mp[0] = __ldg(P.meas_params + b * P.meas_param_stride);
  mp[1] = __ldg(P.meas_params + b * P.meas_param_stride + 1);
  mp[2] = 1.0 / mp[0];
  __syncwarp();

  // The graded-lex position of (a, b) is (a+b)(a+b+1)/2 + a for d = 2; the host verifies that the tables it was given
  // (multi_indices, inds) are in that order, so the accumulators below can use compile-time positions.
  double nell = 0.0;
  int status = -1, reason = 0, fail_stage = 0;
  int64_t t = 0;
  for (; t < P.T; ++t) {
    MFS_ND_PHASE_BARRIER(1)
    const double y = (double)__ldg(P.ys + b * P.T + t);
    // ---------------- prediction ----------------
    int why = quadrature_nd<N, kBatch>(sm, tab, lane, P.stable, smem_all, warp);
    if (why) { status = (int)t; reason = why; fail_stage = 1; break; }
    MFS_ND_PHASE_BARRIER(3)
    double acc[Z];
    const bool tme_full = P.trans_id == MFS_TRANS_TME;   // TME without the Normal approximation
    const bool central = P.mode == MFS_MODE_CENTRAL;
    double nm1 = 0.0, nm2 = 0.0;       // centre of the predicted moments (0 in raw mode, where mean1 = mean2 = 0 too)
    if (central) {
      double s1 = 0.0, s2 = 0.0;
      for (int e = lane; e < SS; e += 32) {
        const double x1 = lam[e / S] + mean1, x2 = lam[S + e % S] + mean2;
        double m1, m2, c11, c12, c22;
        // state_cond_mean = tme.expectation(identity): same expansion for both TME factories (moments.py:401-403, 469-471)
        lv_mean_cov(tme_full ? MFS_TRANS_TME_NORMAL : P.trans_id, P.tme_order, x1, x2, P.dt, tp, m1, m2, c11, c12, c22);
        s1 = fma(wts[e], m1, s1);
        s2 = fma(wts[e], m2, s2);
      }
      nm1 = warp_sum(s1);
      nm2 = warp_sum(s2);
    }
#pragma unroll
    for (int p = 0; p < Z; ++p) acc[p] = 0.0;
    for (int e = lane; e < SS; e += 32) {      // nodes still refer to the old mean
      const double x1 = lam[e / S] + mean1, x2 = lam[S + e % S] + mean2;
      if (tme_full) {
        double G[5][5];
        lv_tme_operator(P.tme_order, x1, x2, P.dt, tp, G);
        accumulate_tme_moments<N>(wts[e], x1 - nm1, x2 - nm2, G, acc);
      } else {
        double m1, m2, c11, c12, c22;
        lv_mean_cov(P.trans_id, P.tme_order, x1, x2, P.dt, tp, m1, m2, c11, c12, c22);
        accumulate_gaussian_moments<N>(wts[e], m1 - nm1, m2 - nm2, c11, c12, c22, acc);
      }
    }
    if (central) { mean1 = nm1; mean2 = nm2; }
    __syncwarp();
    warp_reduce_moments<N>(acc, scratch, ms, 1.0, lane);
    // ---------------- update ----------------
    MFS_ND_PHASE_BARRIER(2)
    why = quadrature_nd<N, kBatch>(sm, tab, lane, P.stable, smem_all, warp);
    if (why) { status = (int)t; reason = why + 4; fail_stage = 2; break; }
    MFS_ND_PHASE_BARRIER(4)
    MeasStep st;
    st.y = y;
    st.c0 = 0.0;
    double cc = 0.0, u1 = 0.0, u2 = 0.0;
    for (int e = lane; e < SS; e += 32) {
      const double x1 = lam[e / S] + mean1, x2 = lam[S + e % S] + mean2;
      const double lik = measurement_pdf(P.meas_id, st, P.obs_dim == 0 ? x1 : x2, mp);
      const double u = wts[e] * lik;
      wts[e] = u;                 // keep w * likelihood for the moment pass
      cc += u;
      u1 = fma(u, x1, u1);
      u2 = fma(u, x2, u2);
    }
    cc = warp_sum(cc);
    u1 = warp_sum(u1);
    u2 = warp_sum(u2);
    const double cinv = 1.0 / cc;
    double c1 = 0.0, c2 = 0.0;     // centre of the posterior moments
    if (P.mode == MFS_MODE_CENTRAL) { c1 = u1 * cinv; c2 = u2 * cinv; }
    __syncwarp();
#pragma unroll
    for (int p = 0; p < Z; ++p) acc[p] = 0.0;
    for (int e = lane; e < SS; e += 32) {
      const double d1 = lam[e / S] + mean1 - c1, d2 = lam[S + e % S] + mean2 - c2;
      double pa[M];
      pa[0] = 1.0;
#pragma unroll
      for (int a = 1; a < M; ++a) pa[a] = pa[a - 1] * d1;
      double pb = wts[e];
#pragma unroll
      for (int bb = 0; bb < M; ++bb) {
#pragma unroll
        for (int a = 0; a < M - bb; ++a) acc[(a + bb) * (a + bb + 1) / 2 + a] = fma(pb, pa[a], acc[(a + bb) * (a + bb + 1) / 2 + a]);
        pb *= d2;
      }
    }
    __syncwarp();
    warp_reduce_moments<N>(acc, scratch, ms, cinv, lane);
    if (P.mode == MFS_MODE_CENTRAL) { mean1 = c1; mean2 = c2; }
    nell -= log(cc);
    __syncwarp();
    if (P.out_mode == MFS_OUT_FULL) {
      double* o = P.ms_out + (b * P.T + t) * Z;
      for (int e = lane; e < Z; e += 32) o[e] = ms[e];
      if (P.mode == MFS_MODE_CENTRAL && lane == 0) {
        P.mean_out[(b * P.T + t) * 2] = mean1;
        P.mean_out[(b * P.T + t) * 2 + 1] = mean2;
      }
    } else if (P.out_mode == MFS_OUT_MEANVAR) {
      // (E x1, E x2, Var x1, Cov(x1, x2), Var x2): graded-lex positions (0,1) = 1, (1,0) = 2, (0,2) = 3, (1,1) = 4, (2,0) = 5
      if (lane == 0) {
        double* o = P.ms_out + (b * P.T + t) * 5;
        if (P.mode == MFS_MODE_CENTRAL) {
          o[0] = mean1; o[1] = mean2; o[2] = ms[5]; o[3] = ms[4]; o[4] = ms[3];
        } else {
          const double e1 = ms[2], e2 = ms[1];
          o[0] = e1; o[1] = e2; o[2] = fma(-e1, e1, ms[5]); o[3] = fma(-e1, e2, ms[4]); o[4] = fma(-e2, e2, ms[3]);
        }
      }
    }
  }
  if (status >= 0) zombie_rest(t, fail_stage);     // a failed filter: the rest of the CTA's barrier sequence
#undef MFS_ND_PHASE_BARRIER
  if (status >= 0) {
    const double qnan = nan("");
    nell = qnan; mean1 = qnan; mean2 = qnan;
    __syncwarp();
    for (int e = lane; e < Z; e += 32) ms[e] = qnan;
    __syncwarp();
    if (P.out_mode == MFS_OUT_MEANVAR) {
      for (int64_t e = t * 5 + lane; e < P.T * 5; e += 32) P.ms_out[b * P.T * 5 + e] = qnan;
    }
    if (P.out_mode == MFS_OUT_FULL) {
      for (; t < P.T; ++t) {
        double* o = P.ms_out + (b * P.T + t) * Z;
        for (int e = lane; e < Z; e += 32) o[e] = qnan;
        if (P.mode == MFS_MODE_CENTRAL && lane == 0) {
          P.mean_out[(b * P.T + t) * 2] = qnan;
          P.mean_out[(b * P.T + t) * 2 + 1] = qnan;
        }
      }
    }
  }
  if (P.out_mode == MFS_OUT_LAST) {
    for (int e = lane; e < Z; e += 32) P.ms_out[b * Z + e] = ms[e];
    if (P.mode == MFS_MODE_CENTRAL && lane == 0) { P.mean_out[b * 2] = mean1; P.mean_out[b * 2 + 1] = mean2; }
  }
  if (lane == 0) {
    P.nell_out[b] = nell;
#ifdef MFS_ND_DEBUG_REASON
    if (status >= 0) status |= reason << 24;
#endif
    (void)reason;
    if (P.status_out) P.status_out[b] = status;
  }
}

// Batched moment_quadrature_nd (mfs/multi_dims/quadratures.py:120-178) as an API of its own: one warp per moment vector,
// the same quadrature_nd<N> as inside the filter.  weights [B][S^2], nodes [B][S^2][2] in the reference's Cartesian
// order (node e = i S + j is (lambda1_i, lambda2_j); the order / signs within each eigen-decomposition are arbitrary).
struct NdQuadArgs {
  int64_t B;
  const double* ms;     // [B][z]
  const double* mean;   // [B][2] or nullptr
  const double* scale;  // [B][2] or nullptr
  const int32_t* inds;  // [3][S][S]
  int32_t stable;
  double* weights;
  double* nodes;
};

template <int N>
__global__ void __launch_bounds__(nd_warps<N>() * 32, nd_min_blocks<N>()) quadrature_nd_kernel(const NdQuadArgs P) {
  using D = NdDims<N>;
  constexpr int kNdWarps = nd_warps<N>();
  constexpr int S = D::S, SS = D::SS, Z = D::Z;
  extern __shared__ double smem_all[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * kNdWarps + warp;
  unsigned char* tab = reinterpret_cast<unsigned char*>(smem_all + kNdWarps * D::kDoubles);
  for (int e = threadIdx.x; e < 3 * SS; e += kNdWarps * 32) tab[e] = (unsigned char)__ldg(P.inds + e);
  __syncthreads();
  if (b >= P.B) return;
  double* sm = smem_all + warp * D::kDoubles;
  double* wts = sm + D::kOffW;
  double* lam = sm + D::kOffLam;
  for (int e = lane; e < Z; e += 32) sm[e] = __ldg(P.ms + b * Z + e);
  __syncwarp();
  const int why = quadrature_nd<N>(sm, tab, lane, P.stable);
  const double m1 = P.mean ? P.mean[b * 2] : 0.0, m2 = P.mean ? P.mean[b * 2 + 1] : 0.0;
  const double s1 = P.scale ? P.scale[b * 2] : 1.0, s2 = P.scale ? P.scale[b * 2 + 1] : 1.0;
  const double qnan = nan("");
  for (int e = lane; e < SS; e += 32) {
    P.weights[b * SS + e] = why ? qnan : wts[e];
    P.nodes[(b * SS + e) * 2] = why ? qnan : fma(s1, lam[e / S], m1);
    P.nodes[(b * SS + e) * 2 + 1] = why ? qnan : fma(s2, lam[S + e % S], m2);
  }
}

template <int N>
cudaError_t launch_quadrature_nd(const NdQuadArgs& a, cudaStream_t stream);

template <int N>
cudaError_t launch_filter_nd(const NdArgs& a, cudaStream_t stream);

}  // namespace mfs
