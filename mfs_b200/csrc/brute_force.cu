// Brute-force grid filter for 1-D state (mfs/classical_filters_smoothers/brute_force.py:26-136), batched over B
// independent measurement records that share one spatial grid.
//
// 'chapman-*' prediction (brute_force.py:80-84, 115-122): one integration sub-step is
//     ps'[i] = trapz_j( N(x_i; m_j, s_j) ps[j], xs ) = sum_j Pw[i][j] ps[j],   Pw[i][j] = N(x_i; m_j, s_j) w_j,
// with w the trapezoid weights of the grid.  For a batch this is the dense contraction  S' = S Pw^T  with
// S [B][n] (one filter's density per row): an FP64 GEMM with both operands K-contiguous, run on the FP64 tensor pipe
// (DMMA, mma.sync.m8n8k4.f64 -- tcgen05 has no FP64 kind) from a cp.async multi-stage shared-memory pipeline.
// Pw (n^2 doubles, 32 MB at the paper's n = 2000) stays L2-resident; S ping-pongs between two workspace buffers.
// The Bayes update (brute_force.py:135) is a row-wise kernel: u = p(y|x) ps, c = trapz(u), ps = u / c.
// 'kolmogorov' (brute_force.py:96-113): explicit Euler on the forward operator with jnp.gradient stencils, one CTA per
// filter, density resident in shared memory for all sub-steps of a time step.
#include <cuda_runtime.h>

#include "common.h"
#include "models.cuh"

namespace mfs {

// ------------------------------------------------------------------------------------------------------------------
// GEMM tile configuration
// ------------------------------------------------------------------------------------------------------------------
constexpr int BF_BM = 128;       // filters per CTA tile
constexpr int BF_BN = 128;       // output grid points per CTA tile
constexpr int BF_BK = 16;        // contraction slice per pipeline stage
constexpr int BF_LDS = BF_BK + 4;  // smem row stride (doubles): 8 rows x 4 k hit 32 distinct banks per half-warp
constexpr int BF_STAGES = 4;
constexpr int BF_THREADS = 256;  // 8 warps as 2 (M) x 4 (N); warp tile 64 x 32 = 8 x 4 DMMA tiles
constexpr int BF_WM = 64, BF_WN = 32;
constexpr size_t BF_SMEM = sizeof(double) * BF_STAGES * (BF_BM + BF_BN) * BF_LDS;

static inline int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

MFS_DEV void dmma_884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

MFS_DEV void cp_async_16(void* smem, const void* gmem, bool pred) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int bytes = pred ? 16 : 0;   // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(bytes));
}
MFS_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
MFS_DEV void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// D[b][i] = sum_j A[b][j] Pw[i][j]   for b < M, i < n;  A, D: [M][ld] (ld = Kpad, pad columns zero), Pw: [Npad][Kpad].
__global__ void __launch_bounds__(BF_THREADS, 1)
bf_gemm_kernel(const double* __restrict__ A, const double* __restrict__ Pw, double* __restrict__ D, int64_t M, int n,
               int Kpad) {
  extern __shared__ __align__(16) double bf_smem[];
  double* As = bf_smem;                                  // [STAGES][BM][LDS]
  double* Bs = bf_smem + BF_STAGES * BF_BM * BF_LDS;     // [STAGES][BN][LDS]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp >> 2, wn = warp & 3;
  const int64_t m0 = (int64_t)blockIdx.y * BF_BM;
  const int n0 = blockIdx.x * BF_BN;
  const int KT = Kpad / BF_BK;

  // each thread moves 4 + 4 16-byte chunks per stage: chunk c -> row c / 8, doubles (c % 8) * 2 .. +1
  auto load_stage = [&](int stage, int kt) {
    const int k0 = kt * BF_BK;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int c = tid + r * BF_THREADS;
      const int row = c >> 3, col = (c & 7) * 2;
      const int64_t gr = m0 + row;
      const bool ok = gr < M;
      cp_async_16(As + (stage * BF_BM + row) * BF_LDS + col, A + (ok ? gr : 0) * Kpad + k0 + col, ok);
      cp_async_16(Bs + (stage * BF_BN + row) * BF_LDS + col, Pw + (int64_t)(n0 + row) * Kpad + k0 + col, true);
    }
  };

  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
  for (int s = 0; s < BF_STAGES - 1; ++s) {
    if (s < KT) load_stage(s, s);
    cp_async_commit();
  }
  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<BF_STAGES - 2>();
    __syncthreads();
    {
      const int nk = kt + BF_STAGES - 1;
      if (nk < KT) load_stage(nk % BF_STAGES, nk);
      cp_async_commit();
    }
    const double* as = As + ((kt % BF_STAGES) * BF_BM + wm * BF_WM + g) * BF_LDS + t;
    const double* bs = Bs + ((kt % BF_STAGES) * BF_BN + wn * BF_WN + g) * BF_LDS + t;
#pragma unroll
    for (int kk = 0; kk < BF_BK / 4; ++kk) {
      double a[8], b[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = as[i * 8 * BF_LDS + kk * 4];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = bs[j * 8 * BF_LDS + kk * 4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma_884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  cp_async_wait<0>();

  // epilogue: c0, c1 of tile (i, j) are (row g, cols 2t, 2t+1)
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t row = m0 + wm * BF_WM + i * 8 + g;
    if (row >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + wn * BF_WN + j * 8 + 2 * t;
      double* o = D + row * Kpad + col;
      if (col + 1 < n) *reinterpret_cast<double2*>(o) = make_double2(acc[i][j][0], acc[i][j][1]);
      else if (col < n) o[0] = acc[i][j][0];
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Small batches (the reference's own call shape is ONE record per grid, dardel/benes_bernoulli/brute_force.py:51-55):
// the contraction degenerates to a matrix-vector product that a 128 x 128 tile cannot fill the machine with.  Here
// every warp owns output rows i, streams row i of Pw (L2-resident) with 16-byte loads and dots it with up to BT
// densities held in shared memory; deterministic lane-strided partial sums + butterfly reduction.
// ------------------------------------------------------------------------------------------------------------------
constexpr int BF_GEMV_THREADS = 256;
constexpr int BF_GEMV_MAX_B = 64;     // batches up to this size take the matrix-vector path (in chunks of <= 8)

template <int BT>
__global__ void __launch_bounds__(BF_GEMV_THREADS)
bf_gemv_kernel(const double* __restrict__ A, const double* __restrict__ Pw, double* __restrict__ D, int64_t M, int n,
               int Kpad, int64_t b0) {
  extern __shared__ __align__(16) double gv_smem[];   // [BT][Kpad]
  for (int e = threadIdx.x * 2; e < BT * Kpad; e += BF_GEMV_THREADS * 2) {
    const int b = e / Kpad, j = e - b * Kpad;
    const bool ok = b0 + b < M;
    const double2 v = ok ? *reinterpret_cast<const double2*>(A + (b0 + b) * Kpad + j) : make_double2(0.0, 0.0);
    *reinterpret_cast<double2*>(gv_smem + e) = v;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kWarps = BF_GEMV_THREADS / 32;
  for (int i = blockIdx.x * kWarps + warp; i < n; i += gridDim.x * kWarps) {
    const double* prow = Pw + (int64_t)i * Kpad;
    double acc[BT];
#pragma unroll
    for (int b = 0; b < BT; ++b) acc[b] = 0.0;
    for (int j = lane * 2; j < Kpad; j += 64) {
      const double2 p = *reinterpret_cast<const double2*>(prow + j);
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        const double2 a = *reinterpret_cast<const double2*>(gv_smem + b * Kpad + j);
        acc[b] = fma(p.y, a.y, fma(p.x, a.x, acc[b]));
      }
    }
#pragma unroll
    for (int b = 0; b < BT; ++b) {
      double v = acc[b];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && b0 + b < M) D[(b0 + b) * Kpad + i] = v;
    }
  }
}

template <int BT>
static cudaError_t launch_gemv(const double* A, const double* Pw, double* D, int64_t M, int n, int Kpad, int64_t b0,
                               int sms, cudaStream_t s) {
  const size_t smem = sizeof(double) * BT * Kpad;
  static bool configured[64] = {};      // the opt-in is per device: tracked per device ordinal
  int dev = 0;
  if (cudaError_t e = cudaGetDevice(&dev)) return e;
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(bf_gemv_kernel<BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  const int rows_per_cta = BF_GEMV_THREADS / 32;
  int grid = (n + rows_per_cta - 1) / rows_per_cta;
  const int cap = sms * (smem > 100 * 1024 ? 1 : smem > 48 * 1024 ? 2 : 4);
  if (grid > cap) grid = cap;
  bf_gemv_kernel<BT><<<grid, BF_GEMV_THREADS, smem, s>>>(A, Pw, D, M, n, Kpad, b0);
  return cudaGetLastError();
}

// one integration sub-step for a small batch: chunks of up to 8 densities
static cudaError_t gemv_substep(const double* A, const double* Pw, double* D, int64_t M, int n, int Kpad, int sms,
                                cudaStream_t s, int64_t* launches) {
  for (int64_t b0 = 0; b0 < M; b0 += 8) {
    const int64_t left = M - b0;
    cudaError_t e;
    if (left >= 5) e = launch_gemv<8>(A, Pw, D, M, n, Kpad, b0, sms, s);
    else if (left >= 3) e = launch_gemv<4>(A, Pw, D, M, n, Kpad, b0, sms, s);
    else if (left == 2) e = launch_gemv<2>(A, Pw, D, M, n, Kpad, b0, sms, s);
    else e = launch_gemv<1>(A, Pw, D, M, n, Kpad, b0, sms, s);
    if (e != cudaSuccess) return e;
    ++*launches;
  }
  return cudaSuccess;
}

// ------------------------------------------------------------------------------------------------------------------
// Transition operator: per column j (m_j, 1/s_j, w_j / (s_j sqrt(2 pi))), then Pw[i][j] (zero in the padding).
// brute_force.py:69-78 (Euler--Maruyama or tme.mean_and_cov on the sub-step ddt) and :84 (norm.pdf * trapz weight).
// ------------------------------------------------------------------------------------------------------------------
__global__ void bf_columns_kernel(int n, int Kpad, int trans_id, int drift_id, int order, double c, double ddt,
                                  const double* __restrict__ prm, const double* __restrict__ xs,
                                  double* __restrict__ col_m, double* __restrict__ col_is, double* __restrict__ col_cf) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= Kpad) return;
  if (j >= n) { col_m[j] = 0.0; col_is[j] = 0.0; col_cf[j] = 0.0; return; }
  double p[MFS_MAX_PARAMS];
#pragma unroll
  for (int k = 0; k < MFS_MAX_PARAMS; ++k) p[k] = prm[k];
  double mean, var;
  normal_mean_var(trans_id, drift_id, order, xs[j], c, ddt, p, mean, var);
  const double s = sqrt(var);   // negative variance -> NaN column, like jnp.sqrt(_cov)
  const double lo = (j > 0) ? xs[j] - xs[j - 1] : 0.0, hi = (j + 1 < n) ? xs[j + 1] - xs[j] : 0.0;
  col_m[j] = mean;
  col_is[j] = 1.0 / s;
  col_cf[j] = 0.5 * (lo + hi) / (s * 2.5066282746310002);
}

__global__ void bf_operator_kernel(int n, int Npad, int Kpad, const double* __restrict__ xs,
                                   const double* __restrict__ col_m, const double* __restrict__ col_is,
                                   const double* __restrict__ col_cf, double* __restrict__ Pw) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j >= Kpad) return;
  double v = 0.0;
  if (i < n && j < n) {
    const double z = (xs[i] - col_m[j]) * col_is[j];
    v = col_cf[j] * exp(-0.5 * z * z);
  }
  Pw[(int64_t)i * Kpad + j] = v;
}

// out[j][i] = in[i][j] for i, j < n (both [Npad][Kpad], padding left as it is: zero), 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(256) bf_transpose_kernel(int n, int Kpad, const double* __restrict__ in, double* __restrict__ out) {
  __shared__ double tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  for (int r = ty; r < 32; r += 8) {
    const int i = i0 + r, j = j0 + tx;
    tile[r][tx] = (i < n && j < n) ? in[(int64_t)i * Kpad + j] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int j = j0 + r, i = i0 + tx;
    if (i < n && j < n) out[(int64_t)j * Kpad + i] = tile[tx][r];
  }
}

// S[b][0..Kpad) <- init_ps[b][0..n), zero padding
__global__ void bf_load_state_kernel(int64_t B, int n, int Kpad, const double* __restrict__ init_ps, int64_t stride,
                                     double* __restrict__ S) {
  const int64_t b = blockIdx.x;
  const int j = blockIdx.y * blockDim.x + threadIdx.x;
  if (b >= B || j >= Kpad) return;
  S[b * Kpad + j] = (j < n) ? init_ps[b * stride + j] : 0.0;
}

MFS_DEV double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double tot = 0.0;
  for (int w = 0; w < nw; ++w) tot += red[w];   // same order in every thread: deterministic
  return tot;
}

struct BfUpdateArgs {
  int n, Kpad, meas_id, ys_dtype, out_full;
  int64_t B, T, t;
  const double* xs;
  const double* meas_params; int64_t meas_param_stride;
  const void* ys; int64_t ys_stride_b, ys_stride_t;
  const double* S_in;   // predicted density [B][Kpad]
  double* S_out;        // posterior density [B][Kpad]
  double* pdfs_out;     // FULL [B][T][n] or LAST [B][n] (written when out_full or t == T-1)
  double* nell;         // [B] accumulated, may be NULL
};

// Bayes update of one filter per CTA: brute_force.py:135
__global__ void __launch_bounds__(256) bf_update_kernel(const BfUpdateArgs P) {
  extern __shared__ double u_sm[];   // n
  __shared__ double red[8];
  const int64_t b = blockIdx.x;
  const double* ps = P.S_in + b * P.Kpad;
  double mp[MFS_MAX_PARAMS];
#pragma unroll
  for (int k = 0; k < MFS_MAX_PARAMS; ++k) mp[k] = P.meas_params[b * P.meas_param_stride + k];
  if (P.meas_id == MFS_MEAS_BERNOULLI_LOGISTIC_CUBIC) mp[2] = 1.0 / mp[0];
  const int64_t yoff = b * P.ys_stride_b + P.t * P.ys_stride_t;
  MeasStep st;
  st.y = (P.ys_dtype == MFS_YS_U8)    ? (double)reinterpret_cast<const unsigned char*>(P.ys)[yoff]
         : (P.ys_dtype == MFS_YS_I32) ? (double)reinterpret_cast<const int*>(P.ys)[yoff]
                                      : reinterpret_cast<const double*>(P.ys)[yoff];
  st.c0 = (P.meas_id == MFS_MEAS_BERNOULLI_LOGISTIC_CUBIC) ? 0.0 : meas_step_constant(P.meas_id, st.y, mp[1]);
  for (int i = threadIdx.x; i < P.n; i += blockDim.x) u_sm[i] = measurement_pdf(P.meas_id, st, P.xs[i], mp) * ps[i];
  __syncthreads();
  double part = 0.0;
  for (int i = threadIdx.x; i + 1 < P.n; i += blockDim.x) part += (P.xs[i + 1] - P.xs[i]) * (u_sm[i + 1] + u_sm[i]);
  const double c = 0.5 * block_sum(part, red);
  double* so = P.S_out + b * P.Kpad;
  const bool emit = P.out_full || P.t == P.T - 1;
  double* po = P.out_full ? P.pdfs_out + (b * P.T + P.t) * P.n : P.pdfs_out + b * P.n;
  for (int i = threadIdx.x; i < P.n; i += blockDim.x) {
    const double v = u_sm[i] / c;
    so[i] = v;
    if (emit) po[i] = v;
  }
  if (P.nell && threadIdx.x == 0) P.nell[b] = (P.t == 0 ? 0.0 : P.nell[b]) - log(c);
}

// ------------------------------------------------------------------------------------------------------------------
// 'kolmogorov': explicit Euler on the Fokker--Planck operator, jnp.gradient stencils, constant dispersion
// (brute_force.py:96-113).  One CTA per filter; shared memory: ps, d1, d2, a, a'  (5 n doubles).
// ------------------------------------------------------------------------------------------------------------------
struct BfKolmogorovArgs {
  int n, drift_id, steps;
  double ddt, gamma;
  const double* xs;
  const double* trans_params;
  double* S;     // [B][Kpad] in/out
  int Kpad;
};

MFS_DEV double grad_at(const double* f, int i, int n, double dx) {
  if (i == 0) return (f[1] - f[0]) / dx;
  if (i == n - 1) return (f[n - 1] - f[n - 2]) / dx;
  return (f[i + 1] - f[i - 1]) / (2.0 * dx);
}

__global__ void __launch_bounds__(256) bf_kolmogorov_kernel(const BfKolmogorovArgs P) {
  extern __shared__ double ksm[];
  const int n = P.n;
  double *ps = ksm, *d1 = ksm + n, *d2 = ksm + 2 * n, *a = ksm + 3 * n, *da = ksm + 4 * n;
  double* S = P.S + (int64_t)blockIdx.x * P.Kpad;
  const double dx = P.xs[1] - P.xs[0];   // brute_force.py:66
  double prm[MFS_MAX_PARAMS];
#pragma unroll
  for (int k = 0; k < MFS_MAX_PARAMS; ++k) prm[k] = P.trans_params[k];
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    ps[i] = S[i];
    const Jet j = drift_jet(P.drift_id, P.xs[i], prm);
    a[i] = j.a0;
    da[i] = j.a1;
  }
  __syncthreads();
  for (int s = 0; s < P.steps; ++s) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) d1[i] = grad_at(ps, i, n, dx);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) d2[i] = grad_at(d1, i, n, dx);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const double part1 = -(da[i] * ps[i] + a[i] * d1[i]);
      ps[i] = ps[i] + (part1 + 0.5 * (P.gamma * d2[i])) * P.ddt;
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) S[i] = ps[i];
}

// ------------------------------------------------------------------------------------------------------------------
// FP64 tensor-pipe peak: 16 independent DMMA accumulator tiles per warp, no memory traffic.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dmma_peak_kernel(int iters, double seed, double* sink) {
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) { c[i][0] = seed + i; c[i][1] = seed - i; }
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) dmma_884(c[i][0], c[i][1], a, b);
  }
  double r = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) r += c[i][0] + c[i][1];
  if (r == 12345.678) sink[0] = r;
}

}  // namespace mfs

using namespace mfs;

extern "C" {

int64_t mfs_brute_force_workspace_bytes_ex(int32_t n_grid, int64_t B, int32_t pred_method, int32_t flags) {
  if (n_grid < 3 || B < 0) return -1;
  const int64_t Kpad = round_up(n_grid, BF_BK), Npad = round_up(n_grid, BF_BN);
  int64_t doubles = 2 * B * Kpad;                                 // two state buffers
  if (pred_method != MFS_BF_KOLMOGOROV) {
    doubles += Npad * Kpad + 3 * Kpad;                            // Pw + column tables
    if (flags & MFS_BF_FLAG_POWER_OPERATOR) doubles += 5 * Npad * Kpad;      // P, Q = P^T, R ping-pong (see mfs_brute_force)
  }
  return doubles * (int64_t)sizeof(double) + 256;
}

int64_t mfs_brute_force_workspace_bytes(int32_t n_grid, int64_t B, int32_t pred_method) {
  return mfs_brute_force_workspace_bytes_ex(n_grid, B, pred_method, 0);
}

int mfs_brute_force(const mfs_brute_force_args* a, void* stream) {
  if (!a) return fail("args is NULL");
  if (a->abi_version != MFS_ABI_VERSION) return fail("abi_version %d != %d", a->abi_version, MFS_ABI_VERSION);
  if (a->pred_method < MFS_BF_CHAPMAN_EULER || a->pred_method > MFS_BF_KOLMOGOROV) return fail("unknown pred_method %d", a->pred_method);
  if (a->pred_method == MFS_BF_CHAPMAN_TME && (a->tme_order < 1 || a->tme_order > 3)) return fail("tme_order=%d outside [1, 3]", a->tme_order);
  if (a->integration_steps < 1) return fail("integration_steps must be >= 1");
  if (a->n_grid < 3) return fail("n_grid must be >= 3");
  if (a->drift_id < MFS_DRIFT_BENES || a->drift_id > MFS_DRIFT_LINEAR) return fail("unknown drift_id %d", a->drift_id);
  if (a->meas_id < MFS_MEAS_BERNOULLI_LOGISTIC_CUBIC || a->meas_id > MFS_MEAS_GAUSSIAN) return fail("unknown meas_id %d", a->meas_id);
  if (a->ys_dtype < MFS_YS_U8 || a->ys_dtype > MFS_YS_F64) return fail("unknown ys_dtype %d", a->ys_dtype);
  if (a->out_mode != MFS_OUT_FULL && a->out_mode != MFS_OUT_LAST) return fail("out_mode must be FULL or LAST");
  if (a->B < 0 || a->T < 0) return fail("negative B or T");
  if (!(a->dt > 0.0)) return fail("dt must be > 0");
  if (a->B == 0 || a->T == 0) return 0;
  if (!a->trans_params || !a->meas_params || !a->xs || !a->init_ps || !a->ys || !a->pdfs_out || !a->workspace)
    return fail("NULL pointer argument");
  if (a->flags & ~MFS_BF_FLAG_POWER_OPERATOR) return fail("unknown flags 0x%x", a->flags);
  const int64_t need = mfs_brute_force_workspace_bytes_ex(a->n_grid, a->B, a->pred_method, a->flags);
  if (a->workspace_bytes < need) return fail("workspace too small: %lld < %lld bytes", (long long)a->workspace_bytes, (long long)need);
  if (a->B > 65535LL * BF_BM) return fail("B too large for one call (max %d)", 65535 * BF_BM);
  if (a->n_grid > 65408) return fail("n_grid too large (max 65408)");

  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int n = a->n_grid;
  const int Kpad = (int)round_up(n, BF_BK), Npad = (int)round_up(n, BF_BN);
  const int64_t B = a->B;
  double* ws = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(a->workspace) + 255) & ~uintptr_t(255));
  double* S[2] = {ws, ws + B * Kpad};
  double* Pw = ws + 2 * B * Kpad;
  double *col_m = Pw + (int64_t)Npad * Kpad, *col_is = col_m + Kpad, *col_cf = col_is + Kpad;
  const double ddt = a->dt / a->integration_steps;
  const double c = 0.5 * a->dispersion * a->dispersion;
  int64_t launches = 0;
#define BF_CHECK()                                                                              \
  do {                                                                                          \
    cudaError_t e__ = cudaGetLastError();                                                       \
    if (e__ != cudaSuccess) return fail("brute-force launch failed: %s", cudaGetErrorString(e__)); \
  } while (0)

  const bool chapman = a->pred_method != MFS_BF_KOLMOGOROV;
  static bool configured[64] = {};      // the opt-in is per device: tracked per device ordinal
  int cur_dev = 0;
  if (cudaGetDevice(&cur_dev) != cudaSuccess) return fail("cudaGetDevice failed");
  if (cur_dev < 0 || cur_dev >= 64 || !configured[cur_dev]) {
    cudaError_t e = cudaFuncSetAttribute(bf_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BF_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(bf_kolmogorov_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return fail("cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    if (cur_dev >= 0 && cur_dev < 64) configured[cur_dev] = true;
  }
  if (chapman) {
    const int trans = a->pred_method == MFS_BF_CHAPMAN_EULER ? MFS_TRANS_EULER : MFS_TRANS_TME_NORMAL;
    bf_columns_kernel<<<(Kpad + 127) / 128, 128, 0, s>>>(n, Kpad, trans, a->drift_id, a->tme_order, c, ddt, a->trans_params,
                                                         a->xs, col_m, col_is, col_cf);
    BF_CHECK();
    bf_operator_kernel<<<dim3((Kpad + 255) / 256, Npad), 256, 0, s>>>(n, Npad, Kpad, a->xs, col_m, col_is, col_cf, Pw);
    BF_CHECK();
    launches += 2;
  } else {
    if ((size_t)n * 5 * sizeof(double) > 200 * 1024) return fail("kolmogorov: n_grid=%d too large for shared memory (max %d)", n, 200 * 1024 / 40);
  }
  // MFS_BF_FLAG_POWER_OPERATOR: the sub-steps of a time step apply the SAME linear operator (brute_force.py:115-122: the
  // scan body does not depend on the sub-step), so  S (Pw^T)^k = S (Pw^k)^T : Pw^k by binary powering on the DMMA GEMM
  // (log2 k + popcount k - 1 products of n x n matrices, once per call), then ONE contraction per time step instead
  // of k.  The kernel computes X Y^T, so the transpose Q = P^T is kept next to P: P P = gemm(P, Q), R P = gemm(R, Q).
  const double* Pw_step = Pw;
  int steps_eff = a->integration_steps;
  if (chapman && (a->flags & MFS_BF_FLAG_POWER_OPERATOR) && a->integration_steps > 1) {
    const int64_t mat = (int64_t)Npad * Kpad;
    double* pool = col_cf + Kpad;
    if (cudaMemsetAsync(pool, 0, sizeof(double) * 5 * mat, s) != cudaSuccess) return fail("cudaMemsetAsync failed");
    double* Pb[2] = {Pw, pool};
    double* Qb[2] = {pool + mat, pool + 2 * mat};
    double* Rb[2] = {pool + 3 * mat, pool + 4 * mat};
    const dim3 tgrid((unsigned)((n + 31) / 32), (unsigned)((n + 31) / 32));
    bf_transpose_kernel<<<tgrid, 256, 0, s>>>(n, Kpad, Pb[0], Qb[0]);
    BF_CHECK();
    launches += 1;
    const dim3 pgrid((unsigned)(Npad / BF_BN), (unsigned)((n + BF_BM - 1) / BF_BM));
    int pc = 0, rc = -1;
    for (int k = a->integration_steps;;) {
      if (k & 1) {
        if (rc < 0) {
          if (cudaMemcpyAsync(Rb[0], Pb[pc], sizeof(double) * mat, cudaMemcpyDeviceToDevice, s) != cudaSuccess)
            return fail("cudaMemcpyAsync failed");
          rc = 0;
        } else {
          bf_gemm_kernel<<<pgrid, BF_THREADS, BF_SMEM, s>>>(Rb[rc], Qb[pc], Rb[rc ^ 1], n, n, Kpad);     // R P
          rc ^= 1;
          launches += 1;
        }
      }
      k >>= 1;
      if (!k) break;
      bf_gemm_kernel<<<pgrid, BF_THREADS, BF_SMEM, s>>>(Pb[pc], Qb[pc], Pb[pc ^ 1], n, n, Kpad);         // P P
      bf_transpose_kernel<<<tgrid, 256, 0, s>>>(n, Kpad, Pb[pc ^ 1], Qb[pc ^ 1]);                       // Q = (P P)^T
      pc ^= 1;
      launches += 2;
    }
    BF_CHECK();
    Pw_step = Rb[rc];
    steps_eff = 1;
  }
  bf_load_state_kernel<<<dim3((unsigned)B, (Kpad + 255) / 256), 256, 0, s>>>(B, n, Kpad, a->init_ps, a->init_ps_stride, S[0]);
  BF_CHECK();
  // the padding columns of the second buffer must be zero too (they meet the zero columns of Pw)
  if (cudaMemsetAsync(S[1], 0, sizeof(double) * B * Kpad, s) != cudaSuccess) return fail("cudaMemsetAsync failed");
  launches += 1;

  int cur = 0;
  // small batches: matrix-vector kernels (the 8-density chunk needs 8 * Kpad doubles of shared memory)
  const bool use_gemv = chapman && B <= BF_GEMV_MAX_B && (size_t)Kpad * 8 * sizeof(double) <= 200 * 1024;
  int sms = 148;
  {
    int devid = 0;
    if (cudaGetDevice(&devid) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, devid);
  }
  const dim3 ggrid((unsigned)(Npad / BF_BN), (unsigned)((B + BF_BM - 1) / BF_BM));
  for (int64_t t = 0; t < a->T; ++t) {
    if (chapman) {
      if (use_gemv) {
        for (int k = 0; k < steps_eff; ++k) {
          cudaError_t e = gemv_substep(S[cur], Pw_step, S[cur ^ 1], B, n, Kpad, sms, s, &launches);
          if (e != cudaSuccess) return fail("brute-force launch failed: %s", cudaGetErrorString(e));
          cur ^= 1;
        }
      } else {
        for (int k = 0; k < steps_eff; ++k) {
          bf_gemm_kernel<<<ggrid, BF_THREADS, BF_SMEM, s>>>(S[cur], Pw_step, S[cur ^ 1], B, n, Kpad);
          cur ^= 1;
        }
        BF_CHECK();
        launches += steps_eff;
      }
    } else {
      BfKolmogorovArgs k;
      k.n = n; k.drift_id = a->drift_id; k.steps = a->integration_steps;
      k.ddt = ddt; k.gamma = a->dispersion * a->dispersion;
      k.xs = a->xs; k.trans_params = a->trans_params; k.S = S[cur]; k.Kpad = Kpad;
      bf_kolmogorov_kernel<<<(unsigned)B, 256, (size_t)n * 5 * sizeof(double), s>>>(k);
      BF_CHECK();
      launches += 1;
    }
    BfUpdateArgs u;
    u.n = n; u.Kpad = Kpad; u.meas_id = a->meas_id; u.ys_dtype = a->ys_dtype; u.out_full = a->out_mode == MFS_OUT_FULL;
    u.B = B; u.T = a->T; u.t = t;
    u.xs = a->xs; u.meas_params = a->meas_params; u.meas_param_stride = a->meas_param_stride;
    u.ys = a->ys; u.ys_stride_b = a->ys_stride_b; u.ys_stride_t = a->ys_stride_t;
    u.S_in = S[cur]; u.S_out = S[cur ^ 1]; u.pdfs_out = a->pdfs_out; u.nell = a->nell_out;
    bf_update_kernel<<<(unsigned)B, 256, (size_t)n * sizeof(double), s>>>(u);
    BF_CHECK();
    cur ^= 1;
    launches += 1;
  }
  count_launches(launches);
  return 0;
#undef BF_CHECK
}

int mfs_dmma_peak(int device, int32_t iters, double* flops, double* ms) {
  if (!flops) return fail("flops is NULL");
  if (cudaSetDevice(device) != cudaSuccess) return fail("cudaSetDevice failed");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail("cudaGetDeviceProperties failed");
  double* sink = nullptr;
  if (cudaMalloc(&sink, sizeof(double)) != cudaSuccess) return fail("cudaMalloc failed");
  const int grid = prop.multiProcessorCount * 4;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  dmma_peak_kernel<<<grid, 256>>>(iters / 4 + 1, 1.0, sink);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    dmma_peak_kernel<<<grid, 256>>>(iters, 1.0, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float t;
    cudaEventElapsedTime(&t, e0, e1);
    if (t < best) best = t;
  }
  count_launches(6);
  const cudaError_t e = cudaGetLastError();
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  if (e != cudaSuccess) return fail("dmma_peak launch failed: %s", cudaGetErrorString(e));
  const double mmas = (double)grid * 8.0 * (double)iters * 16.0;   // warps x iters x tiles
  *flops = 512.0 * mmas / (best * 1e-3);
  if (ms) *ms = best;
  return 0;
}

}  // extern "C"
