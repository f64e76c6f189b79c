// C ABI of libmfs_b200.so (declared in include/mfs_b200.h): validation, dispatch, host-buffer pipeline.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "common.h"
#include "filter1d.cuh"
#include "filter_nd.cuh"

namespace mfs {

static thread_local std::string g_err;
static std::atomic<int64_t> g_launches{0};

void count_launches(int64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int fail(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return -1;
}

#define MFS_CUDA(call)                                                                       \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess) return fail("%s failed: %s", #call, cudaGetErrorString(e__));    \
  } while (0)

// Device workspace of the host-buffer entry point.  Blocks are cached per (device, size) between calls so that a
// pipeline calling mfs_filter_1d_host repeatedly with the same shapes pays cudaMalloc/cudaFree (which synchronise the
// device) once; mfs_release_cached_memory() gives everything back.
struct WsBlock { int device; size_t bytes; void* ptr; bool busy; };
static std::mutex g_ws_mutex;
static std::vector<WsBlock> g_ws;

static cudaError_t ws_alloc(int device, size_t bytes, void** out) {
  std::lock_guard<std::mutex> lock(g_ws_mutex);
  for (WsBlock& b : g_ws)
    if (!b.busy && b.device == device && b.bytes == bytes) { b.busy = true; *out = b.ptr; return cudaSuccess; }
  // a miss means the shapes changed: idle big blocks of the old shapes go back to the driver instead of accumulating
  for (size_t i = 0; i < g_ws.size();) {
    if (!g_ws[i].busy && g_ws[i].device == device && g_ws[i].bytes >= (64u << 20)) {
      cudaFree(g_ws[i].ptr);
      g_ws.erase(g_ws.begin() + i);
    } else {
      ++i;
    }
  }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {   // give cached idle blocks back and retry once
    for (size_t i = 0; i < g_ws.size();) {
      if (!g_ws[i].busy && g_ws[i].device == device) { cudaFree(g_ws[i].ptr); g_ws.erase(g_ws.begin() + i); } else ++i;
    }
    cudaGetLastError();
    e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return e;
  }
  g_ws.push_back({device, bytes, p, true});
  *out = p;
  return cudaSuccess;
}

static void ws_free(void* p) {
  if (!p) return;
  std::lock_guard<std::mutex> lock(g_ws_mutex);
  for (WsBlock& b : g_ws)
    if (b.ptr == p) { b.busy = false; return; }
}

}  // namespace mfs
extern "C" int64_t mfs_filter_1d_workspace_bytes(int32_t N, int64_t B, int64_t T);
namespace mfs {

static int ys_elem_size(int dtype) { return dtype == MFS_YS_U8 ? 1 : dtype == MFS_YS_I32 ? 4 : 8; }

// Argument checks shared by the device and host entry points (pointer NULL-ness only; both address spaces).
static int validate(const mfs_filter1d_args* a) {
  if (!a) return fail("args is NULL");
  if (a->abi_version != MFS_ABI_VERSION) return fail("abi_version %d != %d", a->abi_version, MFS_ABI_VERSION);
  if (a->N < 2 || a->N > MFS_MAX_N) return fail("N=%d outside [2, %d]", a->N, MFS_MAX_N);
  if (a->mode < MFS_MODE_RAW || a->mode > MFS_MODE_SCALED) return fail("unknown mode %d", a->mode);
  if (a->B < 0 || a->T < 0) return fail("negative B or T");
  if (a->trans_id < MFS_TRANS_TME || a->trans_id > MFS_TRANS_NORMAL_AFFINE) return fail("unknown trans_id %d", a->trans_id);
  if (a->trans_id != MFS_TRANS_NORMAL_AFFINE && (a->drift_id < MFS_DRIFT_BENES || a->drift_id > MFS_DRIFT_LINEAR))
    return fail("unknown drift_id %d", a->drift_id);
  if ((a->trans_id == MFS_TRANS_TME || a->trans_id == MFS_TRANS_TME_NORMAL) && (a->tme_order < 1 || a->tme_order > 3))
    return fail("tme_order=%d outside [1, 3]", a->tme_order);
  if (a->meas_id < MFS_MEAS_BERNOULLI_LOGISTIC_CUBIC || a->meas_id > MFS_MEAS_GAUSSIAN) return fail("unknown meas_id %d", a->meas_id);
  if (a->ys_dtype < MFS_YS_U8 || a->ys_dtype > MFS_YS_F64) return fail("unknown ys_dtype %d", a->ys_dtype);
  if (a->out_mode < MFS_OUT_FULL || a->out_mode > MFS_OUT_MEANVAR) return fail("unknown out_mode %d", a->out_mode);
  if (!(a->dt > 0.0)) return fail("dt must be > 0");
  if (a->B == 0) return 0;
  if (!a->ms0 || !a->trans_params || !a->meas_params || !a->nell_out) return fail("ms0 / trans_params / meas_params / nell_out must not be NULL");
  if (a->T > 0 && !a->ys) return fail("ys is NULL");
  if (a->mode != MFS_MODE_RAW && !a->mean0) return fail("mean0 is NULL in central/scaled mode");
  if (a->mode == MFS_MODE_SCALED && !a->scale0) return fail("scale0 is NULL in scaled mode");
  // per-step histories are empty when T == 0: only the LAST outputs (one row per filter) are required then
  const bool need_ms = a->out_mode == MFS_OUT_LAST || ((a->out_mode == MFS_OUT_FULL || a->out_mode == MFS_OUT_MEANVAR) && a->T > 0);
  if (need_ms && !a->ms_out) return fail("ms_out is NULL");
  const bool aux = a->out_mode == MFS_OUT_LAST || (a->out_mode == MFS_OUT_FULL && a->T > 0);
  if (aux && a->mode != MFS_MODE_RAW && !a->mean_out) return fail("mean_out is NULL in central/scaled mode");
  if (aux && a->mode == MFS_MODE_SCALED && !a->scale_out) return fail("scale_out is NULL in scaled mode");
  if (a->t_offset < 0) return fail("negative t_offset");
  if (a->grid_records < 0 || (a->grid_records > 0 && a->B % a->grid_records != 0))
    return fail("grid_records=%lld must be 0 or a divisor of B=%lld", (long long)a->grid_records, (long long)a->B);
  return 0;
}

// Compile-time specialisations: the headline model (Benes TME + Bernoulli-logistic likelihood) and the estimation
// objective (Normal family + Poisson-softplus likelihood); everything else runs the run-time-switched instances.
static int pick_kind(const mfs_filter1d_args& a, int* meas_ct) {
  *meas_ct = -1;
  if (a.trans_id == MFS_TRANS_TME) {
    if (a.drift_id == MFS_DRIFT_BENES && a.dispersion == 1.0 && a.meas_id == MFS_MEAS_BERNOULLI_LOGISTIC_CUBIC) {
      *meas_ct = MFS_MEAS_BERNOULLI_LOGISTIC_CUBIC;
      return KIND_BENES_TME;
    }
    return KIND_TME;
  }
  if (a.meas_id == MFS_MEAS_POISSON_SOFTPLUS) *meas_ct = MFS_MEAS_POISSON_SOFTPLUS;
  return KIND_NORMAL;
}

static cudaError_t dispatch(const mfs_filter1d_args& a, const SegInfo& g, cudaStream_t s) {
  int meas_ct = -1;
  const int kind = pick_kind(a, &meas_ct);
  switch (a.N) {
#define MFS_CASE(n) case n: return launch_filter1d<n>(a, g, kind, meas_ct, s);
    MFS_CASE(2) MFS_CASE(3) MFS_CASE(4) MFS_CASE(5) MFS_CASE(6) MFS_CASE(7) MFS_CASE(8) MFS_CASE(9)
    MFS_CASE(10) MFS_CASE(11) MFS_CASE(12) MFS_CASE(13) MFS_CASE(14) MFS_CASE(15)
#undef MFS_CASE
    default: return cudaErrorInvalidValue;
  }
}

// nell + gradient (filter1d_grad.cuh), one translation unit per N
template <int N>
cudaError_t launch_filter1d_grad(const mfs_filter1d_args& a, const GradInfo& g, cudaStream_t stream);

// NaN tails of the filters that failed (JAX semantics: everything after a failed Cholesky is NaN until the end of the
// scan).  Written here, one warp per failed filter with coalesced stores, instead of by the failing thread itself --
// which would hold its whole warp for up to T - t serial store rounds.
__global__ void __launch_bounds__(128) nan_fill_kernel(const mfs_filter1d_args P) {
  const int64_t b = ((int64_t)blockIdx.x * 128 + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= P.B) return;
  int64_t s = P.status_out[b];
  if (s < 0) return;
  s -= P.t_offset;             // status holds the absolute step; a filter that failed in an earlier chunk starts at 0
  if (s < 0) s = 0;
  const int M = (P.out_mode == MFS_OUT_MEANVAR) ? 2 : 2 * P.N;
  const double qnan = nan("");
  double* o = P.ms_out + b * P.ms_stride_b;
  if (P.ms_stride_t == M) {
    double* base = o + s * M;
    const int64_t cnt = (P.T - s) * M;
    for (int64_t e = lane; e < cnt; e += 32) base[e] = qnan;
  } else {
    for (int64_t t = s; t < P.T; ++t)
      for (int p = lane; p < M; p += 32) o[t * P.ms_stride_t + p] = qnan;
  }
  if (P.out_mode == MFS_OUT_MEANVAR) return;
  if (P.mode != MFS_MODE_RAW && P.mean_out)
    for (int64_t t = s + lane; t < P.T; t += 32) P.mean_out[b * P.aux_stride_b + t] = qnan;
  if (P.mode == MFS_MODE_SCALED && P.scale_out)
    for (int64_t t = s + lane; t < P.T; t += 32) P.scale_out[b * P.aux_stride_b + t] = qnan;
}

// T == 0 (or a carried state with nothing to filter): the reference returns empty histories and nell = 0; with a
// carry the state passes through unchanged.
__global__ void __launch_bounds__(128) empty_scan_kernel(const mfs_filter1d_args P, int state_doubles) {
  const int64_t b = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (b >= P.B) return;
  const int64_t row = P.grid_records > 0 ? b / P.grid_records : b;
  double nell = 0.0;
  int status = -1;
  if (P.carry_in) {
    const double* ci = P.carry_in + b * state_doubles;
    nell = ci[state_doubles - 2];
    const double flag = ci[state_doubles - 1];
    if (flag < 0.0) { status = (int)(-flag) - 1; nell = nan(""); }
    if (P.carry_out)
      for (int k = 0; k < state_doubles; ++k) P.carry_out[b * state_doubles + k] = ci[k];
  } else if (P.carry_out) {
    const int M = 2 * P.N;
    double* co = P.carry_out + b * state_doubles;
    for (int k = 0; k < M; ++k) co[k] = P.ms0[row * P.ms0_stride + k];
    for (int k = M; k < 2 * M; ++k) co[k] = 0.0;
    co[2 * M] = P.mode != MFS_MODE_RAW ? P.mean0[row * P.mean0_stride] : 0.0;
    co[2 * M + 1] = P.mode == MFS_MODE_SCALED ? P.scale0[row * P.scale0_stride] : 1.0;
    co[2 * M + 2] = 0.0;
    co[2 * M + 3] = 0.0;
  }
  if (P.out_mode == MFS_OUT_LAST) {
    // no step ran: the "last" moments are the initial ones
    const int M = 2 * P.N;
    const double* src = P.carry_in ? P.carry_in + b * state_doubles : P.ms0 + row * P.ms0_stride;
    for (int k = 0; k < M; ++k) P.ms_out[b * P.ms_stride_b + k] = src[k];
    if (P.mode != MFS_MODE_RAW && P.mean_out)
      P.mean_out[b * P.aux_stride_b] = P.carry_in ? P.carry_in[b * state_doubles + 2 * M] : P.mean0[row * P.mean0_stride];
    if (P.mode == MFS_MODE_SCALED && P.scale_out)
      P.scale_out[b * P.aux_stride_b] = P.carry_in ? P.carry_in[b * state_doubles + 2 * M + 1] : P.scale0[row * P.scale0_stride];
  }
  P.nell_out[b] = nell;
  if (P.status_out) P.status_out[b] = status;
}

constexpr int64_t kDefaultSegmentSteps = 64;

static int launch_device(const mfs_filter1d_args& a, cudaStream_t s) {
  if (a.B == 0) return 0;
  const bool per_step = a.out_mode == MFS_OUT_FULL || a.out_mode == MFS_OUT_MEANVAR;
  if (a.out_mode != MFS_OUT_NONE) {
    if ((reinterpret_cast<uintptr_t>(a.ms_out) & 15) || (a.ms_stride_b & 1) || (per_step && (a.ms_stride_t & 1)))
      return fail("ms_out must be 16-byte aligned with even strides (vector stores)");
  }
  if (a.B > (int64_t)kBlock * 0x7fffffffLL) return fail("B too large for one launch");
  if (a.T == 0) {   // nothing to scan: empty histories, nell = 0 (the reference's scan over an empty ys)
    empty_scan_kernel<<<(unsigned)((a.B + 127) / 128), 128, 0, s>>>(a, 4 * a.N + 4);
    cudaError_t e0 = cudaGetLastError();
    if (e0 != cudaSuccess) return fail("empty-scan launch failed: %s", cudaGetErrorString(e0));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
  }
  SegInfo g = {};
  g.t0 = 0;
  g.t1 = a.T;
  g.state_in = a.carry_in;
  g.state_out = a.carry_out;
  g.last = 1;
  g.defer_nan_fill = (per_step && a.status_out != nullptr && a.T > 1) ? 1 : 0;
  auto fill_tails = [&]() -> int {
    if (!g.defer_nan_fill) return 0;
    const int64_t ctas = (a.B * 32 + 127) / 128;
    if (ctas > 0x7fffffffLL) return fail("B too large for the NaN-tail launch");
    nan_fill_kernel<<<(unsigned)ctas, 128, 0, s>>>(a);
    cudaError_t e2 = cudaGetLastError();
    if (e2 != cudaSuccess) return fail("nan_fill launch failed: %s", cudaGetErrorString(e2));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
  };
  const int64_t seg = a.segment_steps > 0 ? a.segment_steps : kDefaultSegmentSteps;
  if (!a.workspace || a.T < 2 * seg || a.B > 0x7fffffffLL) {     // one launch for the whole scan
    cudaError_t e = dispatch(a, g, s);
    if (e != cudaSuccess) return fail("kernel launch failed: %s", cudaGetErrorString(e));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return fill_tails();
  }
  // segmented execution with live-filter compaction (filter1d.cuh: SegInfo)
  const int64_t need = mfs_filter_1d_workspace_bytes(a.N, a.B, a.T);
  if (a.workspace_bytes < need) return fail("workspace too small: %lld < %lld bytes", (long long)a.workspace_bytes, (long long)need);
  const int64_t nseg = (a.T + seg - 1) / seg;
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(a.workspace) + 255) & ~uintptr_t(255));
  double* park = reinterpret_cast<double*>(base);
  const size_t state_bytes = ((size_t)a.B * (4 * a.N + 4) * sizeof(double) + 255) & ~size_t(255);
  int32_t* idx[2] = {reinterpret_cast<int32_t*>(base + state_bytes),
                     reinterpret_cast<int32_t*>(base + state_bytes + (((size_t)a.B * 4 + 255) & ~size_t(255)))};
  int32_t* counts = reinterpret_cast<int32_t*>(base + state_bytes + 2 * (((size_t)a.B * 4 + 255) & ~size_t(255)));
  cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)(nseg + 1), s);
  if (e != cudaSuccess) return fail("cudaMemsetAsync failed: %s", cudaGetErrorString(e));
  for (int64_t k = 0; k < nseg; ++k) {
    const bool last = k + 1 == nseg;
    g.t0 = k * seg;
    g.t1 = last ? a.T : (k + 1) * seg;
    g.idx_in = k == 0 ? nullptr : idx[(k - 1) & 1];
    g.count_in = k == 0 ? nullptr : counts + k;
    g.idx_out = last ? nullptr : idx[k & 1];
    g.count_out = last ? nullptr : counts + k + 1;
    g.state_in = k == 0 ? a.carry_in : park;
    g.state_out = last ? a.carry_out : park;
    g.last = last ? 1 : 0;
    e = dispatch(a, g, s);
    if (e != cudaSuccess) return fail("kernel launch failed: %s", cudaGetErrorString(e));
  }
  g_launches.fetch_add(nseg, std::memory_order_relaxed);
  return fill_tails();
}

// ---------------------------------------------------------------------------------------------------------------------
// batched quadrature kernel (mfs/one_dim/quadtures.py:83-133)
// ---------------------------------------------------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(kBlock) quadrature_kernel(int64_t B, const double* __restrict__ ms,
                                                            const double* __restrict__ mean,
                                                            const double* __restrict__ scale, int sort_nodes, int ldl,
                                                            double* __restrict__ weights, double* __restrict__ nodes) {
  const int64_t b = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (b >= B) return;
  double m[2 * N], w[N], x[N];
#pragma unroll
  for (int p = 0; p < 2 * N; ++p) m[p] = ms[b * 2 * N + p];
  const bool ok = moment_quadrature<N>(m, mean ? mean[b] : 0.0, scale ? scale[b] : 1.0, w, x, ldl != 0);
  if (!ok) {
#pragma unroll
    for (int i = 0; i < N; ++i) { w[i] = nan(""); x[i] = nan(""); }
  } else if (sort_nodes) {
    // odd-even transposition network on (x, w): N rounds, compile-time indices
#pragma unroll
    for (int r = 0; r < N; ++r) {
#pragma unroll
      for (int i = (r & 1); i + 1 < N; i += 2) {
        if (x[i] > x[i + 1]) {
          const double tx = x[i]; x[i] = x[i + 1]; x[i + 1] = tx;
          const double tw = w[i]; w[i] = w[i + 1]; w[i + 1] = tw;
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < N; ++i) { weights[b * N + i] = w[i]; nodes[b * N + i] = x[i]; }
}

template <int N>
static cudaError_t launch_quadrature(int64_t B, const double* ms, const double* mean, const double* scale, int sort,
                                     int ldl, double* w, double* x, cudaStream_t s) {
  const unsigned grid = (unsigned)((B + kBlock - 1) / kBlock);
  quadrature_kernel<N><<<grid, kBlock, 0, s>>>(B, ms, mean, scale, sort, ldl, w, x);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------------
// characteristic function by moments (mfs/one_dim/moments.py:309-337): E[exp(i z X)] ~ sum_n w_n exp(i z x_n), the
// post-processing step vmapped over (time steps x z-grid x Monte-Carlo runs) in
// dardel/benes_bernoulli/post_processing_mf.py:37-60.  One CTA per moment vector: one thread derives the quadrature
// (same device code as the filter), all threads sweep the z-grid.  out[b][j] = (re, im) as two doubles (complex128).
// ---------------------------------------------------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(kBlock) characteristic_kernel(int64_t m, const double* __restrict__ ms,
                                                                const double* __restrict__ mean,
                                                                const double* __restrict__ scale,
                                                                const double* __restrict__ zs, double* __restrict__ out) {
  __shared__ double sw[N], sx[N];
  const int64_t b = blockIdx.x;
  if (threadIdx.x == 0) {
    double mm[2 * N], w[N], x[N];
#pragma unroll
    for (int p = 0; p < 2 * N; ++p) mm[p] = ms[b * 2 * N + p];
    const bool ok = moment_quadrature<N>(mm, mean ? mean[b] : 0.0, scale ? scale[b] : 1.0, w, x);
#pragma unroll
    for (int i = 0; i < N; ++i) { sw[i] = ok ? w[i] : nan(""); sx[i] = ok ? x[i] : nan(""); }
  }
  __syncthreads();
  double w[N], x[N];
#pragma unroll
  for (int i = 0; i < N; ++i) { w[i] = sw[i]; x[i] = sx[i]; }
  double2* o = reinterpret_cast<double2*>(out) + b * m;
  for (int64_t j = threadIdx.x; j < m; j += kBlock) {
    const double z = zs[j];
    double re = 0.0, im = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double sn, cs;
      sincos(z * x[i], &sn, &cs);
      re = fma(w[i], cs, re);
      im = fma(w[i], sn, im);
    }
    o[j] = make_double2(re, im);
  }
}

template <int N>
static cudaError_t launch_characteristic(int64_t B, int64_t m, const double* ms, const double* mean, const double* scale,
                                         const double* zs, double* out, cudaStream_t s) {
  characteristic_kernel<N><<<(unsigned)B, kBlock, 0, s>>>(m, ms, mean, scale, zs, out);
  return cudaGetLastError();
}

// The branch-free elementary functions of models.cuh on an array (diagnostic entry point for the parity tests).
__global__ void __launch_bounds__(256) math_selftest_kernel(int64_t n, const double* __restrict__ x, double* __restrict__ out_exp,
                                                            double* __restrict__ out_log, double* __restrict__ out_tanh) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const double v = x[i];
  if (out_exp) out_exp[i] = exp_any(v);
  if (out_log) out_log[i] = log_fast(v);
  if (out_tanh) out_tanh[i] = tanh_fast(v);
}

// ---------------------------------------------------------------------------------------------------------------------
// FP64 FMA peak micro-benchmark: 8 independent accumulator chains per thread, no memory traffic.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fp64_peak_kernel(int iters, double seed, double* sink) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
  }
  const double r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (r == 12345.678) sink[0] = r;  // never true; keeps the chains alive
}

}  // namespace mfs

using namespace mfs;

extern "C" {

int mfs_abi_version(void) { return MFS_ABI_VERSION; }

const char* mfs_last_error(void) { return g_err.c_str(); }

int64_t mfs_launch_count(void) { return g_launches.load(); }

int mfs_functor_lookup(const char* kind, const char* name, int32_t* id) {
  struct Entry { const char* kind; const char* name; int id; };
  static const Entry table[] = {
      {"mode", "raw", MFS_MODE_RAW}, {"mode", "central", MFS_MODE_CENTRAL}, {"mode", "scaled", MFS_MODE_SCALED},
      {"trans", "tme", MFS_TRANS_TME}, {"trans", "tme_normal", MFS_TRANS_TME_NORMAL}, {"trans", "euler", MFS_TRANS_EULER},
      {"trans", "normal_affine", MFS_TRANS_NORMAL_AFFINE},
      {"drift", "benes", MFS_DRIFT_BENES}, {"drift", "well", MFS_DRIFT_WELL}, {"drift", "linear", MFS_DRIFT_LINEAR},
      {"meas", "bernoulli_logistic_cubic", MFS_MEAS_BERNOULLI_LOGISTIC_CUBIC},
      {"meas", "poisson_softplus", MFS_MEAS_POISSON_SOFTPLUS}, {"meas", "gaussian", MFS_MEAS_GAUSSIAN},
  };
  if (!kind || !name || !id) return fail("mfs_functor_lookup: NULL argument");
  for (const Entry& e : table)
    if (!strcmp(e.kind, kind) && !strcmp(e.name, name)) { *id = e.id; return 0; }
  return fail("unknown %s functor '%s'", kind, name);
}

int64_t mfs_filter_1d_workspace_bytes(int32_t N, int64_t B, int64_t T) {
  if (N < 2 || N > MFS_MAX_N || B < 0 || T < 0) return -1;
  const int64_t state = ((int64_t)B * (4 * N + 4) * (int64_t)sizeof(double) + 255) & ~int64_t(255);
  const int64_t idx = ((int64_t)B * 4 + 255) & ~int64_t(255);
  const int64_t counts = (((T + 1) + 2) * 4 + 255) & ~int64_t(255);   // one counter per segment boundary (<= T + 1)
  return state + 2 * idx + counts + 256;
}

int mfs_filter_1d(const mfs_filter1d_args* a, void* stream) {
  if (int rc = validate(a)) return rc;
  return launch_device(*a, static_cast<cudaStream_t>(stream));
}

int mfs_filter_1d_grad(const mfs_filter1d_args* a, int32_t n_tangents, const int32_t* tangent_ids, double* grad_out,
                       void* stream) {
  if (int rc = validate(a)) return rc;
  if (a->mode == MFS_MODE_SCALED) return fail("mfs_filter_1d_grad: raw and central moments only");
  if (a->stable) return fail("mfs_filter_1d_grad: stable=True is not differentiable (eps-substituted pivots); not offered");
  if (a->carry_in || a->carry_out || a->grid_records || a->t_offset)
    return fail("mfs_filter_1d_grad: carry_in / carry_out / t_offset / grid_records are not offered by the gradient kernel");
  if (n_tangents < 1 || n_tangents > MFS_GRAD_MAX_TANGENTS) return fail("mfs_filter_1d_grad: n_tangents=%d outside [1, %d]", n_tangents, MFS_GRAD_MAX_TANGENTS);
  if (!tangent_ids) return fail("mfs_filter_1d_grad: tangent_ids must not be NULL");
  if (a->B == 0) return 0;
  if (!grad_out) return fail("mfs_filter_1d_grad: grad_out must not be NULL");
  GradInfo g = {};
  for (int k = 0; k < 4; ++k) g.tangent_ids[k] = -1;
  for (int k = 0; k < n_tangents; ++k) {
    if (tangent_ids[k] < 0 || tangent_ids[k] >= 2 * MFS_MAX_PARAMS) return fail("mfs_filter_1d_grad: tangent id %d outside [0, %d)", tangent_ids[k], 2 * MFS_MAX_PARAMS);
    g.tangent_ids[k] = tangent_ids[k];
  }
  g.grad_out = grad_out;
  g.n_slots = MFS_GRAD_MAX_TANGENTS;
  if (a->B == 0) return 0;
  cudaError_t e = cudaErrorInvalidValue;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (a->N) {
#define MFS_CASE(n) case n: e = launch_filter1d_grad<n>(*a, g, s); break;
    MFS_CASE(2) MFS_CASE(3) MFS_CASE(4) MFS_CASE(5) MFS_CASE(6) MFS_CASE(7) MFS_CASE(8) MFS_CASE(9)
    MFS_CASE(10) MFS_CASE(11) MFS_CASE(12) MFS_CASE(13) MFS_CASE(14) MFS_CASE(15)
#undef MFS_CASE
    default: break;
  }
  if (e != cudaSuccess) return fail("mfs_filter_1d_grad: kernel launch failed: %s", cudaGetErrorString(e));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int mfs_filter_1d_host(const mfs_filter1d_args* a, int device, int64_t chunk_filters) {
  if (int rc = validate(a)) return rc;
  if (a->B == 0) return 0;
  const int M = 2 * a->N;
  const int64_t T = a->T;
  if (a->carry_in || a->carry_out) return fail("carry_in / carry_out are offered by mfs_filter_1d (device pointers) only");
  if (a->grid_records) return fail("grid_records is offered by mfs_filter_1d (device pointers) only");
  if (a->T > 0 && (a->ys_stride_t != 1 || a->ys_stride_b != T)) return fail("host ys must be contiguous (B, T)");
  if (a->out_mode == MFS_OUT_FULL && (a->ms_stride_t != M || a->ms_stride_b != T * M || a->aux_stride_b != T))
    return fail("host outputs must be contiguous (B, T, 2N) / (B, T)");
  if (a->out_mode == MFS_OUT_MEANVAR && (a->ms_stride_t != 2 || a->ms_stride_b != T * 2))
    return fail("host outputs must be contiguous (B, T, 2)");
  if (a->out_mode == MFS_OUT_LAST && (a->ms_stride_b != M || a->aux_stride_b != 1))
    return fail("host outputs must be contiguous (B, 2N) / (B,)");
  // every per-filter table is copied chunk-wise with its natural row length: check the strides before anything is queued
  if (a->ms0_stride != 0 && a->ms0_stride != M) return fail("host ms0 must be contiguous (B, 2N) or shared (stride 0)");
  if ((a->trans_param_stride != 0 && a->trans_param_stride != MFS_MAX_PARAMS) ||
      (a->meas_param_stride != 0 && a->meas_param_stride != MFS_MAX_PARAMS))
    return fail("host params must be contiguous (B, %d) or shared (stride 0)", MFS_MAX_PARAMS);
  if (a->mean0 && a->mean0_stride != 0 && a->mean0_stride != 1) return fail("host mean0 must be contiguous (B,) or shared (stride 0)");
  if (a->scale0 && a->scale0_stride != 0 && a->scale0_stride != 1) return fail("host scale0 must be contiguous (B,) or shared (stride 0)");
  MFS_CUDA(cudaSetDevice(device));

  const int esz = ys_elem_size(a->ys_dtype);
  const int64_t row_out = a->out_mode == MFS_OUT_FULL ? T * (M + 2) * 8 : a->out_mode == MFS_OUT_MEANVAR ? T * 16 : (M + 4) * 8;
  const int64_t per_filter = row_out + T * esz + (4 * a->N + 12) * 8;
  if (chunk_filters <= 0) {
    // ~1 GiB of staging per chunk; at least two CTAs per SM when that still fits 4 GiB (the D2H copy, not the kernel,
    // is the bottleneck of the full-history mode: small chunks shorten the un-overlapped head and tail of the pipeline)
    chunk_filters = (1LL << 30) / per_filter;
    const int64_t floor_filters = 148LL * 2 * kBlock;
    if (chunk_filters < floor_filters) chunk_filters = ((4LL << 30) / per_filter < floor_filters) ? (4LL << 30) / per_filter : floor_filters;
    if (chunk_filters < kBlock) chunk_filters = kBlock;
  }
  chunk_filters = (chunk_filters + kBlock - 1) / kBlock * kBlock;
  if (chunk_filters > a->B) chunk_filters = a->B;
  const int64_t C = chunk_filters;

  struct Slot {
    cudaStream_t s = nullptr;
    void* ys = nullptr;
    double *ms0 = nullptr, *mean0 = nullptr, *scale0 = nullptr, *tprm = nullptr, *mprm = nullptr;
    double *ms_out = nullptr, *mean_out = nullptr, *scale_out = nullptr, *nell = nullptr;
    int32_t* status = nullptr;
    void* seg_ws = nullptr;
  } slot[2];
  const bool per_ms0 = a->ms0_stride != 0, per_mean0 = a->mean0 && a->mean0_stride != 0,
             per_scale0 = a->scale0 && a->scale0_stride != 0, per_t = a->trans_param_stride != 0,
             per_m = a->meas_param_stride != 0;
  const int64_t out_ms = a->out_mode == MFS_OUT_FULL ? C * T * M : a->out_mode == MFS_OUT_LAST ? C * M
                         : a->out_mode == MFS_OUT_MEANVAR ? C * T * 2 : 0;
  const int64_t out_aux = a->out_mode == MFS_OUT_FULL ? C * T : a->out_mode == MFS_OUT_LAST ? C : 0;

  int rc = 0;
  auto cleanup = [&]() {
    for (Slot& k : slot) {
      if (k.s) cudaStreamSynchronize(k.s);
      ws_free(k.ys); ws_free(k.ms0); ws_free(k.mean0); ws_free(k.scale0); ws_free(k.tprm); ws_free(k.mprm);
      ws_free(k.ms_out); ws_free(k.mean_out); ws_free(k.scale_out); ws_free(k.nell); ws_free(k.status);
      ws_free(k.seg_ws);
      if (k.s) cudaStreamDestroy(k.s);
    }
  };
#define MFS_TRY(call)                                                                          \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) { rc = fail("%s failed: %s", #call, cudaGetErrorString(e__)); cleanup(); return rc; } \
  } while (0)

  // scratch of the segmented execution (live-filter compaction); segment_steps < 0 switches it off
  const int64_t seg_len = a->segment_steps > 0 ? a->segment_steps : kDefaultSegmentSteps;
  const int64_t seg_bytes = (a->segment_steps >= 0 && T >= 2 * seg_len) ? mfs_filter_1d_workspace_bytes(a->N, C, T) : 0;
  const int nslots = a->B > C ? 2 : 1;
  for (int k = 0; k < nslots; ++k) {
    Slot& s = slot[k];
    MFS_TRY(cudaStreamCreateWithFlags(&s.s, cudaStreamNonBlocking));
    if (T > 0) MFS_TRY(ws_alloc(device, (size_t)(C * T * esz), reinterpret_cast<void**>(&s.ys)));
    MFS_TRY(ws_alloc(device, sizeof(double) * (per_ms0 ? C * M : M), reinterpret_cast<void**>(&s.ms0)));
    if (a->mean0) MFS_TRY(ws_alloc(device, sizeof(double) * (per_mean0 ? C : 1), reinterpret_cast<void**>(&s.mean0)));
    if (a->scale0) MFS_TRY(ws_alloc(device, sizeof(double) * (per_scale0 ? C : 1), reinterpret_cast<void**>(&s.scale0)));
    MFS_TRY(ws_alloc(device, sizeof(double) * MFS_MAX_PARAMS * (per_t ? C : 1), reinterpret_cast<void**>(&s.tprm)));
    MFS_TRY(ws_alloc(device, sizeof(double) * MFS_MAX_PARAMS * (per_m ? C : 1), reinterpret_cast<void**>(&s.mprm)));
    if (out_ms) MFS_TRY(ws_alloc(device, sizeof(double) * out_ms, reinterpret_cast<void**>(&s.ms_out)));
    if (out_aux && a->mean_out) MFS_TRY(ws_alloc(device, sizeof(double) * out_aux, reinterpret_cast<void**>(&s.mean_out)));
    if (out_aux && a->scale_out) MFS_TRY(ws_alloc(device, sizeof(double) * out_aux, reinterpret_cast<void**>(&s.scale_out)));
    MFS_TRY(ws_alloc(device, sizeof(double) * C, reinterpret_cast<void**>(&s.nell)));
    if (a->status_out) MFS_TRY(ws_alloc(device, sizeof(int32_t) * C, reinterpret_cast<void**>(&s.status)));
    if (seg_bytes > 0) MFS_TRY(ws_alloc(device, (size_t)seg_bytes, &s.seg_ws));
  }

  int k = 0;
  for (int64_t b0 = 0; b0 < a->B; b0 += C, k ^= (nslots - 1)) {
    Slot& s = slot[k];
    const int64_t n = (a->B - b0 < C) ? a->B - b0 : C;
    const auto H2D = cudaMemcpyHostToDevice;
    const auto D2H = cudaMemcpyDeviceToHost;
    // stream order makes the reuse of this slot's buffers safe: its previous D2H copies precede these H2D copies
    if (T > 0)
      MFS_TRY(cudaMemcpyAsync(s.ys, static_cast<const char*>(a->ys) + b0 * T * esz, (size_t)(n * T * esz), H2D, s.s));
    MFS_TRY(cudaMemcpyAsync(s.ms0, a->ms0 + (per_ms0 ? b0 * a->ms0_stride : 0), sizeof(double) * (per_ms0 ? n * M : M), H2D, s.s));
    if (a->mean0) MFS_TRY(cudaMemcpyAsync(s.mean0, a->mean0 + (per_mean0 ? b0 : 0), sizeof(double) * (per_mean0 ? n : 1), H2D, s.s));
    if (a->scale0) MFS_TRY(cudaMemcpyAsync(s.scale0, a->scale0 + (per_scale0 ? b0 : 0), sizeof(double) * (per_scale0 ? n : 1), H2D, s.s));
    MFS_TRY(cudaMemcpyAsync(s.tprm, a->trans_params + (per_t ? b0 * a->trans_param_stride : 0), sizeof(double) * MFS_MAX_PARAMS * (per_t ? n : 1), H2D, s.s));
    MFS_TRY(cudaMemcpyAsync(s.mprm, a->meas_params + (per_m ? b0 * a->meas_param_stride : 0), sizeof(double) * MFS_MAX_PARAMS * (per_m ? n : 1), H2D, s.s));

    mfs_filter1d_args d = *a;
    d.B = n;
    d.ys = s.ys;
    d.ms0 = s.ms0; d.ms0_stride = per_ms0 ? M : 0;
    d.mean0 = s.mean0; d.mean0_stride = per_mean0 ? 1 : 0;
    d.scale0 = s.scale0; d.scale0_stride = per_scale0 ? 1 : 0;
    d.trans_params = s.tprm; d.trans_param_stride = per_t ? MFS_MAX_PARAMS : 0;
    d.meas_params = s.mprm; d.meas_param_stride = per_m ? MFS_MAX_PARAMS : 0;
    d.ms_out = s.ms_out; d.mean_out = s.mean_out; d.scale_out = s.scale_out;
    d.nell_out = s.nell; d.status_out = s.status;
    d.workspace = s.seg_ws; d.workspace_bytes = seg_bytes; d.segment_steps = (int32_t)(seg_bytes > 0 ? seg_len : 0);
    if (launch_device(d, s.s)) { cleanup(); return -1; }

    if (a->out_mode == MFS_OUT_FULL) {
      MFS_TRY(cudaMemcpyAsync(a->ms_out + b0 * T * M, s.ms_out, sizeof(double) * n * T * M, D2H, s.s));
      if (a->mean_out) MFS_TRY(cudaMemcpyAsync(a->mean_out + b0 * T, s.mean_out, sizeof(double) * n * T, D2H, s.s));
      if (a->scale_out) MFS_TRY(cudaMemcpyAsync(a->scale_out + b0 * T, s.scale_out, sizeof(double) * n * T, D2H, s.s));
    } else if (a->out_mode == MFS_OUT_MEANVAR) {
      MFS_TRY(cudaMemcpyAsync(a->ms_out + b0 * T * 2, s.ms_out, sizeof(double) * n * T * 2, D2H, s.s));
    } else if (a->out_mode == MFS_OUT_LAST) {
      MFS_TRY(cudaMemcpyAsync(a->ms_out + b0 * M, s.ms_out, sizeof(double) * n * M, D2H, s.s));
      if (a->mean_out) MFS_TRY(cudaMemcpyAsync(a->mean_out + b0, s.mean_out, sizeof(double) * n, D2H, s.s));
      if (a->scale_out) MFS_TRY(cudaMemcpyAsync(a->scale_out + b0, s.scale_out, sizeof(double) * n, D2H, s.s));
    }
    MFS_TRY(cudaMemcpyAsync(a->nell_out + b0, s.nell, sizeof(double) * n, D2H, s.s));
    if (a->status_out) MFS_TRY(cudaMemcpyAsync(a->status_out + b0, s.status, sizeof(int32_t) * n, D2H, s.s));
  }
  for (int i = 0; i < nslots; ++i) MFS_TRY(cudaStreamSynchronize(slot[i].s));
  cleanup();
  return 0;
#undef MFS_TRY
}

int mfs_moment_quadrature_1d(int32_t N, int64_t B, const double* ms, const double* mean, const double* scale,
                             int32_t sort_nodes, int32_t ldl, double* weights, double* nodes, void* stream) {
  if (N < 1 || N > MFS_MAX_N) return fail("N=%d outside [1, %d]", N, MFS_MAX_N);
  if (B < 0) return fail("negative B");
  if (B == 0) return 0;
  if (!ms || !weights || !nodes) return fail("NULL pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaErrorInvalidValue;
  switch (N) {
#define MFS_CASE(n) case n: e = launch_quadrature<n>(B, ms, mean, scale, sort_nodes, ldl, weights, nodes, s); break;
    MFS_CASE(1) MFS_CASE(2) MFS_CASE(3) MFS_CASE(4) MFS_CASE(5) MFS_CASE(6) MFS_CASE(7) MFS_CASE(8) MFS_CASE(9)
    MFS_CASE(10) MFS_CASE(11) MFS_CASE(12) MFS_CASE(13) MFS_CASE(14) MFS_CASE(15)
#undef MFS_CASE
  }
  if (e != cudaSuccess) return fail("quadrature launch failed: %s", cudaGetErrorString(e));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int mfs_characteristic_fn_1d(int32_t N, int64_t B, int64_t m, const double* ms, const double* mean, const double* scale,
                             const double* zs, double* out, void* stream) {
  if (N < 1 || N > MFS_MAX_N) return fail("N=%d outside [1, %d]", N, MFS_MAX_N);
  if (B < 0 || m < 0) return fail("negative B or m");
  if (B == 0 || m == 0) return 0;
  if (B > 0x7fffffffLL) return fail("B too large for one launch");
  if (!ms || !zs || !out) return fail("NULL pointer");
  if (reinterpret_cast<uintptr_t>(out) & 15) return fail("out must be 16-byte aligned (complex128)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaErrorInvalidValue;
  switch (N) {
#define MFS_CASE(n) case n: e = launch_characteristic<n>(B, m, ms, mean, scale, zs, out, s); break;
    MFS_CASE(1) MFS_CASE(2) MFS_CASE(3) MFS_CASE(4) MFS_CASE(5) MFS_CASE(6) MFS_CASE(7) MFS_CASE(8) MFS_CASE(9)
    MFS_CASE(10) MFS_CASE(11) MFS_CASE(12) MFS_CASE(13) MFS_CASE(14) MFS_CASE(15)
#undef MFS_CASE
  }
  if (e != cudaSuccess) return fail("characteristic_fn launch failed: %s", cudaGetErrorString(e));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int mfs_filter_nd(const mfs_filternd_args* a, void* stream) {
  if (!a) return fail("args is NULL");
  if (a->abi_version != MFS_ABI_VERSION) return fail("abi_version %d != %d", a->abi_version, MFS_ABI_VERSION);
  if (a->d != 2) return fail("only d = 2 is implemented (got d = %d)", a->d);
  if (a->N < 2 || a->N > 7) return fail("N=%d outside [2, 7] for the 2-D filter (one basis row per lane: N(N+1)/2 <= 32)", a->N);
  if (a->mode != MFS_MODE_RAW && a->mode != MFS_MODE_CENTRAL) return fail("2-D filter: mode must be raw or central");
  if (a->trans_id != MFS_TRANS_EULER && a->trans_id != MFS_TRANS_TME_NORMAL && a->trans_id != MFS_TRANS_TME)
    return fail("2-D filter: transition must be euler, tme_normal or tme (Lotka--Volterra)");
  if (a->trans_id != MFS_TRANS_EULER && (a->tme_order < 1 || a->tme_order > 2)) return fail("2-D filter: tme_order must be 1 or 2");
  if (a->meas_id != MFS_MEAS_BERNOULLI_LOGISTIC_CUBIC) return fail("2-D filter: measurement must be bernoulli_logistic_cubic");
  if (a->obs_dim < 0 || a->obs_dim > 1) return fail("obs_dim must be 0 or 1");
  if (a->B < 0 || a->T < 0) return fail("negative B or T");
  if (a->B == 0) return 0;
  if (!a->trans_params || !a->meas_params || !a->ms0 || !a->inds || !a->nell_out || (a->T > 0 && !a->ys))
    return fail("NULL pointer argument");
  if (a->out_mode < MFS_OUT_FULL || a->out_mode > MFS_OUT_MEANVAR) return fail("unknown out_mode %d", a->out_mode);
  const bool nd_aux = a->out_mode == MFS_OUT_FULL || a->out_mode == MFS_OUT_LAST;    // MEANVAR carries the means in ms_out
  if (a->mode == MFS_MODE_CENTRAL && (!a->mean0 || (nd_aux && !a->mean_out))) return fail("mean0 / mean_out is NULL in central mode");
  if (a->out_mode != MFS_OUT_NONE && !a->ms_out && !(a->out_mode == MFS_OUT_MEANVAR && a->T == 0)) return fail("ms_out is NULL");
  NdArgs k;
  k.mode = a->mode; k.trans_id = a->trans_id; k.tme_order = a->tme_order; k.meas_id = a->meas_id; k.obs_dim = a->obs_dim;
  k.out_mode = a->out_mode; k.stable = a->stable ? 1 : 0; k.B = a->B; k.T = a->T; k.dt = a->dt;
  k.trans_params = a->trans_params; k.trans_param_stride = a->trans_param_stride;
  k.meas_params = a->meas_params; k.meas_param_stride = a->meas_param_stride;
  k.ms0 = a->ms0; k.ms0_stride = a->ms0_stride; k.mean0 = a->mean0; k.mean0_stride = a->mean0_stride;
  k.ys = a->ys; k.inds = a->inds; k.pos = nullptr;
  k.ms_out = a->ms_out; k.mean_out = a->mean_out; k.nell_out = a->nell_out; k.status_out = a->status_out;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaErrorInvalidValue;
  switch (a->N) {
    case 2: e = launch_filter_nd<2>(k, s); break;
    case 3: e = launch_filter_nd<3>(k, s); break;
    case 4: e = launch_filter_nd<4>(k, s); break;
    case 5: e = launch_filter_nd<5>(k, s); break;
    case 6: e = launch_filter_nd<6>(k, s); break;
    case 7: e = launch_filter_nd<7>(k, s); break;
  }
  if (e != cudaSuccess) return fail("2-D filter launch failed: %s", cudaGetErrorString(e));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int mfs_moment_quadrature_nd(int32_t N, int32_t d, int64_t B, const double* ms, const double* mean, const double* scale,
                             const int32_t* inds, int32_t ldl, double* weights, double* nodes, void* stream) {
  if (d != 2) return fail("only d = 2 is implemented (got d = %d)", d);
  if (N < 2 || N > 7) return fail("N=%d outside [2, 7] for the 2-D quadrature", N);
  if (B < 0) return fail("negative B");
  if (B == 0) return 0;
  if (!ms || !inds || !weights || !nodes) return fail("NULL pointer argument");
  NdQuadArgs q;
  q.B = B; q.ms = ms; q.mean = mean; q.scale = scale; q.inds = inds; q.stable = ldl ? 1 : 0; q.weights = weights; q.nodes = nodes;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaErrorInvalidValue;
  switch (N) {
    case 2: e = launch_quadrature_nd<2>(q, s); break;
    case 3: e = launch_quadrature_nd<3>(q, s); break;
    case 4: e = launch_quadrature_nd<4>(q, s); break;
    case 5: e = launch_quadrature_nd<5>(q, s); break;
    case 6: e = launch_quadrature_nd<6>(q, s); break;
    case 7: e = launch_quadrature_nd<7>(q, s); break;
  }
  if (e != cudaSuccess) return fail("2-D quadrature launch failed: %s", cudaGetErrorString(e));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int mfs_math_selftest(int64_t n, const double* x, double* out_exp, double* out_log, double* out_tanh, void* stream) {
  if (n < 0) return fail("negative n");
  if (n == 0) return 0;
  if (!x) return fail("x is NULL");
  math_selftest_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(n, x, out_exp, out_log, out_tanh);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("math selftest launch failed: %s", cudaGetErrorString(e));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int mfs_release_cached_memory(void) {
  std::lock_guard<std::mutex> lock(g_ws_mutex);
  for (size_t i = 0; i < g_ws.size();) {
    if (!g_ws[i].busy) { cudaFree(g_ws[i].ptr); g_ws.erase(g_ws.begin() + i); } else ++i;
  }
  return 0;
}

int mfs_fp64_peak(int device, int32_t iters, double* flops, double* ms) {
  if (!flops) return fail("flops is NULL");
  MFS_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  MFS_CUDA(cudaGetDeviceProperties(&prop, device));
  double* sink = nullptr;
  MFS_CUDA(cudaMalloc(&sink, sizeof(double)));
  const int grid = prop.multiProcessorCount * 8;
  cudaEvent_t e0, e1;
  MFS_CUDA(cudaEventCreate(&e0));
  MFS_CUDA(cudaEventCreate(&e1));
  fp64_peak_kernel<<<grid, 256>>>(iters / 4 + 1, 1.0, sink);  // warm-up
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    MFS_CUDA(cudaEventRecord(e0));
    fp64_peak_kernel<<<grid, 256>>>(iters, 1.0, sink);
    MFS_CUDA(cudaEventRecord(e1));
    MFS_CUDA(cudaEventSynchronize(e1));
    float t;
    MFS_CUDA(cudaEventElapsedTime(&t, e0, e1));
    if (t < best) best = t;
  }
  g_launches.fetch_add(6, std::memory_order_relaxed);
  MFS_CUDA(cudaGetLastError());
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  const double fmas = (double)grid * 256.0 * (double)iters * 16.0 * 8.0;
  *flops = 2.0 * fmas / (best * 1e-3);
  if (ms) *ms = best;
  return 0;
}

}  // extern "C"
