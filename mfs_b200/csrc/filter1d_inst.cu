// One translation unit per quadrature order N (compiled with -DMFS_N=<N>) so the orders build in parallel.
#include "filter1d.cuh"

#ifndef MFS_N
#error "compile with -DMFS_N=<quadrature order>"
#endif

namespace mfs {

template <int N, int MODE, int KIND, int MEAS>
static cudaError_t launch_one(const mfs_filter1d_args& a, const SegInfo& g, cudaStream_t stream) {
  const unsigned grid = (unsigned)((a.B + kBlock - 1) / kBlock);
  constexpr size_t smem = sizeof(double) * smem_rows<N, MODE, KIND>() * kBlock;
  if (smem > 48 * 1024) {
    // the opt-in is per device (and per context): track it per device ordinal, not per process
    static bool configured[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !configured[dev]) {
      e = cudaFuncSetAttribute(filter1d_kernel<N, MODE, KIND, MEAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      if (dev >= 0 && dev < 64) configured[dev] = true;   // benign race: the attribute is idempotent
    }
  }
  filter1d_kernel<N, MODE, KIND, MEAS><<<grid, kBlock, smem, stream>>>(a, g);
  return cudaGetLastError();
}

template <int N, int MODE>
static cudaError_t launch_kind(const mfs_filter1d_args& a, const SegInfo& g, int kind, int meas_ct, cudaStream_t stream) {
  switch (kind) {
    case KIND_TME: return launch_one<N, MODE, KIND_TME, kMeasRuntime>(a, g, stream);
    case KIND_NORMAL:
      if (meas_ct == MFS_MEAS_POISSON_SOFTPLUS) return launch_one<N, MODE, KIND_NORMAL, MFS_MEAS_POISSON_SOFTPLUS>(a, g, stream);
      return launch_one<N, MODE, KIND_NORMAL, kMeasRuntime>(a, g, stream);
    case KIND_BENES_TME: return launch_one<N, MODE, KIND_BENES_TME, MFS_MEAS_BERNOULLI_LOGISTIC_CUBIC>(a, g, stream);
    default: return cudaErrorInvalidValue;
  }
}

template <>
cudaError_t launch_filter1d<MFS_N>(const mfs_filter1d_args& a, const SegInfo& g, int kind, int meas_ct, cudaStream_t stream) {
  switch (a.mode) {
    case MFS_MODE_RAW: return launch_kind<MFS_N, MFS_MODE_RAW>(a, g, kind, meas_ct, stream);
    case MFS_MODE_CENTRAL: return launch_kind<MFS_N, MFS_MODE_CENTRAL>(a, g, kind, meas_ct, stream);
    case MFS_MODE_SCALED: return launch_kind<MFS_N, MFS_MODE_SCALED>(a, g, kind, meas_ct, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace mfs
