// One translation unit per quadrature order N (compiled with -DMFS_N=<N>) so the orders build in parallel.
#include "filter1d.cuh"

#ifndef MFS_N
#error "compile with -DMFS_N=<quadrature order>"
#endif

namespace mfs {

template <int N, int MODE, int KIND>
static cudaError_t launch_one(const mfs_filter1d_args& a, const SegInfo& g, cudaStream_t stream) {
  const unsigned grid = (unsigned)((a.B + kBlock - 1) / kBlock);
#ifdef MFS_QL_SMEM
  const size_t smem = sizeof(double) * 3 * N * kBlock;
  static bool configured = false;   // benign race: the attribute is idempotent
  if (!configured && smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(filter1d_kernel<N, MODE, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  filter1d_kernel<N, MODE, KIND><<<grid, kBlock, smem, stream>>>(a, g);
#else
  filter1d_kernel<N, MODE, KIND><<<grid, kBlock, 0, stream>>>(a, g);
#endif
  return cudaGetLastError();
}

template <int N, int MODE>
static cudaError_t launch_kind(const mfs_filter1d_args& a, const SegInfo& g, int kind, cudaStream_t stream) {
  switch (kind) {
    case KIND_TME: return launch_one<N, MODE, KIND_TME>(a, g, stream);
    case KIND_NORMAL:
      // the reference's scaled Normal factories divide every order by prod(scale**k) (moments.py:205,243): not offered
      if constexpr (MODE == MFS_MODE_SCALED) return cudaErrorInvalidValue;
      else return launch_one<N, MODE, KIND_NORMAL>(a, g, stream);
    case KIND_BENES_TME: return launch_one<N, MODE, KIND_BENES_TME>(a, g, stream);
    default: return cudaErrorInvalidValue;
  }
}

template <>
cudaError_t launch_filter1d<MFS_N>(const mfs_filter1d_args& a, const SegInfo& g, int kind, cudaStream_t stream) {
  switch (a.mode) {
    case MFS_MODE_RAW: return launch_kind<MFS_N, MFS_MODE_RAW>(a, g, kind, stream);
    case MFS_MODE_CENTRAL: return launch_kind<MFS_N, MFS_MODE_CENTRAL>(a, g, kind, stream);
    case MFS_MODE_SCALED: return launch_kind<MFS_N, MFS_MODE_SCALED>(a, g, kind, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace mfs
