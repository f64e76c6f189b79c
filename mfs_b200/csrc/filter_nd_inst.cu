// Instantiations of the 2-D moment filter kernel (one translation unit; N = 2..7: S = N(N+1)/2 <= 28 basis functions, one
// row per lane; the reference runs N = 5 and N = 7, reproduce_paper_plots/plot_prey_predator_errs.py:10).
#include "filter_nd.cuh"

namespace mfs {

// The dynamic-shared-memory opt-in is per device (and per context): tracked per device ordinal, not per process.
template <typename Kernel>
static cudaError_t opt_in_smem(Kernel kernel, bool (&configured)[64], size_t smem) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64 && configured[dev]) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess && dev >= 0 && dev < 64) configured[dev] = true;   // benign race: the attribute is idempotent
  return e;
}

template <int N>
cudaError_t launch_filter_nd(const NdArgs& a, cudaStream_t stream) {
  using D = NdDims<N>;
  constexpr int kNdWarps = nd_warps<N>();
  const size_t smem = sizeof(double) * kNdWarps * D::kDoubles + D::kTabBytes;
  static bool configured[64] = {};
  if (cudaError_t e = opt_in_smem(filter_nd_kernel<N>, configured, smem)) return e;
  const unsigned grid = (unsigned)((a.B + kNdWarps - 1) / kNdWarps);
  filter_nd_kernel<N><<<grid, kNdWarps * 32, smem, stream>>>(a);
  return cudaGetLastError();
}

template <int N>
cudaError_t launch_quadrature_nd(const NdQuadArgs& a, cudaStream_t stream) {
  using D = NdDims<N>;
  constexpr int kNdWarps = nd_warps<N>();
  const size_t smem = sizeof(double) * kNdWarps * D::kDoubles + D::kTabBytes;
  static bool configured[64] = {};
  if (cudaError_t e = opt_in_smem(quadrature_nd_kernel<N>, configured, smem)) return e;
  const unsigned grid = (unsigned)((a.B + kNdWarps - 1) / kNdWarps);
  quadrature_nd_kernel<N><<<grid, kNdWarps * 32, smem, stream>>>(a);
  return cudaGetLastError();
}

template cudaError_t launch_quadrature_nd<2>(const NdQuadArgs&, cudaStream_t);
template cudaError_t launch_quadrature_nd<3>(const NdQuadArgs&, cudaStream_t);
template cudaError_t launch_quadrature_nd<4>(const NdQuadArgs&, cudaStream_t);
template cudaError_t launch_quadrature_nd<5>(const NdQuadArgs&, cudaStream_t);
template cudaError_t launch_quadrature_nd<6>(const NdQuadArgs&, cudaStream_t);
template cudaError_t launch_quadrature_nd<7>(const NdQuadArgs&, cudaStream_t);

template cudaError_t launch_filter_nd<2>(const NdArgs&, cudaStream_t);
template cudaError_t launch_filter_nd<3>(const NdArgs&, cudaStream_t);
template cudaError_t launch_filter_nd<4>(const NdArgs&, cudaStream_t);
template cudaError_t launch_filter_nd<5>(const NdArgs&, cudaStream_t);
template cudaError_t launch_filter_nd<6>(const NdArgs&, cudaStream_t);
template cudaError_t launch_filter_nd<7>(const NdArgs&, cudaStream_t);

}  // namespace mfs
