// Instantiations of the 2-D moment filter kernel (one translation unit; N = 2..6).
#include "filter_nd.cuh"

namespace mfs {

template <int N>
cudaError_t launch_filter_nd(const NdArgs& a, cudaStream_t stream) {
  using D = NdDims<N>;
  const size_t smem = sizeof(double) * kNdWarps * D::kDoubles + sizeof(int) * D::kTabInts;
  static bool configured = false;   // benign race: the attribute is idempotent
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(filter_nd_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const unsigned grid = (unsigned)((a.B + kNdWarps - 1) / kNdWarps);
  filter_nd_kernel<N><<<grid, kNdWarps * 32, smem, stream>>>(a);
  return cudaGetLastError();
}

template <int N>
cudaError_t launch_quadrature_nd(const NdQuadArgs& a, cudaStream_t stream) {
  using D = NdDims<N>;
  const size_t smem = sizeof(double) * kNdWarps * D::kDoubles + sizeof(int) * D::kTabInts;
  static bool configured = false;   // benign race: the attribute is idempotent
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(quadrature_nd_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const unsigned grid = (unsigned)((a.B + kNdWarps - 1) / kNdWarps);
  quadrature_nd_kernel<N><<<grid, kNdWarps * 32, smem, stream>>>(a);
  return cudaGetLastError();
}

template cudaError_t launch_quadrature_nd<2>(const NdQuadArgs&, cudaStream_t);
template cudaError_t launch_quadrature_nd<3>(const NdQuadArgs&, cudaStream_t);
template cudaError_t launch_quadrature_nd<4>(const NdQuadArgs&, cudaStream_t);
template cudaError_t launch_quadrature_nd<5>(const NdQuadArgs&, cudaStream_t);
template cudaError_t launch_quadrature_nd<6>(const NdQuadArgs&, cudaStream_t);

template cudaError_t launch_filter_nd<2>(const NdArgs&, cudaStream_t);
template cudaError_t launch_filter_nd<3>(const NdArgs&, cudaStream_t);
template cudaError_t launch_filter_nd<4>(const NdArgs&, cudaStream_t);
template cudaError_t launch_filter_nd<5>(const NdArgs&, cudaStream_t);
template cudaError_t launch_filter_nd<6>(const NdArgs&, cudaStream_t);

}  // namespace mfs
