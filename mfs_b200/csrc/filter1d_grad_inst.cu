// Gradient kernel, one translation unit per quadrature order N (compiled with -DMFS_N=<N>).
#include "filter1d_grad.cuh"

#ifndef MFS_N
#error "compile with -DMFS_N=<quadrature order>"
#endif

namespace mfs {

template <int N>
cudaError_t launch_filter1d_grad(const mfs_filter1d_args& a, const GradInfo& g, cudaStream_t stream);

template <>
cudaError_t launch_filter1d_grad<MFS_N>(const mfs_filter1d_args& a, const GradInfo& g, cudaStream_t stream) {
  const unsigned grid = (unsigned)((a.B + MFS_GRAD_BLOCK - 1) / MFS_GRAD_BLOCK);
  filter1d_grad_kernel<MFS_N, 2><<<grid, MFS_GRAD_BLOCK, 0, stream>>>(a, g);
  return cudaGetLastError();
}

}  // namespace mfs
