// FP64 pipe micro-benchmark (B200): dependent-chain latency of DFMA and throughput as a function of ILP and resident
// warps per scheduler.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_lat fp64_lat.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void chain(int iters, double* out, long long* cyc) {
  double a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
  const double m = 1.0000001, c = 1e-9;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], m, c);
  }
  const long long t1 = clock64();
  double r = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) r += a[i];
  if (r == 1234.5) out[0] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int ILP>
void run(int warps) {
  double* out; long long* cyc; long long h;
  cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  chain<ILP><<<1, warps * 32>>>(iters, out, cyc);
  chain<ILP><<<1, warps * 32>>>(iters, out, cyc);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double per = (double)h / (iters * 8.0);
  printf("ILP %d warps/SM %2d (per scheduler %d): %.2f cyc per chain step, %.2f DFMA warp-instr/cyc/SM\n", ILP, warps,
         (warps + 3) / 4, per, ILP * warps / per);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {1, 4, 8, 16, 32}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
  return 0;
}
