// Dev tool: FP64 FMA dependent-issue latency and throughput vs (warps per SM sub-partition, ILP) on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/fp64_latency tools/fp64_latency.cu && build/fp64_latency
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void chain(double* sink, int iters, double a, double b, long long* cycles) {
  double x[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) x[k] = threadIdx.x * 1e-3 + k;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
#pragma unroll
      for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], a, b);
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += x[k];
  if (s == 123.456) sink[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int ILP>
void run(int warps_per_sm) {
  double* sink; long long* cyc;
  cudaMalloc(&sink, 8); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  // one CTA per SM with warps_per_sm warps (spread over the 4 sub-partitions round robin)
  chain<ILP><<<148, 32 * warps_per_sm>>>(sink, 16, 1.0000001, 1e-9, cyc);
  cudaDeviceSynchronize();
  chain<ILP><<<148, 32 * warps_per_sm>>>(sink, iters, 1.0000001, 1e-9, cyc);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double per_fma_warp = (double)c / (iters * 16.0 * ILP);          // cycles per DFMA issued by one warp
  const double per_smsp = per_fma_warp / (warps_per_sm / 4.0);             // cycles per DFMA per sub-partition
  printf("warps/SMSP %4.1f ILP %d : %.2f cycles per dependent step, %.2f cycles/DFMA per warp, %.2f cycles/DFMA per SMSP (pipe floor 2.0)\n",
         warps_per_sm / 4.0, ILP, (double)c / (iters * 16.0), per_fma_warp, per_smsp);
  cudaFree(sink); cudaFree(cyc);
}

int main() {
  for (int w : {4, 8, 12, 16}) {
    run<1>(w); run<2>(w); run<3>(w); run<4>(w); run<6>(w); run<8>(w);
  }
  return 0;
}
