// Dev tool: does a non-FP64 instruction cost FP64 throughput on sm_100a?  Two warps per SM sub-partition (the step kernels'
// occupancy), four independent DFMA chains per warp (enough ILP for the 2.17-cycle issue floor), plus M integer ALU
// instructions (independent 32-bit add/xor chains, via asm volatile so none is folded) per 4 DFMAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/fp64_mix tools/fp64_mix.cu && build/fp64_mix
#include <cstdio>
#include <cuda_runtime.h>

template <int M, int KIND>
__global__ void mix(double* sink, int iters, double a, double b, long long* cycles) {
  double x[4];
  unsigned n[4];
  float f[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) { x[k] = threadIdx.x * 1e-3 + k; n[k] = threadIdx.x + k; f[k] = threadIdx.x + k; }
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
#pragma unroll
      for (int k = 0; k < 4; ++k) x[k] = fma(x[k], a, b);
#pragma unroll
      for (int m = 0; m < M; ++m) {
        if (KIND == 0) asm volatile("add.u32 %0, %0, %1;" : "+r"(n[m & 3]) : "r"(i));          // IADD3 (ALU)
        else if (KIND == 1) asm volatile("mov.b32 %0, %1;" : "=r"(n[m & 3]) : "r"(n[(m + 1) & 3]));   // MOV
        else asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[m & 3]) : "f"(1.0001f));       // FFMA (FMA pipe)
      }
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) s += x[k] + n[k] + f[k];
  if (s == 123.456) sink[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int M, int KIND>
void run() {
  double* sink; long long* cyc;
  cudaMalloc(&sink, 8); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  mix<M, KIND><<<148, 256>>>(sink, 16, 1.0000001, 1e-9, cyc);
  cudaDeviceSynchronize();
  mix<M, KIND><<<148, 256>>>(sink, iters, 1.0000001, 1e-9, cyc);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double per_dfma_smsp = (double)c / (iters * 16.0 * 4.0) / 2.0;     // two warps per sub-partition
  const char* kind[] = {"IADD", "MOV", "FFMA"};
  printf("2 warps/SMSP, 4 DFMA chains + %d %s per 4 DFMA: %.3f cycles per DFMA per SMSP, %.3f cycles per instruction per SMSP\n",
         M, kind[KIND], per_dfma_smsp, per_dfma_smsp * 4.0 / (4.0 + M));
  cudaFree(sink); cudaFree(cyc);
}

int main() {
  run<0, 0>(); run<1, 0>(); run<2, 0>(); run<4, 0>(); run<8, 0>();
  run<1, 1>(); run<2, 1>(); run<4, 1>();
  run<1, 2>(); run<2, 2>(); run<4, 2>();
  return 0;
}
