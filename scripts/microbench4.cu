// microbench4.cu — FP64 pipe throughput per SM sub-partition on sm_100a: independent DFMA streams from 1..16 warps.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, int ilp_mode) {
  double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double y = 1.0000001, z = 1e-9;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 1000; ++i) {
    a0 = fma(a0, y, z); a1 = fma(a1, y, z); a2 = fma(a2, y, z); a3 = fma(a3, y, z);
    a4 = fma(a4, y, z); a5 = fma(a5, y, z); a6 = fma(a6, y, z); a7 = fma(a7, y, z);
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 2048 * 8); cudaMalloc(&cyc, 8);
  for (int nw : {1, 2, 4, 8, 16, 32}) {
    long long h;
    for (int r = 0; r < 2; ++r) k<<<1, nw * 32>>>(out, cyc, 0);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("warps=%2d: %lld cycles for 8000 DFMA per warp -> %.2f cycles per warp-DFMA per SM, %.1f lanes/clk/SM\n", nw, h, double(h) / (8000.0 * nw), 32.0 * 8000 * nw / h);
  }
  return 0;
}
