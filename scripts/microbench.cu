// microbench.cu — latency measurements that size the sparse wavefront kernel's serial chain on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -I. scripts/microbench.cu -o build/microbench
#include <cstdio>
#include <cuda_runtime.h>
#include "../sgdnet_b200/csrc/common.cuh"
using namespace sgd;

#define N_IT 2000

__global__ void k_chain(double* out, long long* cyc, double seed) {
  double x = seed, y = 1.0000001, z = 0.9999999;
  long long t0, t1;
  // DFMA
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N_IT; ++i) x = fma(x, y, z);
  t1 = clock64(); cyc[0] = t1 - t0; out[0] = x;
  // DADD
  x = seed; t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N_IT; ++i) x = x + z;
  t1 = clock64(); cyc[1] = t1 - t0; out[1] = x;
  // DMUL
  x = seed; t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N_IT; ++i) x = x * y;
  t1 = clock64(); cyc[2] = t1 - t0; out[2] = x;
  // DDIV
  x = seed; t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N_IT; ++i) x = y / x;
  t1 = clock64(); cyc[3] = t1 - t0; out[3] = x;
  // sgd_exp
  x = seed; t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N_IT; ++i) x = sgd_exp(x) * 0.3;
  t1 = clock64(); cyc[4] = t1 - t0; out[4] = x;
  // full binomial intercept chain (token section arithmetic)
  double b = 0.1, gsi = 0.01; const double nd = 1e6, gamma = 0.3, dot = seed, gm = 0.2, yv = 1.0;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N_IT; ++i) {
    const double lp = dot + b;
    const double g = gradient_scalar(kBinomial, lp, yv);
    const double gch = g - gm;
    gsi += gch / nd;
    b -= gamma * (gsi * 0.01 + gch / nd);
  }
  t1 = clock64(); cyc[5] = t1 - t0; out[5] = b + gsi;
  // warp_sum of doubles
  x = seed + threadIdx.x; t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N_IT; ++i) x = warp_sum(x) * 0.03125;
  t1 = clock64(); cyc[6] = t1 - t0; out[6] = x;
  // gaussian chain
  b = 0.1; gsi = 0.01;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N_IT; ++i) {
    const double lp = dot + b;
    const double g = lp - yv;
    const double gch = g - gm;
    gsi += gch / nd;
    b -= gamma * (gsi * 0.01 + gch / nd);
  }
  t1 = clock64(); cyc[7] = t1 - t0; out[7] = b + gsi;
  // reciprocal-multiply exact division candidate: q = a*rn; r = fma(-q, nd, a); q' = fma(r, rn, q)
  x = seed; const double rn = 1.0 / nd; t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N_IT; ++i) { double q = x * rn; double rr = fma(-q, nd, x); x = fma(rr, rn, q) + 1.0; }
  t1 = clock64(); cyc[8] = t1 - t0; out[8] = x;
}

// token ping-pong between W warps through mbarriers (ring of 32, like the kernel) or through polled shared words
__global__ void k_handoff(long long* cyc, int mode, int W) {
  __shared__ uint64_t bar[32];
  __shared__ double tokv[32];
  __shared__ volatile unsigned long long flag[32 * 2];   // mode 1: {value bits, seq} pairs
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 32; ++i) { mbar_init(&bar[i], 1); flag[2 * i] = 0; flag[2 * i + 1] = 0; }
    fence_barrier_init();
    tokv[0] = 1.0; mbar_arrive(&bar[0]);
    flag[0] = __double_as_longlong(1.0); flag[1] = 1;   // seq q+1
  }
  __syncthreads();
  const int rows = 20000;
  long long t0 = clock64();
  for (int q = warp; q < rows; q += W) {
    const int sq = q & 31;
    double v;
    if (mode == 0) {
      mbar_wait(&bar[sq], (q >> 5) & 1);
      v = tokv[sq];
    } else {
      while (flag[2 * sq + 1] != (unsigned long long)(q + 1)) {}
      v = __longlong_as_double(flag[2 * sq]);
    }
    v = v * 1.0000001;
    if (lane == 0) {
      const int s1 = (sq + 1) & 31;
      if (mode == 0) { tokv[s1] = v; mbar_arrive(&bar[s1]); }
      else { flag[2 * s1] = __double_as_longlong(v); __threadfence_block(); flag[2 * s1 + 1] = (unsigned long long)(q + 2); }
    }
    __syncwarp();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = (t1 - t0) / rows;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 64 * 8); cudaMalloc(&cyc, 64 * 8);
  long long h[16];
  for (int rep = 0; rep < 2; ++rep) {
    k_chain<<<1, 32>>>(out, cyc, 0.5);
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  }
  const char* names[] = {"DFMA", "DADD", "DMUL", "DDIV", "sgd_exp(+mul)", "binomial intercept chain", "warp_sum(double)+mul", "gaussian intercept chain", "recip-div(3 fma)+add"};
  for (int i = 0; i < 9; ++i) printf("%-28s %8.1f cycles/iter\n", names[i], double(h[i]) / N_IT);
  for (int mode = 0; mode < 2; ++mode)
    for (int W : {2, 4, 8}) {
      k_handoff<<<1, W * 32>>>(cyc, mode, W);
      cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
      printf("handoff %-10s W=%d %6lld cycles/row\n", mode == 0 ? "mbarrier" : "smem-poll", W, h[0]);
    }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return 0;
}
