// microbench7.cu — what a global store ahead of an mbarrier arrive costs the issuing warp (sm_100a).
// One warp runs a dependent FP64 chain per iteration and, per variant, adds shared/global stores, an mbarrier
// arrive (release.cta) and a test_wait probe + shared loads, the way the wavefront kernel's chain warp does.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false scripts/microbench7.cu -o scripts/microbench7.bin
#include <cstdio>
#include <cuda_runtime.h>
#include "../sgdnet_b200/csrc/common.cuh"
using namespace sgd;

#define N_IT 4000

template <int V>
__global__ void k(double* out, long long* cyc, double* gbuf, const uint32_t* idx, double seed) {
  __shared__ uint64_t bar[32];
  __shared__ double sv[32 * 4];
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 32; ++i) mbar_init(&bar[i], 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 128; i += 32) sv[i] = 1.0 + i * 1e-9;
  __syncthreads();
  const uint32_t sb = smem_u32(sv), bb = smem_u32(bar);
  double x = seed, acc = 0.0;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N_IT; ++i) {
    const uint32_t o8 = (i & 31) * 8u;
    double a = 0.0, b2 = 0.0, c = 0.0;
    if (V >= 4) {   // probe + dependent loads for the "next row"
      const bool rdy = mbar_test_wait_a(bb + ((i + 1) & 31) * 8u, ((i + 1) >> 5) & 1u);
      const uint32_t od = rdy ? o8 : o8 + (uint32_t)(seed == 12345.0);
      a = lds_f64(sb + od); b2 = lds_f64(sb + 256 + od); c = lds_f64(sb + 512 + od);
    }
#pragma unroll
    for (int j = 0; j < 12; ++j) x = fma(x, 0.9999999, 1e-9);   // ~12 dependent DFMA
    if (V == 1 || V >= 3) sts_f64(sb + 768 + o8, x);
    if (V == 2 || V >= 3) gbuf[idx[i & 1023]] = x;                // global store, scattered addresses
    if (V == 1 || V >= 3) mbar_arrive_if(lane == 0, bb + o8);
    if (V == 5) {   // store issued AFTER the arrive
    }
    acc += a + b2 + c;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { cyc[V] = t1 - t0; out[V] = x + acc; }
}

int main() {
  double *out, *gbuf; long long* cyc; uint32_t* idx;
  cudaMalloc(&out, 64 * 8); cudaMalloc(&cyc, 64 * 8); cudaMalloc(&gbuf, 8 << 20); cudaMalloc(&idx, 4096);
  uint32_t h_idx[1024];
  for (int i = 0; i < 1024; ++i) h_idx[i] = (uint32_t)((i * 2654435761u) % (1u << 20));
  cudaMemcpy(idx, h_idx, sizeof(h_idx), cudaMemcpyHostToDevice);
  long long h[16];
  for (int rep = 0; rep < 2; ++rep) {
    k<0><<<1, 32>>>(out, cyc, gbuf, idx, 0.5);
    k<1><<<1, 32>>>(out, cyc, gbuf, idx, 0.5);
    k<2><<<1, 32>>>(out, cyc, gbuf, idx, 0.5);
    k<3><<<1, 32>>>(out, cyc, gbuf, idx, 0.5);
    k<4><<<1, 32>>>(out, cyc, gbuf, idx, 0.5);
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  }
  const char* names[] = {"12 dependent DFMA", "+ sts + arrive", "+ global store (no arrive)", "+ sts + global store + arrive",
                         "+ probe + 3 lds + sts + global store + arrive"};
  for (int i = 0; i < 5; ++i) printf("%-48s %8.1f cycles/iter\n", names[i], double(h[i]) / N_IT);
  printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
