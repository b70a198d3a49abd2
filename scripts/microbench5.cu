// microbench5.cu — how much do warps that wait on an mbarrier (try_wait loop, lane-0-only try_wait loop,
// test_wait + nanosleep) slow down a warp doing dependent FP64 work on the same SM?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../sgdnet_b200/csrc/common.cuh"
using namespace sgd;

__global__ void k(long long* out, double* sink, int mode, int nwait) {
  __shared__ uint64_t bar;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); stop = 0; }
  __syncthreads();
  if (warp == 0) {
    double x = 0.5 + lane;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 4000; ++i) x = fma(x, 1.0000001, 1e-9);
    long long t1 = clock64();
    if (lane == 0) { out[0] = t1 - t0; stop = 1; mbar_arrive(&bar); }
    sink[lane] = x;
  } else if (warp <= nwait) {
    if (mode == 0) { while (!mbar_try_wait(&bar, 0)) {} }
    else if (mode == 1) { if (lane == 0) { while (!mbar_try_wait(&bar, 0)) {} } __syncwarp(); }
    else if (mode == 2) { while (!mbar_test_wait(&bar, 0)) __nanosleep(100); }
    else if (mode == 3) { if (lane < 7) { while (!mbar_try_wait(&bar, 0)) {} } __syncwarp(); }
    else { while (!stop) {} }
  }
}
int main() {
  long long* d; double* sink; cudaMalloc(&d, 8); cudaMalloc(&sink, 4096);
  const char* names[] = {"try_wait (all lanes)", "try_wait (lane 0)", "test_wait+nanosleep(100)", "try_wait (7 lanes)", "volatile smem poll"};
  for (int mode = 0; mode < 5; ++mode)
    for (int nwait : {0, 1, 3, 4, 7, 11}) {
      long long h;
      for (int r = 0; r < 2; ++r) k<<<1, 12 * 32>>>(d, sink, mode, nwait);
      cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      printf("%-26s waiters=%2d: 4000 dependent DFMA take %lld cycles (%.1f each)\n", names[mode], nwait, h, h / 4000.0);
    }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
