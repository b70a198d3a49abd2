// microbench3.cu — mbarrier wait flavours: how long one try_wait probe suspends, with and without a suspend-time
// hint, and what a waiting warp costs the rest of the SM (probes issued per wait).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../sgdnet_b200/csrc/common.cuh"
using namespace sgd;

__device__ __forceinline__ bool try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(ns) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

// warp 0 arrives on bar[i] every `gap` cycles; warp 1 (all lanes or lane 0 only) waits; reports wake latency and probes
__global__ void k_wait(long long* out, int mode, int gap, int lane0_only, uint32_t hint) {
  __shared__ uint64_t bar[32];
  __shared__ long long t_arrive[32];
  __shared__ volatile int ack;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int i = 0; i < 32; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); ack = 0; }
  __syncthreads();
  const int rounds = 2000;
  if (warp == 0) {
    for (int q = 0; q < rounds; ++q) {
      while (ack < q) {}
      long long t0 = clock64();
      while (clock64() - t0 < gap) {}
      if (lane == 0) { t_arrive[q & 31] = clock64(); mbar_arrive(&bar[q & 31]); }
      __syncwarp();
    }
  } else if (warp == 1) {
    long long lat = 0, probes = 0;
    for (int q = 0; q < rounds; ++q) {
      if (!lane0_only || lane == 0) {
        const uint32_t par = (q >> 5) & 1;
        bool ok = false;
        while (!ok) {
          ++probes;
          if (mode == 0) ok = mbar_try_wait(&bar[q & 31], par);
          else if (mode == 1) ok = try_wait_hint(&bar[q & 31], par, hint);
          else { ok = test_wait(&bar[q & 31], par); if (!ok) __nanosleep(hint); }
        }
        lat += clock64() - t_arrive[q & 31];
      }
      __syncwarp();
      if (lane == 0) ack = q + 1;
    }
    if (lane == 0) { out[0] = lat / rounds; out[1] = probes / rounds; }
  }
}

int main() {
  long long* d; cudaMalloc(&d, 64); long long h[2];
  const char* names[] = {"try_wait", "try_wait+hint", "test_wait+nanosleep"};
  for (int lane0 = 0; lane0 < 2; ++lane0)
    for (int mode = 0; mode < 3; ++mode)
      for (uint32_t hint : {100u, 1000u, 20000u}) {
        if (mode == 0 && hint != 100u) continue;
        for (int gap : {500, 5000}) {
          k_wait<<<1, 64>>>(d, mode, gap, lane0, hint);
          cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
          printf("%-22s hint=%5u ns  lane0_only=%d gap=%5d: wake latency %5lld cycles, probes per wait %5lld\n", names[mode], hint, lane0, gap, h[0], h[1]);
        }
      }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
