// rng.cu — R's Mersenne-Twister and the sampling sequence on the device.
//
// The reference draws one sample index per update from R's global generator: `floor(R::runif(0.0, n_samples))`
// (src/saga-sparse.h:261, src/saga-dense.h:152; R core src/main/RNG.c MT_genrand + fixup, src/nmath/runif.c). With
// R's default generator (SGDNET_RNG_MT) the library produces that stream where it is consumed: the generator state
// (624 words + position) lives in HBM, one CTA regenerates it a block of 624 words at a time and writes the indices of
// the next launch; the host neither draws nor uploads indices. Every operation is integer arithmetic except
// word * 2^-32-ish and n * u, single IEEE multiplications that round identically on both sides, so the sequence equals
// the host generator's (host_setup.cu mt_unif / draw_indices) word for word; tests/test_abi_cpu.py checks the
// parallel regeneration schedule against the sequential one on the CPU, tests/test_parity_gpu.py checks the kernel.
//
// Block regeneration, MT19937 (N = 624, M = 397): new[k] = src[k] ^ twist(old[k], old[k+1]) with
//   src[k] = old[k + 397]      for k in [0, 227)        (phase 1)
//          = new[k - 227]      for k in [227, 454)      (phase 2: reads phase 1)
//          = new[k - 227]      for k in [454, 623)      (phase 3: reads phase 2)
//   new[623] = new[396] ^ twist(old[623], new[0])       (phase 4)
// The twist operands old[k], old[k+1] are always the OLD words (the sequential loop reads mt[k+1] before it writes
// it), so every thread takes them before the first phase writes anything.
#include <cmath>
#include <cstring>

#include "common.cuh"
#include "kernels.h"

namespace sgd {

namespace {

constexpr int kN = 624, kM = 397;
constexpr int kMtThreads = 640;

__host__ __device__ inline uint32_t mt_twist(uint32_t a, uint32_t b) {
  const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
  return (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}
__host__ __device__ inline uint32_t mt_temper(uint32_t y) {
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}
// unif_rand() = fixup(word * 2.3283064365386963e-10); index = floor(0 + (n - 0) * u)
__host__ __device__ inline uint32_t mt_index(uint32_t word, double nd) {
  double u = static_cast<double>(word) * 2.3283064365386963e-10;
  const double half_step = 0.5 * 2.328306437080797e-10;
  if (u <= 0.0) u = half_step;
  if (1.0 - u <= 0.0) u = 1.0 - half_step;
  return static_cast<uint32_t>(floor(0.0 + (nd - 0.0) * u));
}

}  // namespace

__global__ void __launch_bounds__(kMtThreads)
mt_indices_kernel(const MtState* __restrict__ src, uint32_t n, int n_epochs, uint32_t* __restrict__ seq,
                  MtState* __restrict__ snaps) {
  __shared__ uint32_t mt[kN];
  const int k = threadIdx.x;
  const double nd = static_cast<double>(n);
  if (k < kN) mt[k] = src->mt[k];
  int mti = src->mti;
  __syncthreads();
  const uint64_t total = uint64_t(n) * uint64_t(n_epochs);
  uint64_t pos = 0;          // draws written so far
  int next_snap = 0;         // snapshots written so far (snaps[e] = state after e epochs)
  for (;;) {
    // snapshots whose epoch boundary is at `pos` or inside the words still unused in this block
    const int avail = (mti >= kN) ? 0 : kN - mti;
    while (next_snap <= n_epochs && uint64_t(next_snap) * n <= pos + uint64_t(avail)) {
      const uint64_t at = uint64_t(next_snap) * n;
      if (k < kN) snaps[next_snap].mt[k] = mt[k];
      if (k == 0) snaps[next_snap].mti = mti + static_cast<int>(at - pos);
      ++next_snap;
    }
    // the unused words of this block
    if (k < kN && k >= mti) {
      const uint64_t q = pos + uint64_t(k - mti);
      if (q < total) seq[q] = mt_index(mt_temper(mt[k]), nd);
    }
    pos += uint64_t(avail);
    if (pos >= total) break;
    // regenerate (all threads; uniform control flow)
    uint32_t tw = 0, old623 = 0;
    if (k < kN - 1) tw = mt_twist(mt[k], mt[k + 1]);
    if (k == kN - 1) old623 = mt[kN - 1];
    __syncthreads();
    if (k < kN - kM) mt[k] = mt[k + kM] ^ tw;
    __syncthreads();
    if (k >= kN - kM && k < 2 * (kN - kM)) mt[k] = mt[k - (kN - kM)] ^ tw;
    __syncthreads();
    if (k >= 2 * (kN - kM) && k < kN - 1) mt[k] = mt[k - (kN - kM)] ^ tw;
    __syncthreads();
    if (k == kN - 1) mt[kN - 1] = mt[kM - 1] ^ mt_twist(old623, mt[0]);
    __syncthreads();
    mti = 0;
  }
}

cudaError_t launch_mt_indices(const MtState* src, uint32_t n, int n_epochs, uint32_t* seq, MtState* snaps, cudaStream_t st) {
  mt_indices_kernel<<<1, kMtThreads, 0, st>>>(src, n, n_epochs, seq, snaps);
  return cudaGetLastError();
}

// The kernel's schedule executed phase by phase on the host (each phase reads only what the previous phases wrote).
void mt_indices_host(const MtState* src, uint32_t n, int n_epochs, uint32_t* seq, MtState* snaps) {
  uint32_t mt[kN], tw[kN];
  std::memcpy(mt, src->mt, sizeof(mt));
  int mti = src->mti;
  const double nd = static_cast<double>(n);
  const uint64_t total = uint64_t(n) * uint64_t(n_epochs);
  uint64_t pos = 0;
  int next_snap = 0;
  for (;;) {
    const int avail = (mti >= kN) ? 0 : kN - mti;
    while (next_snap <= n_epochs && uint64_t(next_snap) * n <= pos + uint64_t(avail)) {
      const uint64_t at = uint64_t(next_snap) * n;
      std::memcpy(snaps[next_snap].mt, mt, sizeof(mt));
      snaps[next_snap].mti = mti + static_cast<int>(at - pos);
      ++next_snap;
    }
    for (int k = mti; k < kN; ++k) {
      const uint64_t q = pos + uint64_t(k - mti);
      if (q < total) seq[q] = mt_index(mt_temper(mt[k]), nd);
    }
    pos += uint64_t(avail);
    if (pos >= total) break;
    for (int k = 0; k < kN - 1; ++k) tw[k] = mt_twist(mt[k], mt[k + 1]);
    const uint32_t old623 = mt[kN - 1];
    for (int k = 0; k < kN - kM; ++k) mt[k] = mt[k + kM] ^ tw[k];
    for (int k = kN - kM; k < 2 * (kN - kM); ++k) mt[k] = mt[k - (kN - kM)] ^ tw[k];
    for (int k = 2 * (kN - kM); k < kN - 1; ++k) mt[k] = mt[k - (kN - kM)] ^ tw[k];
    mt[kN - 1] = mt[kM - 1] ^ mt_twist(old623, mt[0]);
    mti = 0;
  }
}

}  // namespace sgd
