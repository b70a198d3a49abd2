// microbench6.cu — dependent-latency of the binomial intercept chain as the wavefront kernel runs it, and of candidate
// rearrangements of sgd_exp (same value set is NOT required here: this only sizes the latency of each form).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false scripts/microbench6.cu -o scripts/microbench6.bin
#include <cstdio>
#include <cuda_runtime.h>
#include "../sgdnet_b200/csrc/common.cuh"
using namespace sgd;

#define N_IT 4000

// Estrin arrangement of the same degree-6 expm1 polynomial
__device__ __forceinline__ double exp_estrin(double x) {
  const double kInvStep = 46.16624130844683, kStepHi = 0.021660849392446835, kStepLo = 5.145609244655338e-14;
  const double kShift = 6755399441055744.0;
  const double ts = fma(x, kInvStep, kShift);
  const double kd = ts - kShift;
  const int32_t k = __double2loint(ts);
  double r = fma(-kd, kStepHi, x);
  r = fma(-kd, kStepLo, r);
  const double r2 = r * r;
  const double t1 = fma(r, 1.0 / 6.0, 0.5);
  const double t2 = fma(r, 1.0 / 120.0, 1.0 / 24.0);
  const double t3 = fma(r2, 1.0 / 720.0, t2);
  const double pp = fma(r2, t3, t1);
  const double q = fma(r2, pp, r);
  const int32_t j = k & 31, m = k >> 5;
  const double thi = sgd_exp_tab_dev[2 * j], tlo = sgd_exp_tab_dev[2 * j + 1];
  const double res = thi + fma(thi, q, tlo);
  return res * __hiloint2double((m + 1023) << 20, 0);
}

// one-step range reduction (single fma with the rounded constant; the low part folded into the polynomial argument later)
__device__ __forceinline__ double exp_estrin_1step(double x) {
  const double kInvStep = 46.16624130844683, kStepHi = 0.021660849392446835, kStepLo = 5.145609244655338e-14;
  const double kShift = 6755399441055744.0;
  const double ts = fma(x, kInvStep, kShift);
  const double kd = ts - kShift;
  const int32_t k = __double2loint(ts);
  const double rh = fma(-kd, kStepHi, x);
  const double rl = kd * kStepLo;          // parallel with rh
  const double r = rh - rl;
  const double r2 = r * r;
  const double t1 = fma(r, 1.0 / 6.0, 0.5);
  const double t2 = fma(r, 1.0 / 120.0, 1.0 / 24.0);
  const double t3 = fma(r2, 1.0 / 720.0, t2);
  const double pp = fma(r2, t3, t1);
  const double q = fma(r2, pp, r);
  const int32_t j = k & 31, m = k >> 5;
  const double thi = sgd_exp_tab_dev[2 * j], tlo = sgd_exp_tab_dev[2 * j + 1];
  const double res = thi + fma(thi, q, tlo);
  return res * __hiloint2double((m + 1023) << 20, 0);
}

template <int V>
__device__ __forceinline__ double expv(double x) {
  if (V == 0) return sgd_exp_inrange(x);
  if (V == 1) return exp_estrin(x);
  return exp_estrin_1step(x);
}

template <int V>
__device__ void chain(double* out, long long* cyc, double seed, int slot) {
  double b = 0.1, gsi = 0.01;
  const double nd = 1e6, rn = 1.0 / nd, gamma = 0.3, dot = seed, gm = 0.2, ya = 0.0;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N_IT; ++i) {
    const double lp = dot + b;
    const double g = ya - 1.0 / (1.0 + expv<V>(lp));
    const double gch = g - gm;
    const double gn = div_by_n(gch, nd, rn);
    gsi += gn;
    b -= gamma * (gsi * 0.01 + gn);
  }
  long long t1 = clock64();
  cyc[slot] = t1 - t0;
  out[slot] = b + gsi;
}

__global__ void k(double* out, long long* cyc, double seed) {
  chain<0>(out, cyc, seed, 0);
  chain<1>(out, cyc, seed, 1);
  chain<2>(out, cyc, seed, 2);
  double x = seed; long long t0, t1;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N_IT; ++i) x = sgd_exp_inrange(x) * 0.3;
  t1 = clock64(); cyc[3] = t1 - t0; out[3] = x;
  x = seed; t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N_IT; ++i) x = exp_estrin(x) * 0.3;
  t1 = clock64(); cyc[4] = t1 - t0; out[4] = x;
  x = seed; t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N_IT; ++i) x = 1.0 / (1.0 + x);
  t1 = clock64(); cyc[5] = t1 - t0; out[5] = x;
  x = seed; const double nd = 1e6, rn = 1.0 / nd; t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N_IT; ++i) x = div_by_n(x, nd, rn) + 1.0;
  t1 = clock64(); cyc[6] = t1 - t0; out[6] = x;
  // intercept tail: 5 dependent ops
  double b = 0.1, gsi = 0.01; t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N_IT; ++i) { const double gn = b * 1e-7; gsi += gn; b -= 0.3 * (gsi * 0.01 + gn); }
  t1 = clock64(); cyc[7] = t1 - t0; out[7] = b;
  x = seed; t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N_IT; ++i) x = warp_sum(x) * 0.03125;
  t1 = clock64(); cyc[8] = t1 - t0; out[8] = x;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 64 * 8); cudaMalloc(&cyc, 64 * 8);
  long long h[16];
  for (int rep = 0; rep < 2; ++rep) { k<<<1, 32>>>(out, cyc, 0.5); cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost); }
  const char* names[] = {"chain (kernel form: sgd_exp_inrange, div_by_n)", "chain, Estrin exp", "chain, Estrin exp + 1-step reduction",
                         "sgd_exp_inrange (+mul)", "exp Estrin (+mul)", "1/(1+x)", "div_by_n (+add)", "intercept tail (mul,add,mul,add,mul,sub)", "warp_sum(double) (+mul)"};
  for (int i = 0; i < 9; ++i) printf("%-52s %8.1f cycles/iter\n", names[i], double(h[i]) / N_IT);
  printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
