// saga_dense.cu — dense SAGA epochs, one persistent CTA per fit (reference: src/saga-dense.h:147-212).
//
// Data layout: X is [n][ld] row-major in HBM (each sample contiguous, the reference's AdaptiveTranspose,
// src/utils.h:283-288), already standardised; rows are streamed into a shared-memory ring with 1-D bulk copies
// (cp.async.bulk -> UBLKCP) driven by the host-precomputed sample sequence, kRing-1 rows ahead of use.
// State: W and g_sum (K x p each, class-major) live in shared memory for the whole launch when they fit
// (16*K*p bytes), otherwise they are used in place in HBM/L2. Every thread owns the features j = tid, tid+T, ...
// for all classes, so W/g_sum are only ever touched by their owner: the only cross-thread traffic per update is
// the K-value dot-product reduction and the K-value gradient change.
//
// Per update, T = blockDim.x:
//   A. partial dot products over the owned features, warp shuffle + one shared-memory stage     -> lp[k]
//   B. gradient, g_change, gradient memory, intercept (K == 1: every thread redundantly, no second barrier;
//      K > 1: one lane per class in warp 0, broadcast through shared memory)
//   C. fused sweep over the owned features: W -= g_change*x*(gamma/wscale); prox; g_sum += g_change*x/n
// Algorithmic HBM bytes per update: 8*p (row) + 4 (index) + 8*K_y (y) + 16*K (gradient memory read+write).
#include "common.cuh"
#include "kernels.h"

namespace sgd {

constexpr int kDenseThreads = 256;
constexpr int kRing = 4;

struct DenseSmem {
  double* W;
  double* G;
  double* ring;     // [kRing][ld]
  double* red;      // [2][nwarps][K], double-buffered by update parity
  double* gch;      // [K]
  uint64_t* full;   // [kRing]
};

__device__ __forceinline__ DenseSmem carve_dense(unsigned char* base, int K, int p, int ld, int nwarps, bool state_in_smem) {
  DenseSmem s;
  size_t off = 0;
  s.ring = reinterpret_cast<double*>(base + off); off += sizeof(double) * kRing * ld;
  s.full = reinterpret_cast<uint64_t*>(base + off); off += sizeof(uint64_t) * kRing;
  s.red = reinterpret_cast<double*>(base + off); off += sizeof(double) * 2 * nwarps * K;
  s.gch = reinterpret_cast<double*>(base + off); off += sizeof(double) * K;
  off = (off + 15) & ~size_t(15);
  if (state_in_smem) {
    s.W = reinterpret_cast<double*>(base + off); off += sizeof(double) * K * p;
    s.G = reinterpret_cast<double*>(base + off);
  } else {
    s.W = nullptr;
    s.G = nullptr;
  }
  return s;
}

size_t dense_smem_bytes(int K, int p, int ld, int* state_in_smem) {
  const int nwarps = kDenseThreads / 32;
  size_t fixed = sizeof(double) * kRing * ld + sizeof(uint64_t) * kRing + sizeof(double) * 2 * nwarps * K + sizeof(double) * K;
  fixed = (fixed + 15) & ~size_t(15);
  size_t state = sizeof(double) * 2 * size_t(K) * p;
  const size_t budget = 227 * 1024;
  if (fixed + state <= budget) {
    *state_in_smem = 1;
    return fixed + state;
  }
  *state_in_smem = 0;
  return fixed;
}

template <bool kScalar>
__global__ void __launch_bounds__(kDenseThreads, 1)
saga_dense_kernel(FitDev* __restrict__ fits, Progress* __restrict__ prog, const RoundArgs* __restrict__ args) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int fit_id = blockIdx.x;
  const RoundArgs ra = args[fit_id];
  Progress& pg = prog[fit_id];
  if (ra.n_epochs <= 0 || pg.status != kRunning) return;
  const bool free_run = (ra.flags & 1) != 0;
  const FitDev& f = fits[fit_id];

  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
  const int K = kScalar ? 1 : f.K, p = f.p, ld = f.ld, Ky = f.Ky;
  const int64_t n = f.n;
  const double nd = static_cast<double>(static_cast<uint32_t>(n));
  const int family = f.family, pen = f.penalty;
  const bool fit_intercept = f.fit_intercept != 0;

  int state_in_smem;
  {
    // same decision as the host made when sizing dynamic shared memory
    size_t fixed = sizeof(double) * kRing * ld + sizeof(uint64_t) * kRing + sizeof(double) * 2 * nwarps * K + sizeof(double) * K;
    fixed = (fixed + 15) & ~size_t(15);
    state_in_smem = (fixed + sizeof(double) * 2 * size_t(K) * p <= size_t(227 * 1024)) ? 1 : 0;
  }
  DenseSmem sm = carve_dense(smem_raw, K, p, ld, nwarps, state_in_smem != 0);
  double* W = state_in_smem ? sm.W : f.W;
  double* G = state_in_smem ? sm.G : f.gsum;

  const int li = pg.lambda_ind;
  const double gamma = f.gamma[li], alpha = f.alpha[li], beta = f.beta[li];
  const double r = 1.0 - alpha * gamma;     // wscale_update
  const uint32_t row_bytes = static_cast<uint32_t>(ld) * 8u;

  if (tid == 0) {
    for (int i = 0; i < kRing; ++i) mbar_init(&sm.full[i], 1);
    fence_barrier_init();
  }
  if (state_in_smem) {
    for (int e = tid; e < K * p; e += T) {
      W[e] = f.W[e];
      G[e] = f.gsum[e];
    }
  }
  __syncthreads();

  const uint32_t* __restrict__ seq = ra.seq;
  const int64_t total = n * ra.n_epochs;    // updates this launch may run
  int64_t issued = 0;                        // rows whose copy has been issued (thread 0 only)
  if (tid == 0) {
    for (; issued < kRing - 1 && issued < total; ++issued) {
      const int slot = static_cast<int>(issued % kRing);
      mbar_expect_tx(&sm.full[slot], row_bytes);
      bulk_g2s(sm.ring + size_t(slot) * ld, f.xd + size_t(seq[issued]) * ld, row_bytes, &sm.full[slot]);
    }
  }

  // intercept state: K == 1 keeps a private, identical copy in every thread; K > 1 keeps class k in lane k of warp 0
  double b_reg = 0.0, gsi_reg = 0.0;
  if (kScalar || (warp == 0 && lane < K)) {
    b_reg = f.b[kScalar ? 0 : lane];
    gsi_reg = f.gsi[kScalar ? 0 : lane];
  }

  double wscale = 1.0;
  uint32_t it_outer = pg.it_outer;
  uint32_t epochs_done = 0;
  int64_t tg = 0;                            // global update counter within this launch
  uint32_t prev_s = 0xffffffffu;
  double prev_g = 0.0;
  bool finished = false;

  for (int ep = 0; ep < ra.n_epochs && !finished; ++ep) {
    for (int64_t t = 0; t < n; ++t, ++tg) {
      const uint32_t s = seq[tg];
      const int slot = static_cast<int>(tg % kRing);
      const uint32_t parity = static_cast<uint32_t>((tg / kRing) & 1);

      // small per-sample operands, issued before the wait so their latency overlaps the dot product
      double y_val = 0.0, gm_val = 0.0;
      if (kScalar) {
        y_val = f.yt[s];
        gm_val = (s == prev_s) ? prev_g : f.gmem[s];   // the previous update's store may still be in flight
      } else if (warp == 0 && lane < K) {
        y_val = f.yt[size_t(s) * Ky + (Ky == 1 ? 0 : lane)];
        gm_val = f.gmem[size_t(s) * K + lane];
      }

      mbar_wait(&sm.full[slot], parity);
      const double* __restrict__ xr = sm.ring + size_t(slot) * ld;
      double* red = sm.red + size_t(tg & 1) * nwarps * K;   // step B of update t may still be reading the other half

      // ---- A: dot products
      if (kScalar) {
        double acc = 0.0;
        for (int j = tid; j < p; j += T) acc += W[j] * xr[j];
        acc = warp_sum(acc);
        if (lane == 0) red[warp] = acc;
      } else {
        for (int k = 0; k < K; ++k) {
          double acc = 0.0;
          const double* Wk = W + size_t(k) * p;
          for (int j = tid; j < p; j += T) acc += Wk[j] * xr[j];
          acc = warp_sum(acc);
          if (lane == 0) red[warp * K + k] = acc;
        }
      }
      __syncthreads();   // (1) partial sums visible; every thread is past step C of the previous update

      if (tid == 0 && issued < total) {      // refill the slot the previous update just released
        const int fslot = static_cast<int>(issued % kRing);
        mbar_expect_tx(&sm.full[fslot], row_bytes);
        bulk_g2s(sm.ring + size_t(fslot) * ld, f.xd + size_t(seq[issued]) * ld, row_bytes, &sm.full[fslot]);
        ++issued;
      }

      // ---- B: gradient, gradient memory, intercept
      double gch_scalar = 0.0;
      if (kScalar) {
        double dot = 0.0;
        for (int w = 0; w < nwarps; ++w) dot += red[w];
        const double lp = dot * wscale + b_reg;
        const double g = gradient_scalar(family, lp, y_val);
        gch_scalar = g - gm_val;
        if (tid == 0) f.gmem[s] = g;
        prev_s = s;
        prev_g = g;
        if (wscale < kSmall) {
          for (int j = tid; j < p; j += T) W[j] *= wscale;
          wscale = 1.0;
        }
        wscale *= r;
        if (fit_intercept) {
          gsi_reg += gch_scalar / nd;
          b_reg -= gamma * (gsi_reg + gch_scalar / nd);
        }
      } else {
        if (warp == 0) {
          const bool valid = lane < K;
          double lp = 0.0;
          if (valid) {
            double dot = 0.0;
            for (int w = 0; w < nwarps; ++w) dot += red[w * K + lane];
            lp = dot * wscale + b_reg;
          }
          double g;
          if (family == kMultinomial) {
            // every lane needs the class id of this sample; y has one column
            const double yc = __shfl_sync(0xffffffffu, y_val, 0);
            const double lse = lse_warp(lp, valid);
            g = sgd_exp(lp - lse);
            if (static_cast<unsigned>(lane) == static_cast<unsigned>(yc + 0.5)) g -= 1.0;
          } else {
            g = lp - y_val;
          }
          if (valid) {
            const double gch = g - gm_val;
            f.gmem[size_t(s) * K + lane] = g;
            if (fit_intercept) {
              gsi_reg += gch / nd;
              b_reg -= gamma * (gsi_reg + gch / nd);
            }
            sm.gch[lane] = gch;
          }
        }
        if (wscale < kSmall) {
          for (int j = tid; j < p; j += T)
            for (int k = 0; k < K; ++k) W[size_t(k) * p + j] *= wscale;
          wscale = 1.0;
        }
        wscale *= r;
        __syncthreads();   // (2) g_change visible
      }

      // ---- C: fused coefficient step, prox, gradient-average update on the owned features
      const double gw = gamma / wscale;
      const PenCoef pc = pen_coef(gamma, beta, wscale, 1.0);
      if (kScalar) {
        for (int j = tid; j < p; j += T) {
          const double xj = xr[j];
          const double gx = gch_scalar * xj;
          const double gs = G[j];
          double w = W[j] - gx * gw;
          W[j] = penalty_scalar(pen, w, gs, pc);
          G[j] = gs + gx / nd;
        }
      } else {
        for (int j = tid; j < p; j += T) {
          const double xj = xr[j];
          for (int k = 0; k < K; ++k) W[size_t(k) * p + j] -= sm.gch[k] * xj * gw;
          apply_penalty(pen, W + j, G + j, K, p, pc);
          for (int k = 0; k < K; ++k) G[size_t(k) * p + j] += sm.gch[k] * xj / nd;
        }
      }
    }

    // ---- epoch end: unscale, convergence (src/saga-dense.h:188-208, src/utils.h:240-262)
    double mc = 0.0, ms = 0.0;
    for (int j = tid; j < p; j += T)
      for (int k = 0; k < K; ++k) {
        const size_t e = size_t(k) * p + j;
        const double w = W[e] * wscale;
        W[e] = w;
        mc = fmax(mc, fabs(w - f.Wprev[e]));
        ms = fmax(ms, fabs(w));
        f.Wprev[e] = w;
      }
    wscale = 1.0;
    mc = warp_max(mc);
    ms = warp_max(ms);
    __syncthreads();       // red[] is free again (all threads are past step B of the last update)
    if (lane == 0) {
      sm.red[warp] = mc;   // nwarps*K >= nwarps doubles
    }
    __syncthreads();
    double mc_all = 0.0;
    for (int w = 0; w < nwarps; ++w) mc_all = fmax(mc_all, sm.red[w]);
    __syncthreads();
    if (lane == 0) sm.red[warp] = ms;
    __syncthreads();
    double ms_all = 0.0;
    for (int w = 0; w < nwarps; ++w) ms_all = fmax(ms_all, sm.red[w]);
    __syncthreads();

    const bool all_zero = (ms_all == 0.0) && (mc_all == 0.0);
    const bool no_change = (ms_all != 0.0) && (mc_all / ms_all <= f.tol);
    ++it_outer;
    ++epochs_done;
    finished = !free_run && ((all_zero || no_change) || !(it_outer < f.max_iter));
  }

  // drain copies that were issued but never consumed (early stop) before the shared memory goes away
  if (tid == 0)
    for (int64_t q = tg; q < issued; ++q)
      mbar_wait(&sm.full[static_cast<int>(q % kRing)], static_cast<uint32_t>((q / kRing) & 1));
  __syncthreads();

  if (state_in_smem) {
    for (int e = tid; e < K * p; e += T) {
      f.W[e] = W[e];
      f.gsum[e] = G[e];
    }
  }
  if (kScalar ? (tid == 0) : (warp == 0 && lane < K)) {
    f.b[kScalar ? 0 : lane] = b_reg;
    f.gsi[kScalar ? 0 : lane] = gsi_reg;
  }
  if (tid == 0) {
    pg.it_outer = it_outer;
    pg.epochs_last_launch = epochs_done;
    if (finished) {
      pg.status = kLambdaDone;
      f.epochs[li] = it_outer;
      f.codes[li] = (it_outer == f.max_iter) ? 1u : 0u;
      pg.npasses += it_outer;
    }
  }
}

cudaError_t launch_saga_dense(int n_fits, bool scalar, size_t smem, FitDev* fits, Progress* prog, const RoundArgs* args,
                              cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(saga_dense_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(saga_dense_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (scalar)
    saga_dense_kernel<true><<<n_fits, kDenseThreads, smem, st>>>(fits, prog, args);
  else
    saga_dense_kernel<false><<<n_fits, kDenseThreads, smem, st>>>(fits, prog, args);
  return cudaGetLastError();
}

}  // namespace sgd
