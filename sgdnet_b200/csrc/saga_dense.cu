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
  return fixed;     // may exceed the budget for very wide rows: the host refuses such a fit (engine.cu, finalize_batch)
}
size_t dense_smem_budget() { return 227 * 1024; }

// KT: number of classes rounded up to 1 (K == 1: gaussian / binomial), 4, 8, 16 or 32; PEN: the penalty functor.
// Both are compile-time so that the per-class loops unroll: the KT dot-product butterflies of a warp then run
// interleaved (one butterfly's latency instead of K of them) and the coefficient sweep is straight-line code.
template <int KT, int PEN>
__global__ void __launch_bounds__(kDenseThreads, 1)
saga_dense_kernel(FitDev* __restrict__ fit, Progress* __restrict__ prog, const RoundArgs ra) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr bool kScalar = (KT == 1);
  Progress& pg = *prog;
  const FitDev& f = *fit;
  if (ra.n_epochs <= 0 || pg.status != kRunning) {
    if (threadIdx.x == 0) {
      pg.epochs_last_launch = 0;
      publish_progress(f.mirror, pg, ra.round_id);
    }
    return;
  }
  const bool free_run = (ra.flags & 1) != 0;
  const uint64_t t_start = globaltimer_ns();

  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
  const int K = kScalar ? 1 : f.K, p = f.p, ld = f.ld, Ky = f.Ky;
  const int64_t n = f.n;
  const double nd = static_cast<double>(static_cast<uint32_t>(n));
  const double rn = 1.0 / nd;
  const int family = f.family;
  const bool fit_intercept = f.fit_intercept != 0;

  int state_in_smem;
  {
    // W / g_sum live in shared memory when this fit's own layout fits the dynamic shared memory the host gave the
    // launch (sized from the largest fit of the batch; fits of a batch may differ in K and p)
    size_t fixed = sizeof(double) * kRing * ld + sizeof(uint64_t) * kRing + sizeof(double) * 2 * nwarps * K + sizeof(double) * K;
    fixed = (fixed + 15) & ~size_t(15);
    uint32_t dyn_bytes;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_bytes));
    state_in_smem = (fixed + sizeof(double) * 2 * size_t(K) * p <= size_t(dyn_bytes)) ? 1 : 0;
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
  uint32_t prev_s = 0xffffffffu, prev2_s = 0xffffffffu;
  double prev_g = 0.0, prev2_g = 0.0;
  bool finished = false;

  // Per-sample operands (index, y, gradient memory) are fetched ONE UPDATE AHEAD so that their HBM latency never sits
  // on the gradient step. A gradient-memory value fetched that early can be stale for the last two samples (their
  // stores are not ordered before the fetch), so those are forwarded from registers instead.
  const bool owner = kScalar || (warp == 0 && lane < K);
  auto fetch_y = [&](uint32_t sx) { return kScalar ? f.yt[sx] : f.yt[size_t(sx) * Ky + (Ky == 1 ? 0 : lane)]; };
  auto fetch_gm = [&](uint32_t sx) { return kScalar ? f.gmem[sx] : f.gmem[size_t(sx) * K + lane]; };
  uint32_t s_cur = seq[0], s_nxt = (total > 1) ? seq[1] : 0u;
  double y_cur = 0.0, gm_cur = 0.0;
  if (owner) {
    y_cur = fetch_y(s_cur);
    gm_cur = fetch_gm(s_cur);
  }
  uint32_t s_refill = (tid == 0 && issued < total) ? seq[issued] : 0u;   // sample of the next row copy (thread 0)

  for (int ep = 0; ep < ra.n_epochs && !finished; ++ep) {
    for (int64_t t = 0; t < n; ++t, ++tg) {
      const uint32_t s = s_cur;
      const int slot = static_cast<int>(tg % kRing);
      const uint32_t parity = static_cast<uint32_t>((tg / kRing) & 1);
      const double y_val = y_cur;
      const double gm_val = (s == prev_s) ? prev_g : ((s == prev2_s) ? prev2_g : gm_cur);
      // next update's operands
      s_cur = s_nxt;
      if (tg + 2 < total) s_nxt = seq[tg + 2];
      if (owner && tg + 1 < total) {
        y_cur = fetch_y(s_cur);
        gm_cur = fetch_gm(s_cur);
      }

      mbar_wait(&sm.full[slot], parity);
      const double* __restrict__ xr = sm.ring + size_t(slot) * ld;
      double* red = sm.red + size_t(tg & 1) * nwarps * K;   // step B of update t may still be reading the other half

      // ---- A: dot products. Thread tid owns the running sums of features j = tid (mod 256) for every class; the
      // warp's KT butterflies are independent and interleave (sgdnet_arith.h, item 2).
      {
        double acc[KT];
#pragma unroll
        for (int k = 0; k < KT; ++k) acc[k] = 0.0;
        for (int j = tid; j < p; j += T) {
          const double xj = xr[j];
#pragma unroll
          for (int k = 0; k < KT; ++k)
            if (kScalar || k < K) acc[k] += W[size_t(k) * p + j] * xj;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
          for (int k = 0; k < KT; ++k)
            if (kScalar || k < K) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
        }
        if (lane == 0) {
#pragma unroll
          for (int k = 0; k < KT; ++k)
            if (kScalar || k < K) red[warp * K + k] = acc[k];
        }
      }
      __syncthreads();   // (1) partial sums visible; every thread is past step C of the previous update

      if (tid == 0 && issued < total) {      // refill the slot the previous update just released
        const int fslot = static_cast<int>(issued % kRing);
        mbar_expect_tx(&sm.full[fslot], row_bytes);
        bulk_g2s(sm.ring + size_t(fslot) * ld, f.xd + size_t(s_refill) * ld, row_bytes, &sm.full[fslot]);
        ++issued;
        if (issued < total) s_refill = seq[issued];
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
        if (s != prev_s) {
          prev2_s = prev_s;
          prev2_g = prev_g;
        }
        prev_s = s;
        prev_g = g;
        if (wscale < kSmall) {
          for (int j = tid; j < p; j += T) W[j] *= wscale;
          wscale = 1.0;
        }
        wscale *= r;
        if (fit_intercept) {
          const double gn = div_by_n(gch_scalar, nd, rn);
          gsi_reg += gn;
          b_reg -= gamma * (gsi_reg + gn);
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
              const double gn = div_by_n(gch, nd, rn);
              gsi_reg += gn;
              b_reg -= gamma * (gsi_reg + gn);
            }
            sm.gch[lane] = gch;
          }
          if (s != prev_s) {
            prev2_s = prev_s;
            prev2_g = prev_g;
          }
          prev_s = s;
          prev_g = g;
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
      // (src/saga-dense.h:176-183; penalty functors src/penalties.h:27-79 with scaling = 1)
      const double gw = gamma / wscale;
      const double step = gamma / wscale * 1.0;
      const double bgs = beta * gamma * 1.0;
      const double thr = bgs / wscale;
      double gch[KT];
#pragma unroll
      for (int k = 0; k < KT; ++k) gch[k] = kScalar ? gch_scalar : ((k < K) ? sm.gch[k] : 0.0);
      for (int j = tid; j < p; j += T) {
        const double xj = xr[j];
        double w[KT], gs[KT];
        double sq = 0.0;
#pragma unroll
        for (int k = 0; k < KT; ++k) {
          if (kScalar || k < K) {
            gs[k] = G[size_t(k) * p + j];
            const double gx = gch[k] * xj;
            const double v = (W[size_t(k) * p + j] - gx * gw) - step * gs[k];
            w[k] = (PEN == kElasticNet) ? soft_threshold(v, thr) : v;
            if (PEN == kGroupLasso) sq += v * v;
            G[size_t(k) * p + j] = gs[k] + div_by_n(gx, nd, rn);
          }
        }
        if (PEN == kGroupLasso) {
          const double factor = bgs / sqrt(sq);
          const double mult = 1.0 - factor / wscale;
#pragma unroll
          for (int k = 0; k < KT; ++k)
            if (kScalar || k < K) w[k] = (factor < 1.0) ? w[k] * mult : 0.0;
        }
#pragma unroll
        for (int k = 0; k < KT; ++k)
          if (kScalar || k < K) W[size_t(k) * p + j] = w[k];
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
    pg.solver_ns += globaltimer_ns() - t_start;
    publish_progress(f.mirror, pg, ra.round_id);
  }
}

template <int KT, int PEN>
static cudaError_t launch_dense_variant(size_t smem, FitDev* fit, Progress* prog, const RoundArgs& ra, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(saga_dense_kernel<KT, PEN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  saga_dense_kernel<KT, PEN><<<1, kDenseThreads, smem, st>>>(fit, prog, ra);
  return cudaGetLastError();
}

template <int KT>
static cudaError_t launch_dense_kt(int pen, size_t smem, FitDev* fit, Progress* prog, const RoundArgs& ra, cudaStream_t st) {
  switch (pen) {
    case kRidge: return launch_dense_variant<KT, kRidge>(smem, fit, prog, ra, st);
    case kElasticNet: return launch_dense_variant<KT, kElasticNet>(smem, fit, prog, ra, st);
    default: return launch_dense_variant<KT, kGroupLasso>(smem, fit, prog, ra, st);
  }
}

int dense_kt_bucket(int K) { return K == 1 ? 1 : K <= 4 ? 4 : K <= 8 ? 8 : K <= 16 ? 16 : 32; }

// One launch per fit: the instantiation for the fit's class-count bucket (1, 4, 8, 16, 32) and penalty.
cudaError_t launch_saga_dense(int K, int pen, size_t smem, FitDev* fit, Progress* prog, const RoundArgs& ra, cudaStream_t st) {
  switch (dense_kt_bucket(K)) {
    case 1: return launch_dense_kt<1>(pen, smem, fit, prog, ra, st);
    case 4: return launch_dense_kt<4>(pen, smem, fit, prog, ra, st);
    case 8: return launch_dense_kt<8>(pen, smem, fit, prog, ra, st);
    case 16: return launch_dense_kt<16>(pen, smem, fit, prog, ra, st);
    default: return launch_dense_kt<32>(pen, smem, fit, prog, ra, st);
  }
}

}  // namespace sgd
