// saga_sparse.cu — sparse SAGA epochs with just-in-time lagged prox (reference: src/saga-sparse.h:76-155, 256-371).
//
// One persistent CTA per fit. Two kernels:
//
//  saga_sparse_k1_kernel  K == 1 (gaussian / binomial), no virtual centring: the headline path (BASELINE configs
//     2 and 5). Warp 0 is the solver: one lane per nonzero of the sampled row, the row's W / g_sum / lag entries are
//     gathered once into registers, caught up (LaggedUpdate k = t), dotted with a warp-shuffle reduction, stepped,
//     proxed (LaggedUpdate k = t+1, lag 1) and scattered back - each touched coefficient is read once and written
//     once per update. Warp 1 is the producer: it walks the host-precomputed sample sequence ahead of the solver and
//     stages each row's padded-CSR index/value runs plus y into a shared-memory ring with 1-D bulk copies
//     (cp.async.bulk, complete_tx on an mbarrier per slot), so the solver never waits on HBM for row data.
//     Coefficient state (W, g_sum, lag) is addressed in place: at these sizes (2 MB at p = 100k) it is L2-resident.
//     Algorithmic HBM bytes per update: 12*nnz_row + 16 (row info) + 4 (index) + 8 (y) + 16 (gradient memory).
//
//  saga_sparse_generic_kernel  any K <= 32 and/or standardize = TRUE (the reference's O(p*K) virtual-centring sweeps,
//     src/saga-sparse.h:127-128, 276-277, reproduced as block-wide passes). Phases are separated by block barriers.
//
// Epoch end (both): Reset(n) over all features by the whole CTA, W *= wscale, lag = 0, convergence test
// (src/saga-sparse.h:340-348, 367; src/utils.h:240-262).
#include "common.cuh"
#include "kernels.h"

namespace sgd {

constexpr int kSpThreads = 128;
constexpr int kSlots = 16;      // ring depth (rows in flight)
constexpr int kCap = 128;       // entries per ring slot; longer rows are read in place
constexpr int kChunks = kCap / 32;

struct SpSlotMeta {
  uint32_t s;
  int32_t nnz;
  int64_t start;
  double y;
};

struct __align__(128) SpRing {
  int32_t idx[kSlots][kCap];
  double val[kSlots][kCap];
  SpSlotMeta meta[kSlots];
  uint64_t full[kSlots];
  uint64_t empty[kSlots];
};

__device__ __forceinline__ double lag_scale(bool identity, const double* __restrict__ table, uint32_t m) {
  return identity ? static_cast<double>(m) : table[m];
}

// Reset(k) on feature range owned by the caller (src/saga-sparse.h:132-155) fused with the epoch-end bookkeeping.
__device__ __forceinline__ void epoch_end_feature(const FitDev& f, int K, int p, int j, uint32_t k_it, double wscale,
                                                  int pen, double gamma, double beta, bool identity, double& mc,
                                                  double& ms) {
  const uint32_t m = k_it - f.lag[j];
  if (m != 0) {
    const PenCoef pc = pen_coef(gamma, beta, wscale, lag_scale(identity, f.lag_scaling, m));
    apply_penalty(pen, f.W + j, f.gsum + j, K, p, pc);
  }
  f.lag[j] = 0;
  for (int k = 0; k < K; ++k) {
    const size_t e = size_t(k) * p + j;
    const double w = f.W[e] * wscale;
    f.W[e] = w;
    mc = fmax(mc, fabs(w - f.Wprev[e]));
    ms = fmax(ms, fabs(w));
    f.Wprev[e] = w;
  }
}

// Block-wide max of (mc, ms) and the convergence decision; `red` holds 2*nwarps doubles.
__device__ __forceinline__ bool block_converged(double mc, double ms, double* red, double tol) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  mc = warp_max(mc);
  ms = warp_max(ms);
  if (lane == 0) {
    red[warp] = mc;
    red[nwarps + warp] = ms;
  }
  __syncthreads();
  double mc_all = 0.0, ms_all = 0.0;
  for (int w = 0; w < nwarps; ++w) {
    mc_all = fmax(mc_all, red[w]);
    ms_all = fmax(ms_all, red[nwarps + w]);
  }
  __syncthreads();
  const bool all_zero = (ms_all == 0.0) && (mc_all == 0.0);
  const bool no_change = (ms_all != 0.0) && (mc_all / ms_all <= tol);
  return all_zero || no_change;
}

// ============================================================================================ K == 1 fast path
__global__ void __launch_bounds__(kSpThreads, 1)
saga_sparse_k1_kernel(FitDev* __restrict__ fits, Progress* __restrict__ prog, const RoundArgs* __restrict__ args) {
  __shared__ SpRing ring;
  __shared__ double red[2 * (kSpThreads / 32)];
  __shared__ double wscale_s;      // only the solver warp tracks wscale; the epoch-end Reset needs it block-wide

  const int fit_id = blockIdx.x;
  const RoundArgs ra = args[fit_id];
  Progress& pg = prog[fit_id];
  if (ra.n_epochs <= 0 || pg.status != kRunning) return;
  const bool free_run = (ra.flags & 1) != 0;
  const FitDev& f = fits[fit_id];

  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int p = f.p;
  const int64_t n = f.n;
  const uint32_t n32 = static_cast<uint32_t>(n);
  const double nd = static_cast<double>(n32);
  const int family = f.family, pen = f.penalty;
  const bool fit_intercept = f.fit_intercept != 0;

  const int li = pg.lambda_ind;
  const double gamma = f.gamma[li], alpha = f.alpha[li], beta = f.beta[li];
  const double r = 1.0 - alpha * gamma;
  const bool identity = (r == 1.0);          // lasso: lag_scaling[m] == m exactly
  const double sc2 = 1.0 / nd;
  const double bg = beta * gamma;            // (beta*gamma)*1.0

  if (tid == 0) {
    for (int i = 0; i < kSlots; ++i) {
      mbar_init(&ring.full[i], 1);
      mbar_init(&ring.empty[i], 1);
    }
    fence_barrier_init();
  }
  __syncthreads();

  double* __restrict__ W = f.W;
  double* __restrict__ G = f.gsum;
  uint32_t* __restrict__ lag = f.lag;
  const uint32_t* __restrict__ seq = ra.seq;

  double b_reg = f.b[0], gsi_reg = f.gsi[0];
  double wscale = 1.0;
  uint32_t it_outer = pg.it_outer, epochs_done = 0;
  bool finished = false;
  int64_t q_base = 0;       // ring sequence number of the first row of the current epoch

  for (int ep = 0; ep < ra.n_epochs && !finished; ++ep, q_base += n) {
    const uint32_t* __restrict__ eseq = seq + size_t(ep) * n;

    if (warp == 1) {
      // ------------------------------------------------------------------ producer: lane l feeds ring slot l
      // Warp-synchronous: the 16 lanes load the sample index, row descriptor and response of their next row together
      // (16 independent HBM requests in flight), one batch ahead of the batch being issued, then the warp polls the
      // slots' "empty" barriers and issues each row's two bulk copies as soon as its slot is released.
      const bool feeder = lane < kSlots;
      int64_t t = (lane - static_cast<int>(q_base % kSlots) + kSlots) % kSlots;   // q = q_base + t lands on slot `lane`
      bool have = feeder && t < n;
      uint32_t s = 0;
      RowInfo ri{};
      double y = 0.0;
      if (have) {
        s = eseq[t];
        ri = f.rows[s];
        y = f.yt[s];
      }
      while (__any_sync(0xffffffffu, have)) {
        const int64_t tn = t + kSlots;
        const bool have_n = feeder && tn < n;
        uint32_t sn = 0;
        RowInfo rin{};
        double yn = 0.0;
        if (have_n) {
          sn = eseq[tn];
          rin = f.rows[sn];
          yn = f.yt[sn];
        }
        bool pending = have;
        const int64_t q = q_base + t;
        const uint32_t par = static_cast<uint32_t>(((q / kSlots) & 1) ^ 1);
        while (__any_sync(0xffffffffu, pending)) {
          if (pending && mbar_try_wait(&ring.empty[lane], par)) {
            SpSlotMeta m;
            m.s = s;
            m.nnz = ri.nnz;
            m.start = ri.start;
            m.y = y;
            ring.meta[lane] = m;
            if (ri.nnz > 0 && ri.nnz <= kCap) {
              const uint32_t bi = static_cast<uint32_t>((ri.nnz + 3) / 4) * 16u;
              const uint32_t bv = static_cast<uint32_t>((ri.nnz + 1) / 2) * 16u;
              mbar_expect_tx(&ring.full[lane], bi + bv);
              bulk_g2s(ring.idx[lane], f.ci + ri.start, bi, &ring.full[lane]);
              bulk_g2s(ring.val[lane], f.cv + ri.start, bv, &ring.full[lane]);
            } else {
              mbar_arrive(&ring.full[lane]);
            }
            pending = false;
          }
        }
        t = tn;
        have = have_n;
        s = sn;
        ri = rin;
        y = yn;
      }
    } else if (warp == 0) {
      // ------------------------------------------------------------------ solver warp
      for (int64_t t = 0; t < n; ++t) {
        const int64_t q = q_base + t;
        const int slot = static_cast<int>(q % kSlots);
        const uint32_t t32 = static_cast<uint32_t>(t);
        mbar_wait(&ring.full[slot], static_cast<uint32_t>((q / kSlots) & 1));
        const SpSlotMeta m = ring.meta[slot];
        const double gm = f.gmem[m.s];
        double gch;

        if (m.nnz <= kCap) {
          // ---- gather the row and its coefficient state into registers
          int jr[kChunks];
          double vr[kChunks], wr[kChunks], gr[kChunks];
          uint32_t lr[kChunks];
#pragma unroll
          for (int c = 0; c < kChunks; ++c) {
            const int e = c * 32 + lane;
            jr[c] = (e < m.nnz) ? ring.idx[slot][e] : -1;
            vr[c] = (e < m.nnz) ? ring.val[slot][e] : 0.0;
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&ring.empty[slot]);     // slot can be refilled
#pragma unroll
          for (int c = 0; c < kChunks; ++c) {
            if (jr[c] >= 0) {
              wr[c] = W[jr[c]];
              gr[c] = G[jr[c]];
              lr[c] = lag[jr[c]];
            }
          }
          // ---- LaggedUpdate(k = t) and the sparse dot product
          const double step0 = gamma / wscale;
          double acc = 0.0;
#pragma unroll
          for (int c = 0; c < kChunks; ++c) {
            if (jr[c] >= 0) {
              const uint32_t lagged = t32 - lr[c];
              if (lagged != 0) {
                const double scal = lag_scale(identity, f.lag_scaling, lagged);
                PenCoef pc;
                pc.step = step0 * scal;
                pc.bgs = bg * scal;
                pc.thr = pc.bgs / wscale;
                pc.w_scale = wscale;
                wr[c] = penalty_scalar(pen, wr[c], gr[c], pc);
              }
              acc += vr[c] * wr[c];
            }
          }
          const double lp = warp_sum(acc) * wscale + b_reg;
          const double g = gradient_scalar(family, lp, m.y);
          gch = g - gm;
          if (lane == 0) f.gmem[m.s] = g;

          if (wscale < kSmall) {
            // rare: materialise the caught-up row, Reset(t) over all features, lag = t (src/saga-sparse.h:285-295)
#pragma unroll
            for (int c = 0; c < kChunks; ++c)
              if (jr[c] >= 0) {
                W[jr[c]] = wr[c];
                lag[jr[c]] = t32;
              }
            __syncwarp();
            for (int j = lane; j < p; j += 32) {
              const uint32_t lagged = t32 - lag[j];
              double w = W[j];
              if (lagged != 0)
                w = penalty_scalar(pen, w, G[j], pen_coef(gamma, beta, wscale, lag_scale(identity, f.lag_scaling, lagged)));
              W[j] = w * wscale;
              lag[j] = t32;
            }
            __syncwarp();
            wscale = 1.0;
#pragma unroll
            for (int c = 0; c < kChunks; ++c)
              if (jr[c] >= 0) wr[c] = W[jr[c]];
          }
          wscale *= r;
          if (fit_intercept) {
            gsi_reg += gch / nd;
            b_reg -= gamma * (gsi_reg * 0.01 + gch / nd);
          }
          // ---- AddWeighted(w), LaggedUpdate(k = t+1, lag 1), AddWeighted(g_sum), scatter
          const double sc = -gamma / wscale;
          PenCoef pc1;
          pc1.step = gamma / wscale * 1.0;
          pc1.bgs = bg;
          pc1.thr = bg / wscale;
          pc1.w_scale = wscale;
#pragma unroll
          for (int c = 0; c < kChunks; ++c) {
            if (jr[c] >= 0) {
              const double gx = vr[c] * gch;
              double w = wr[c] + gx * sc;
              w = penalty_scalar(pen, w, gr[c], pc1);
              W[jr[c]] = w;
              lag[jr[c]] = t32 + 1u;
              G[jr[c]] = gr[c] + gx * sc2;
            }
          }
        } else {
          // ---- long row: same operations, operands streamed from HBM in place (no ring copy was made)
          __syncwarp();
          if (lane == 0) mbar_arrive(&ring.empty[slot]);
          const int32_t* __restrict__ ci = f.ci + m.start;
          const double* __restrict__ cv = f.cv + m.start;
          double acc = 0.0;
          for (int e = lane; e < m.nnz; e += 32) {
            const int j = ci[e];
            const uint32_t lagged = t32 - lag[j];
            double w = W[j];
            if (lagged != 0) {
              w = penalty_scalar(pen, w, G[j], pen_coef(gamma, beta, wscale, lag_scale(identity, f.lag_scaling, lagged)));
              W[j] = w;
              lag[j] = t32;
            }
            acc += cv[e] * w;
          }
          const double lp = warp_sum(acc) * wscale + b_reg;
          const double g = gradient_scalar(family, lp, m.y);
          gch = g - gm;
          if (lane == 0) f.gmem[m.s] = g;
          if (wscale < kSmall) {
            __syncwarp();
            for (int j = lane; j < p; j += 32) {
              const uint32_t lagged = t32 - lag[j];
              double w = W[j];
              if (lagged != 0)
                w = penalty_scalar(pen, w, G[j], pen_coef(gamma, beta, wscale, lag_scale(identity, f.lag_scaling, lagged)));
              W[j] = w * wscale;
              lag[j] = t32;
            }
            __syncwarp();
            wscale = 1.0;
          }
          wscale *= r;
          if (fit_intercept) {
            gsi_reg += gch / nd;
            b_reg -= gamma * (gsi_reg * 0.01 + gch / nd);
          }
          const double sc = -gamma / wscale;
          const PenCoef pc1 = pen_coef(gamma, beta, wscale, 1.0);
          for (int e = lane; e < m.nnz; e += 32) {
            const int j = ci[e];
            const double gx = cv[e] * gch;
            const double gs = G[j];
            double w = W[j] + gx * sc;
            w = penalty_scalar(pen, w, gs, pc1);
            W[j] = w;
            lag[j] = t32 + 1u;
            G[j] = gs + gx * sc2;
          }
        }
        __syncwarp();   // order this update's scatter before the next update's gather (other lanes, same addresses)
      }
      if (lane == 0) wscale_s = wscale;
    }
    __syncthreads();

    // ---- epoch end: Reset(n), unscale, lag = 0, convergence
    wscale = wscale_s;
    double mc = 0.0, ms = 0.0;
    for (int j = tid; j < p; j += T) epoch_end_feature(f, 1, p, j, n32, wscale, pen, gamma, beta, identity, mc, ms);
    wscale = 1.0;
    const bool conv = block_converged(mc, ms, red, f.tol);
    ++it_outer;
    ++epochs_done;
    finished = !free_run && (conv || !(it_outer < f.max_iter));
  }

  if (tid == 0) {
    f.b[0] = b_reg;
    f.gsi[0] = gsi_reg;
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

// ============================================================================================ generic path
constexpr int kGenThreads = 256;

__global__ void __launch_bounds__(kGenThreads, 1)
saga_sparse_generic_kernel(FitDev* __restrict__ fits, Progress* __restrict__ prog, const RoundArgs* __restrict__ args) {
  __shared__ double red[2][2 * 32 * (kGenThreads / 32)];   // [parity][warp][2K]  (K <= 32)
  __shared__ double gch_s[32];
  __shared__ double cred[2 * (kGenThreads / 32)];

  const int fit_id = blockIdx.x;
  const RoundArgs ra = args[fit_id];
  Progress& pg = prog[fit_id];
  if (ra.n_epochs <= 0 || pg.status != kRunning) return;
  const bool free_run = (ra.flags & 1) != 0;
  const FitDev& f = fits[fit_id];

  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
  const int K = f.K, Ky = f.Ky, p = f.p;
  const int64_t n = f.n;
  const uint32_t n32 = static_cast<uint32_t>(n);
  const double nd = static_cast<double>(n32);
  const int family = f.family, pen = f.penalty;
  const bool fit_intercept = f.fit_intercept != 0, stdz = f.standardize != 0;

  const int li = pg.lambda_ind;
  const double gamma = f.gamma[li], alpha = f.alpha[li], beta = f.beta[li];
  const double r = 1.0 - alpha * gamma;
  const bool identity = (r == 1.0);
  const double sc2 = 1.0 / nd;

  double* __restrict__ W = f.W;
  double* __restrict__ G = f.gsum;
  uint32_t* __restrict__ lag = f.lag;
  const double* __restrict__ c = f.c;

  double b_reg = 0.0, gsi_reg = 0.0;
  if (warp == 0 && lane < K) {
    b_reg = f.b[lane];
    gsi_reg = f.gsi[lane];
  }
  double wscale = 1.0;
  uint32_t it_outer = pg.it_outer, epochs_done = 0;
  bool finished = false;

  for (int ep = 0; ep < ra.n_epochs && !finished; ++ep) {
    const uint32_t* __restrict__ eseq = ra.seq + size_t(ep) * n;
    for (int64_t t = 0; t < n; ++t) {
      const uint32_t t32 = static_cast<uint32_t>(t);
      const uint32_t s = eseq[t];
      const RowInfo ri = f.rows[s];
      const int32_t* __restrict__ ci = f.ci + ri.start;
      const double* __restrict__ cv = f.cv + ri.start;
      double* rb = red[t & 1];

      // 1. LaggedUpdate(k = t) on the row's features (thread e owns nonzero e, e+T, ...)
      for (int e = tid; e < ri.nnz; e += T) {
        const int j = ci[e];
        const uint32_t lagged = t32 - lag[j];
        if (lagged != 0) {
          apply_penalty(pen, W + j, G + j, K, p, pen_coef(gamma, beta, wscale, lag_scale(identity, f.lag_scaling, lagged)));
          lag[j] = t32;
        }
      }
      // 2. linear predictor. Sparse dot: one warp per class, 32 interleaved running sums over the row's nonzero
      //    positions + butterfly (sgdnet_arith.h); W.c (virtual centring): 256 interleaved sums over all features.
      __syncthreads();   // coefficients caught up by other threads are read below
      for (int k = warp; k < K; k += nwarps) {
        double a = 0.0;
        const double* Wk = W + size_t(k) * p;
        for (int e = lane; e < ri.nnz; e += 32) a += cv[e] * Wk[ci[e]];
        a = warp_sum(a);
        if (lane == 0) rb[k] = a;
      }
      if (stdz) {
        for (int k = 0; k < K; ++k) {
          double wc = 0.0;
          const double* Wk = W + size_t(k) * p;
          for (int j = tid; j < p; j += T) wc += Wk[j] * c[j];
          wc = warp_sum(wc);
          if (lane == 0) rb[32 + warp * 32 + k] = wc;
        }
      }
      __syncthreads();
      // 3. gradient, gradient memory, intercept (one lane per class)
      if (warp == 0) {
        const bool valid = lane < K;
        double lp = 0.0;
        double y_val = 0.0, gm = 0.0;
        if (valid) {
          const double a = rb[lane];
          double wc = 0.0;
          if (stdz)
            for (int w = 0; w < nwarps; ++w) wc += rb[32 + w * 32 + lane];
          lp = a * wscale + b_reg;
          if (stdz) lp -= wc * wscale;
          y_val = f.yt[size_t(s) * Ky + (Ky == 1 ? 0 : lane)];
          gm = f.gmem[size_t(s) * K + lane];
        }
        double g;
        if (family == kMultinomial) {
          const double yc = __shfl_sync(0xffffffffu, y_val, 0);
          const double lse = lse_warp(lp, valid);
          g = sgd_exp(lp - lse);
          if (static_cast<unsigned>(lane) == static_cast<unsigned>(yc + 0.5)) g -= 1.0;
        } else if (family == kBinomial) {
          g = 1.0 - y_val - 1.0 / (1.0 + sgd_exp(lp));
        } else {
          g = lp - y_val;
        }
        if (valid) {
          const double gch = g - gm;
          f.gmem[size_t(s) * K + lane] = g;
          if (fit_intercept) {
            gsi_reg += gch / nd;
            b_reg -= gamma * (gsi_reg * 0.01 + gch / nd);
          }
          gch_s[lane] = gch;
        }
      }
      if (wscale < kSmall) {       // uniform: every thread tracks the same wscale
        __syncthreads();
        for (int j = tid; j < p; j += T) {
          const uint32_t lagged = t32 - lag[j];
          if (lagged != 0)
            apply_penalty(pen, W + j, G + j, K, p, pen_coef(gamma, beta, wscale, lag_scale(identity, f.lag_scaling, lagged)));
          for (int k = 0; k < K; ++k) W[size_t(k) * p + j] *= wscale;
          lag[j] = t32;
        }
        wscale = 1.0;
      }
      wscale *= r;
      __syncthreads();
      // 4. AddWeighted(w)
      const double sc = -gamma / wscale;
      for (int e = tid; e < ri.nnz; e += T) {
        const int j = ci[e];
        for (int k = 0; k < K; ++k) W[size_t(k) * p + j] += cv[e] * gch_s[k] * sc;
      }
      if (stdz) {
        __syncthreads();
        for (int j = tid; j < p; j += T)
          for (int k = 0; k < K; ++k) W[size_t(k) * p + j] -= c[j] * gch_s[k] * sc;
        __syncthreads();
      }
      // 5. LaggedUpdate(k = t+1) then 6. AddWeighted(g_sum)
      const PenCoef pc1 = pen_coef(gamma, beta, wscale, lag_scale(identity, f.lag_scaling, 1u));
      for (int e = tid; e < ri.nnz; e += T) {
        const int j = ci[e];
        const uint32_t lagged = (t32 + 1u) - lag[j];
        if (lagged != 0) {
          if (lagged == 1u)
            apply_penalty(pen, W + j, G + j, K, p, pc1);
          else
            apply_penalty(pen, W + j, G + j, K, p, pen_coef(gamma, beta, wscale, lag_scale(identity, f.lag_scaling, lagged)));
          lag[j] = t32 + 1u;
        }
        for (int k = 0; k < K; ++k) G[size_t(k) * p + j] += cv[e] * gch_s[k] * sc2;
      }
      if (stdz) {
        __syncthreads();
        for (int j = tid; j < p; j += T)
          for (int k = 0; k < K; ++k) G[size_t(k) * p + j] -= c[j] * gch_s[k] * sc2;
      }
      __syncthreads();
    }

    double mc = 0.0, ms = 0.0;
    for (int j = tid; j < p; j += T) epoch_end_feature(f, K, p, j, n32, wscale, pen, gamma, beta, identity, mc, ms);
    wscale = 1.0;
    const bool conv = block_converged(mc, ms, cred, f.tol);
    ++it_outer;
    ++epochs_done;
    finished = !free_run && (conv || !(it_outer < f.max_iter));
  }

  if (warp == 0 && lane < K) {
    f.b[lane] = b_reg;
    f.gsi[lane] = gsi_reg;
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

cudaError_t launch_saga_sparse(int n_fits, bool fast_k1, FitDev* fits, Progress* prog, const RoundArgs* args,
                               cudaStream_t st) {
  if (fast_k1)
    saga_sparse_k1_kernel<<<n_fits, kSpThreads, 0, st>>>(fits, prog, args);
  else
    saga_sparse_generic_kernel<<<n_fits, kGenThreads, 0, st>>>(fits, prog, args);
  return cudaGetLastError();
}

}  // namespace sgd
