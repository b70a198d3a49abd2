// saga_sparse_centred.cu — sparse SAGA epochs with virtual centring (`standardize = TRUE` on sparse input), K == 1
// (reference: src/saga-sparse.h:114-130, 256-371).
//
// With virtual centring the reference touches EVERY coefficient on EVERY update: the linear predictor takes
// w . x_center_scaled over all p features (:276-277), and both AddWeighted calls subtract
// x_center_scaled * g_change * scaling from all of w and all of g_sum (:127-128). There is no exact lazy form of that
// (tests/test_arith_cpu.py: a running scalar applied once rounds differently from T subtractions), so the work per
// update is O(p) here as well. What this kernel organises is who does which part of it:
//
//   owners: thread tid of the one 256-thread CTA owns the features j = tid, tid + 256, ... - the partition the
//   arithmetic specification uses for the dense dot product w . c (256 interleaved running sums, butterfly per warp,
//   warps ascending). Once per update an owner visits its features, eight independent chains at a time: the centring
//   parts of the two AddWeighted calls of update t (state read once, written once) and the feature's term of the
//   w . c of update t + 1, branch-free.
//
//   row lanes: the few features of the sampled rows carry the sparse operations in between (AddWeighted row part,
//   LaggedUpdate, the product for the sparse dot) - a dependent chain of a dozen FP64 operations per feature that an
//   owner would execute once per feature, one after the other, under divergence. They are taken out of the owners'
//   pass and done one lane per nonzero position, all positions in parallel: R1 completes update t on the features of
//   row t; R2 finishes update t and does the head of update t + 1 (LaggedUpdate(k = t + 1), product) on the features
//   of row t + 1. Position maps in shared memory (set by the row's lanes one update ahead, cleared by the owners) tell
//   an owner which features a row lane has handled. Per element the sequence of floating point operations is the
//   reference's.
//
//   warp 0 adds the row's products with the association of the solver's sparse dot product (position e -> running
//   sum e mod 32, xor-butterfly) and runs the scalar gradient step; one lane of warp 1 prepares the update's step
//   constants (the divisions by wscale); the last warp moves the pipeline of the next rows on (sample index -> row
//   descriptor and response -> index / value runs by cp.async), each stage one update ahead of the next, so no warp
//   waits for HBM. Four block barriers per update.
//
// State (w, g_sum, c, lag, the two position maps) lives in shared memory for the whole launch when p allows it
// (p <= 6.5 k; the owners' pass then moves 44 B per feature through shared memory, which is what bounds it), otherwise in
// HBM / L2 with the same code. Rows longer than kCentCap nonzeros and K > 1 stay on saga_sparse_generic_kernel.
#include "common.cuh"
#include "kernels.h"

namespace sgd {

namespace {

constexpr int kCentThreads = 256;   // tied to the association of w . c (256 interleaved sums, 8 warps ascending)
constexpr int kCentWarps = kCentThreads / 32;
constexpr int kCentRing = 4;        // rows staged ahead (power of two)

struct CentSlot {
  uint32_t s;
  int32_t nnz;
  double y;
};

struct __align__(16) CentFixed {
  double cv[kCentRing][kCentCap];
  int32_t ci[kCentRing][kCentCap];
  double prod[kCentCap];            // the current row's products cv[e] * w[ci[e]], by nonzero position
  CentSlot slot[kCentRing];
  double red[kCentWarps];           // warp sums of w . c
  double cred[2 * kCentWarps];
  double gch;                       // g_change of the current update
  double sc, step1, thr1, bgs1;     // the update's step constants (uniform; computed once, by one lane, off warp 0's path)
};

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Block-wide max of (mc, ms) and the convergence decision (src/utils.h:240-262); `red` holds 2 * kCentWarps doubles.
__device__ __forceinline__ bool block_converged(double mc, double ms, double* red, double tol) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  mc = warp_max(mc);
  ms = warp_max(ms);
  if (lane == 0) {
    red[warp] = mc;
    red[kCentWarps + warp] = ms;
  }
  __syncthreads();
  double mc_all = 0.0, ms_all = 0.0;
  for (int w = 0; w < kCentWarps; ++w) {
    mc_all = fmax(mc_all, red[w]);
    ms_all = fmax(ms_all, red[kCentWarps + w]);
  }
  __syncthreads();
  const bool all_zero = (ms_all == 0.0) && (mc_all == 0.0);
  const bool no_change = (ms_all != 0.0) && (mc_all / ms_all <= tol);
  return all_zero || no_change;
}

__device__ __forceinline__ double lag_scale_c(bool identity, const double* __restrict__ table, uint32_t m) {
  return identity ? static_cast<double>(m) : table[m];
}

// LaggedUpdate of one feature that lags by `lagged` steps (src/saga-sparse.h:76-100; penalties.h with scaling =
// lag_scaling[lagged]). Out of line: it is taken by a handful of lanes per update and carries two divisions, and the
// pass over the owned features has to stay small enough for the instruction cache.
__device__ __noinline__ double centred_catch_up(int pen, double w, double g, double gamma, double beta, double ws, double scaling) {
  return penalty_scalar(pen, w, g, pen_coef(gamma, beta, ws, scaling));
}

#ifdef SGD_CENT_TRACE
// measurement build: cycles thread 0 spends in the phases of an update, summed over a launch
__device__ long long g_cent_trace[8];
__device__ __forceinline__ long long cent_clock(double dep) {      // a clock read that depends on `dep` having arrived
  long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "d"(dep) : "memory");
  return t;
}
#define CT(i, dep) do { if (tid == 0) { const long long t_ = cent_clock(dep); ct_acc[i] += t_ - ct_last; ct_last = t_; } } while (0)
#else
#define CT(i, dep)
#endif

}  // namespace

#ifdef SGD_CENT_TRACE
extern "C" void sgdnet_debug_centred_trace(long long* out) { cudaMemcpyFromSymbol(out, g_cent_trace, sizeof(g_cent_trace)); }
#endif

size_t centred_smem_bytes(int p, bool* state_in_smem) {
  const size_t fixed = (sizeof(CentFixed) + 15) & ~size_t(15);
  const size_t state = size_t(p) * (3 * sizeof(double) + sizeof(uint32_t) + 2 * sizeof(uint16_t)) + 64;
  const bool in = fixed + state <= dense_smem_budget();
  if (state_in_smem) *state_in_smem = in;
  return in ? fixed + state : fixed;
}

// MODE: 0 elastic net with alpha gamma == 0 (the lasso: wscale stays exactly 1, lag_scaling[m] == m), 1 elastic net,
// 2 ridge, 3 any penalty through the general functors - compile-time so that the pass over the owned features carries
// only the arithmetic its fit needs.
template <bool SMEM, int MODE>
__global__ void __launch_bounds__(kCentThreads, 1)
saga_sparse_centred_kernel(FitDev* __restrict__ fit, Progress* __restrict__ prog, const RoundArgs ra, uint16_t* __restrict__ pos_global) {
  extern __shared__ __align__(16) unsigned char cent_smem[];
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
  CentFixed& sm = *reinterpret_cast<CentFixed*>(cent_smem);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int p = f.p;
  const int64_t n = f.n;
  const uint32_t n32 = static_cast<uint32_t>(n);
  const double nd = static_cast<double>(n32);
  const int family = f.family, pen = f.penalty;
  const bool fit_intercept = f.fit_intercept != 0;
  const int li = pg.lambda_ind;
  const double gamma = f.gamma[li], alpha = f.alpha[li], beta = f.beta[li];
  const double r = 1.0 - alpha * gamma;
  const bool identity = (MODE == 0) ? true : ((MODE == 3) ? (r == 1.0) : false);
  const double sc2 = 1.0 / nd, rn = 1.0 / nd;
  const double* __restrict__ ls_table = f.lag_scaling;

  // ---- state: shared memory (SMEM) or the fit's arrays in HBM
  double *W, *G;
  const double* C;
  uint32_t* lag;
  uint16_t *pos0, *pos1;
  if (SMEM) {
    unsigned char* base = cent_smem + ((sizeof(CentFixed) + 15) & ~size_t(15));
    W = reinterpret_cast<double*>(base);
    G = W + p;
    double* Cs = G + p;
    lag = reinterpret_cast<uint32_t*>(Cs + p);
    pos0 = reinterpret_cast<uint16_t*>(lag + p);
    pos1 = pos0 + p;
    for (int j = tid; j < p; j += kCentThreads) {
      W[j] = f.W[j];
      G[j] = f.gsum[j];
      Cs[j] = f.c[j];
      lag[j] = 0u;
      pos0[j] = 0;
      pos1[j] = 0;
    }
    C = Cs;
  } else {
    W = f.W;
    G = f.gsum;
    C = f.c;
    lag = f.lag;
    pos0 = pos_global;
    pos1 = pos_global + p;
    for (int j = tid; j < p; j += kCentThreads) {
      lag[j] = 0u;
      pos0[j] = 0;
      pos1[j] = 0;
    }
  }

  const uint32_t* __restrict__ seq = ra.seq;
  const int64_t total = n * ra.n_epochs;

  // ---- row pipeline (last warp): row u lives in ring slot u % kCentRing
  const bool loader = warp == kCentWarps - 1;
  auto copy_row = [&](int64_t u, const RowInfo& ri, uint32_t s, double y) {      // the loader warp; asynchronous
    const int slot = static_cast<int>(u & (kCentRing - 1));
    const int32_t* src_i = f.ci + ri.start;
    const double* src_v = f.cv + ri.start;
    for (int q = lane; q < (ri.nnz + 3) / 4; q += 32) cp_async16(&sm.ci[slot][4 * q], src_i + 4 * q);
    for (int q = lane; q < (ri.nnz + 1) / 2; q += 32) cp_async16(&sm.cv[slot][2 * q], src_v + 2 * q);
    if (lane == 0) sm.slot[slot] = CentSlot{s, ri.nnz, y};
  };
  // loader registers: the row two ahead has its descriptor, the row three ahead its sample index
  uint32_t s_d = 0, s_i = 0;          // samples of rows u + 2 (descriptor loaded) and u + 3 (index loaded), relative to the loop
  RowInfo ri_d{};
  double y_d = 0.0;
  if (loader) {
    // rows 0 and 1 synchronously, row 2's descriptor, row 3's sample index
    for (int64_t u = 0; u < 2 && u < total; ++u) {
      const uint32_t s = seq[u];
      copy_row(u, f.rows[s], s, f.yt[s]);
    }
    cp_async_wait_all();
    if (2 < total) {
      s_d = seq[2];
      ri_d = f.rows[s_d];
      y_d = f.yt[s_d];
    }
    if (3 < total) s_i = seq[3];
  }

  double b_reg = f.b[0], gsi_reg = f.gsi[0];      // used by thread 0
  double gm_next = 0.0;
  double wscale = 1.0;
  uint32_t it_outer = pg.it_outer, epochs_done = 0;
  bool finished = false;
  int64_t u = 0;                                   // update index within the launch
  __syncthreads();

  // position map of row 0, then X(0): products and partial w . c for the first update
  {
    const CentSlot s0 = sm.slot[0];
    if (tid < s0.nnz) pos0[sm.ci[0][tid]] = static_cast<uint16_t>(tid + 1);
    if (tid == 0) gm_next = f.gmem[s0.s];
  }
  __syncthreads();

  // X(u_next): what update u_next does to an owned feature before its gradient; returns the feature's term of w . c
  constexpr bool en = MODE == 0 || MODE == 1;
  constexpr bool plain = MODE != 3;
  const double bg = beta * gamma;
  // LaggedUpdate of one feature that lags by `lagged` steps (src/saga-sparse.h:76-100; the functors of penalties.h with
  // scaling = lag_scaling[lagged]): step = gamma / w_scale * scaling, threshold = (beta gamma scaling) / w_scale. When
  // alpha gamma == 0 (lasso, lambda == 0) w_scale stays exactly 1: both divisions are by 1.0, i.e. exact identities.
  auto catch_up = [&](double w, double g, uint32_t lagged, double ws) {
    const double scal = lag_scale_c(identity, ls_table, lagged);
    if (plain) {
      const double step = identity ? gamma * scal : gamma / ws * scal;
      const double v = w - step * g;
      if (!en) return v;
      const double bgs = bg * scal;
      return soft_threshold(v, identity ? bgs : bgs / ws);
    }
    return centred_catch_up(pen, w, g, gamma, beta, ws, scal);
  };
  // X(u_next): what update u_next does to an owned feature before its gradient; returns the feature's term of w . c
  auto visit_next = [&](int j, double& w, double g, double cj, const uint16_t* __restrict__ mapN, int slotN, uint32_t itN, double ws) {
    const uint32_t mN = mapN[j];
    if (mN != 0) {
      const uint32_t lagged = itN - lag[j];
      if (lagged != 0) {
        w = catch_up(w, g, lagged, ws);
        lag[j] = itN;
      }
      sm.prod[mN - 1] = sm.cv[slotN][mN - 1] * w;
    }
    return w * cj;
  };
  auto publish_wc = [&](double wc) {
    wc = warp_sum(wc);
    if (lane == 0) sm.red[warp] = wc;
  };
  {
    double wc = 0.0;
    for (int j = tid; j < p; j += kCentThreads) {
      double w = W[j];
      wc += visit_next(j, w, G[j], C[j], pos0, 0, 0u, wscale);
      W[j] = w;
    }
    publish_wc(wc);
  }

#ifdef SGD_CENT_TRACE
  long long ct_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ct_last = clock64();
#endif
  for (int ep = 0; ep < ra.n_epochs && !finished; ++ep) {
    for (uint32_t it = 0; it < n32; ++it, ++u) {
      const int slotC = static_cast<int>(u & (kCentRing - 1)), slotN = static_cast<int>((u + 1) & (kCentRing - 1));
      uint16_t* const mapC = (u & 1) ? pos1 : pos0;
      uint16_t* const mapN = (u & 1) ? pos0 : pos1;
      const bool last_of_epoch = it + 1u == n32;
      const bool have_next = u + 1 < total;
      CT(0, 0.0);                                    // pass over the owned features (previous update) + loop overhead
      if (loader) cp_async_wait_all();               // row u + 1 (issued an update ago) is in its slot
      __syncthreads();                               // (1) products, warp sums of w . c, row u + 1 visible
      const CentSlot rowC = sm.slot[slotC];
      CT(1, rowC.y);                                 // barrier 1
      const CentSlot rowN = have_next ? sm.slot[slotN] : CentSlot{0u, 0, 0.0};
      if (warp == 0) {
        // ---- sparse dot product (position e -> running sum e mod 32, butterfly), w . c, gradient, intercept
        double a = 0.0;
        for (int e = lane; e < rowC.nnz; e += 32) a += sm.prod[e];
        a = warp_sum(a);
        if (lane == 0) {
          double wc = 0.0;
#pragma unroll
          for (int w8 = 0; w8 < kCentWarps; ++w8) wc += sm.red[w8];
          double lp = a * wscale + b_reg;
          lp -= wc * wscale;
          const double g = (family == kBinomial) ? 1.0 - rowC.y - 1.0 / (1.0 + sgd_exp_inrange(lp)) : lp - rowC.y;
          const double gch = g - gm_next;
          f.gmem[rowC.s] = g;
          if (have_next) gm_next = f.gmem[rowN.s];     // behind the store in program order: sees it when the samples coincide
          sm.gch = gch;
          if (fit_intercept) {
            const double gn = div_by_n(gch, nd, rn);      // the bits of gch / nd
            gsi_reg += gn;
            b_reg -= gamma * (gsi_reg * 0.01 + gn);
          }
        }
      } else if (warp == 1) {
        // ---- this update's step constants (functions of the deterministic wscale track): four divisions, once
        if (lane == 0) {
          const double ws_new = ((wscale < kSmall) ? 1.0 : wscale) * r;
          const PenCoef pc = pen_coef(gamma, beta, ws_new, lag_scale_c(identity, ls_table, 1u));
          sm.sc = -gamma / ws_new;
          sm.step1 = pc.step;
          sm.thr1 = pc.thr;
          sm.bgs1 = pc.bgs;
        }
      } else if (loader) {
        // ---- pipeline: copy row u + 2, descriptor of row u + 3, sample index of row u + 4
        if (u + 2 < total) copy_row(u + 2, ri_d, s_d, y_d);
        if (u + 3 < total) {
          s_d = s_i;
          ri_d = f.rows[s_d];
          y_d = f.yt[s_d];
        }
        if (u + 4 < total) s_i = seq[u + 4];
      }
      // position map of row u + 1 (its lanes; the owners read it in the pass below)
      if (have_next && !last_of_epoch && tid < rowN.nnz) mapN[sm.ci[slotN][tid]] = static_cast<uint16_t>(tid + 1);
      CT(2, b_reg);                                  // warp 0: dot products, gradient, intercept
      const bool reset = wscale < kSmall;              // src/saga-sparse.h:285-295, decided on wscale before this step
      const double ws_old = wscale;
      wscale = (reset ? 1.0 : wscale) * r;
      __syncthreads();                               // (2) g_change, step constants and the next row's map visible
      const double gch = sm.gch;
      CT(3, gch);                                    // barrier 2
      const double sc = sm.sc;
      PenCoef pc1;
      pc1.step = sm.step1;
      pc1.thr = sm.thr1;
      pc1.bgs = sm.bgs1;
      pc1.w_scale = wscale;

      // ---- one visit per owned feature: the rest of update u, then (unless the epoch ends) the head of update u + 1
      double wc = 0.0, mc = 0.0, ms = 0.0;
      if (reset) {                                     // Reset(it), rare: catch up, fold the scale in, lag = it
        for (int j = tid; j < p; j += kCentThreads) {
          double w = W[j];
          const uint32_t lagged = it - lag[j];
          if (lagged != 0) w = catch_up(w, G[j], lagged, ws_old);
          W[j] = w * ws_old;
          lag[j] = it;
        }
        __syncthreads();
      }
      const bool head_next = !last_of_epoch && have_next;
      // ---- R1: the features of this update's row, one lane per nonzero position (an owner would meet them one after
      // the other, each a dependent chain of a dozen FP64 operations): AddWeighted(w) row part then centring part,
      // LaggedUpdate(k = it + 1), AddWeighted(g_sum) row part then centring part
      if (tid < rowC.nnz) {
        const int j = sm.ci[slotC][tid];
        const double g0 = G[j];
        const double cg = C[j] * gch, xg = sm.cv[slotC][tid] * gch;
        double w = W[j] + xg * sc;
        w -= cg * sc;
        if (plain) {
          const double v = w - pc1.step * g0;
          w = en ? soft_threshold(v, pc1.thr) : v;
        } else {
          w = penalty_scalar(pen, w, g0, pc1);
        }
        lag[j] = it + 1u;
        double g = g0 + xg * sc2;
        g -= cg * sc2;
        W[j] = w;
        G[j] = g;
      }
      __syncthreads();                               // (3)
      // ---- R2: the features of the next row, one lane per position: what is left of this update for them (the
      // centring parts, unless R1 just did the whole update), then LaggedUpdate(k = it + 1) and the product for the
      // next sparse dot product
      if (head_next && tid < rowN.nnz) {
        const int j = sm.ci[slotN][tid];
        double w = W[j], g = G[j];
        if (mapC[j] == 0) {
          const double cg = C[j] * gch;
          w -= cg * sc;
          g -= cg * sc2;
          G[j] = g;
        }
        const uint32_t lagged = (it + 1u) - lag[j];
        if (lagged != 0) {
          w = catch_up(w, g, lagged, wscale);
          lag[j] = it + 1u;
        }
        W[j] = w;
        sm.prod[tid] = sm.cv[slotN][tid] * w;
      }
      __syncthreads();                               // (4)
      // ---- O: every owner over its features, kCh independent chains at a time: the centring parts of the two
      // AddWeighted calls for the features no row lane handled, and the feature's term of the next w . c (ascending j)
      constexpr int kCh = 8;
      for (int j0 = tid; j0 < p; j0 += kCh * kCentThreads) {
        double w[kCh], g[kCh], cj[kCh];
        bool done[kCh];
#pragma unroll
        for (int i = 0; i < kCh; ++i) {
          const int j = j0 + i * kCentThreads;
          const bool in = j < p;
          w[i] = in ? W[j] : 0.0;
          g[i] = in ? G[j] : 0.0;
          cj[i] = in ? C[j] : 0.0;
          const uint32_t mC = in ? mapC[j] : 0u;
          const uint32_t mN = (in && head_next) ? mapN[j] : 0u;
          done[i] = (mC | mN) != 0;
          if (mC != 0) mapC[j] = 0;
        }
#pragma unroll
        for (int i = 0; i < kCh; ++i) {
          const double cg = cj[i] * gch;
          const double w2 = w[i] - cg * sc, g2 = g[i] - cg * sc2;
          w[i] = done[i] ? w[i] : w2;
          g[i] = done[i] ? g[i] : g2;
        }
#pragma unroll
        for (int i = 0; i < kCh; ++i) {
          const int j = j0 + i * kCentThreads;
          if (j < p) {
            if (head_next) wc += w[i] * cj[i];
            W[j] = w[i];
            G[j] = g[i];
          }
        }
      }
      if (last_of_epoch) {
        // Reset(n) + unscale + convergence bookkeeping (src/saga-sparse.h:340-348, 367; src/utils.h:240-262)
        for (int j = tid; j < p; j += kCentThreads) {
          double w = W[j];
          const uint32_t lagged = n32 - lag[j];
          if (lagged != 0) w = catch_up(w, G[j], lagged, wscale);
          w *= wscale;
          lag[j] = 0u;
          mc = fmax(mc, fabs(w - f.Wprev[j]));
          ms = fmax(ms, fabs(w));
          f.Wprev[j] = w;
          W[j] = w;
        }
      }
      if (!last_of_epoch) publish_wc(wc);
      if (last_of_epoch) {
        wscale = 1.0;
        const bool conv = block_converged(mc, ms, sm.cred, f.tol);
        ++it_outer;
        ++epochs_done;
        finished = !free_run && (conv || !(it_outer < f.max_iter));
        if (!finished && ep + 1 < ra.n_epochs && have_next) {
          // head of the next epoch's first update on the unscaled coefficients (every lag is 0: nothing to catch up)
          if (tid < rowN.nnz) mapN[sm.ci[slotN][tid]] = static_cast<uint16_t>(tid + 1);
          __syncthreads();
          double wc0 = 0.0;
          for (int j = tid; j < p; j += kCentThreads) {
            double w = W[j];
            wc0 += visit_next(j, w, G[j], C[j], mapN, slotN, 0u, wscale);
            W[j] = w;
          }
          publish_wc(wc0);
        }
      }
    }
  }

#ifdef SGD_CENT_TRACE
  if (tid == 0)
    for (int i = 0; i < 8; ++i) g_cent_trace[i] = ct_acc[i];
#endif
  if (loader) cp_async_wait_all();
  __syncthreads();
  if (SMEM) {
    for (int j = tid; j < p; j += kCentThreads) {
      f.W[j] = W[j];
      f.gsum[j] = G[j];
    }
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
    pg.solver_ns += globaltimer_ns() - t_start;
    __threadfence();
    publish_progress(f.mirror, pg, ra.round_id);
  }
}

template <bool SMEM, int MODE>
static cudaError_t launch_centred_variant(size_t smem, FitDev* fit, Progress* prog, const RoundArgs& ra, uint16_t* pos_global, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(saga_sparse_centred_kernel<SMEM, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  saga_sparse_centred_kernel<SMEM, MODE><<<1, kCentThreads, smem, st>>>(fit, prog, ra, pos_global);
  return cudaGetLastError();
}

// mode: 0 lasso-like (elastic net, the path's alpha * gamma == 0 at every lambda), 1 elastic net, 2 ridge, 3 general
cudaError_t launch_saga_sparse_centred(int p, int mode, FitDev* fit, Progress* prog, const RoundArgs& ra, uint16_t* pos_global, cudaStream_t st) {
  bool in_smem = false;
  const size_t smem = centred_smem_bytes(p, &in_smem);
  if (in_smem) {
    switch (mode) {
      case 0: return launch_centred_variant<true, 0>(smem, fit, prog, ra, pos_global, st);
      case 1: return launch_centred_variant<true, 1>(smem, fit, prog, ra, pos_global, st);
      case 2: return launch_centred_variant<true, 2>(smem, fit, prog, ra, pos_global, st);
      default: return launch_centred_variant<true, 3>(smem, fit, prog, ra, pos_global, st);
    }
  }
  switch (mode) {
    case 0: return launch_centred_variant<false, 0>(smem, fit, prog, ra, pos_global, st);
    case 1: return launch_centred_variant<false, 1>(smem, fit, prog, ra, pos_global, st);
    case 2: return launch_centred_variant<false, 2>(smem, fit, prog, ra, pos_global, st);
    default: return launch_centred_variant<false, 3>(smem, fit, prog, ra, pos_global, st);
  }
}

}  // namespace sgd
