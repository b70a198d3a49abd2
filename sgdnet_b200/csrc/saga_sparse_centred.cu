// saga_sparse_centred.cu — sparse SAGA epochs with virtual centring (`standardize = TRUE` on sparse input), K == 1
// (reference: src/saga-sparse.h:114-130, 256-371).
//
// With virtual centring the reference touches EVERY coefficient on EVERY update: the linear predictor takes
// w . x_center_scaled over all p features (:276-277), and both AddWeighted calls subtract
// x_center_scaled * g_change * scaling from all of w and all of g_sum (:127-128). There is no exact lazy form of that
// (tests/test_arith_cpu.py: a running scalar applied once rounds differently from T subtractions), so the work per
// update is O(p) here as well; what this kernel changes is who does it and how often the block synchronises:
//
//   owner computes: thread tid of the one 256-thread CTA owns the features j = tid, tid + 256, ... - the partition the
//   arithmetic specification uses for the dense dot product w . c (256 interleaved running sums, butterfly per warp,
//   warps ascending) - and performs EVERY operation of an update on them, the sparse ones included, in the
//   reference's order. Which of its features are in the sampled row, and at which position, it reads from a position
//   map in shared memory that the row's lanes set one update ahead and the owners clear behind them.
//
//   one pass per update: everything update t does to feature j after its g_change is known (AddWeighted(w) sparse then
//   dense part, LaggedUpdate(k = t + 1), AddWeighted(g_sum) sparse then dense part) and everything update t + 1 does to
//   it before ITS gradient (LaggedUpdate(k = t + 1) on the next row's features, the products for the next sparse dot
//   product, the next partial sum of w . c) is one visit of the feature by its owner: state read once, written once.
//   Two block barriers per update (partial sums visible; g_change visible) instead of eight.
//
//   the row's products land in shared memory by nonzero position, and warp 0 adds them with the association of the
//   solver's sparse dot product (position e -> running sum e mod 32, xor-butterfly), then runs the scalar gradient
//   step; meanwhile the last warp moves the pipeline of the next rows on (sample index -> row descriptor and response
//   -> index / value runs by cp.async), each stage one update ahead of the next, so no warp ever waits for HBM.
//
// State (w, g_sum, c, lag, the two position maps) lives in shared memory for the whole launch when p allows it
// (p <= 6.5 k), otherwise in HBM / L2 with the same code. Rows longer than kCentCap nonzeros, K > 1: the generic kernel.
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

}  // namespace

size_t centred_smem_bytes(int p, bool* state_in_smem) {
  const size_t fixed = (sizeof(CentFixed) + 15) & ~size_t(15);
  const size_t state = size_t(p) * (3 * sizeof(double) + sizeof(uint32_t) + 2 * sizeof(uint16_t)) + 64;
  const bool in = fixed + state <= dense_smem_budget();
  if (state_in_smem) *state_in_smem = in;
  return in ? fixed + state : fixed;
}

template <bool SMEM>
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
  const bool identity = (r == 1.0);
  const double sc2 = 1.0 / nd;
  const double* __restrict__ ls_table = f.lag_scaling;

  // ---- state: shared memory (SMEM) or the fit's arrays in HBM
  double *W, *G;
  const double* C;
  uint32_t* lag;
  uint16_t* pos[2];
  if (SMEM) {
    unsigned char* base = cent_smem + ((sizeof(CentFixed) + 15) & ~size_t(15));
    W = reinterpret_cast<double*>(base);
    G = W + p;
    double* Cs = G + p;
    lag = reinterpret_cast<uint32_t*>(Cs + p);
    pos[0] = reinterpret_cast<uint16_t*>(lag + p);
    pos[1] = pos[0] + p;
    for (int j = tid; j < p; j += kCentThreads) {
      W[j] = f.W[j];
      G[j] = f.gsum[j];
      Cs[j] = f.c[j];
      lag[j] = 0u;
      pos[0][j] = 0;
      pos[1][j] = 0;
    }
    C = Cs;
  } else {
    W = f.W;
    G = f.gsum;
    C = f.c;
    lag = f.lag;
    pos[0] = pos_global;
    pos[1] = pos_global + p;
    for (int j = tid; j < p; j += kCentThreads) {
      lag[j] = 0u;
      pos[0][j] = 0;
      pos[1][j] = 0;
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
    if (tid < s0.nnz) pos[0][sm.ci[0][tid]] = static_cast<uint16_t>(tid + 1);
    if (tid == 0) gm_next = f.gmem[s0.s];
  }
  __syncthreads();

  // X(u_next): what update u_next does to an owned feature before its gradient; returns the feature's term of w . c
  auto visit_next = [&](int j, double& w, double g, double cj, const uint16_t* __restrict__ mapN, int slotN, uint32_t itN, double ws) {
    const uint32_t mN = mapN[j];
    if (mN != 0) {
      const uint32_t lagged = itN - lag[j];
      if (lagged != 0) {
        w = penalty_scalar(pen, w, g, pen_coef(gamma, beta, ws, lag_scale_c(identity, ls_table, lagged)));
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
      wc += visit_next(j, w, G[j], C[j], pos[0], 0, 0u, wscale);
      W[j] = w;
    }
    publish_wc(wc);
  }

  for (int ep = 0; ep < ra.n_epochs && !finished; ++ep) {
    for (uint32_t it = 0; it < n32; ++it, ++u) {
      const int slotC = static_cast<int>(u & (kCentRing - 1)), slotN = static_cast<int>((u + 1) & (kCentRing - 1));
      uint16_t* const mapC = pos[u & 1];
      uint16_t* const mapN = pos[(u + 1) & 1];
      const bool last_of_epoch = it + 1u == n32;
      const bool have_next = u + 1 < total;
      if (loader) cp_async_wait_all();               // row u + 1 (issued an update ago) is in its slot
      __syncthreads();                               // (1) products, warp sums of w . c, row u + 1 visible
      const CentSlot rowC = sm.slot[slotC];
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
          const double g = gradient_scalar(family, lp, rowC.y);
          const double gch = g - gm_next;
          f.gmem[rowC.s] = g;
          if (have_next) gm_next = f.gmem[rowN.s];     // behind the store in program order: sees it when the samples coincide
          sm.gch = gch;
          if (fit_intercept) {
            gsi_reg += gch / nd;
            b_reg -= gamma * (gsi_reg * 0.01 + gch / nd);
          }
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
      const bool reset = wscale < kSmall;              // src/saga-sparse.h:285-295, decided on wscale before this step
      const double ws_old = wscale;
      wscale = (reset ? 1.0 : wscale) * r;
      const double sc = -gamma / wscale;
      const PenCoef pc1 = pen_coef(gamma, beta, wscale, lag_scale_c(identity, ls_table, 1u));
      __syncthreads();                               // (2) g_change and the next row's map visible
      const double gch = sm.gch;

      // ---- one visit per owned feature: the rest of update u, then (unless the epoch ends) the head of update u + 1
      double wc = 0.0, mc = 0.0, ms = 0.0;
      for (int j = tid; j < p; j += kCentThreads) {
        double w = W[j], g = G[j];
        const double cj = C[j];
        const uint32_t mC = mapC[j];
        if (reset) {                                   // Reset(it): catch up, fold the scale in, lag = it
          const uint32_t lagged = it - lag[j];
          if (lagged != 0) w = penalty_scalar(pen, w, g, pen_coef(gamma, beta, ws_old, lag_scale_c(identity, ls_table, lagged)));
          w *= ws_old;
          lag[j] = it;
        }
        if (mC != 0) w += sm.cv[slotC][mC - 1] * gch * sc;          // AddWeighted(w): the row's part ...
        w -= cj * gch * sc;                                          // ... and the centring part (all features)
        if (mC != 0) {                                               // LaggedUpdate(k = it + 1): the row's features lag by one
          w = penalty_scalar(pen, w, g, pc1);
          lag[j] = it + 1u;
          g += sm.cv[slotC][mC - 1] * gch * sc2;                     // AddWeighted(g_sum): the row's part ...
          mapC[j] = 0;
        }
        g -= cj * gch * sc2;                                         // ... and the centring part
        if (last_of_epoch) {
          // Reset(n) + unscale + convergence bookkeeping (src/saga-sparse.h:340-348, 367; src/utils.h:240-262)
          const uint32_t lagged = n32 - lag[j];
          if (lagged != 0) w = penalty_scalar(pen, w, g, pen_coef(gamma, beta, wscale, lag_scale_c(identity, ls_table, lagged)));
          w *= wscale;
          lag[j] = 0u;
          mc = fmax(mc, fabs(w - f.Wprev[j]));
          ms = fmax(ms, fabs(w));
          f.Wprev[j] = w;
        } else if (have_next) {
          wc += visit_next(j, w, g, cj, mapN, slotN, it + 1u, wscale);
        }
        W[j] = w;
        G[j] = g;
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

cudaError_t launch_saga_sparse_centred(int p, FitDev* fit, Progress* prog, const RoundArgs& ra, uint16_t* pos_global, cudaStream_t st) {
  bool in_smem = false;
  const size_t smem = centred_smem_bytes(p, &in_smem);
  cudaError_t e;
  if (in_smem) {
    e = cudaFuncSetAttribute(saga_sparse_centred_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    saga_sparse_centred_kernel<true><<<1, kCentThreads, smem, st>>>(fit, prog, ra, pos_global);
  } else {
    e = cudaFuncSetAttribute(saga_sparse_centred_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    saga_sparse_centred_kernel<false><<<1, kCentThreads, smem, st>>>(fit, prog, ra, pos_global);
  }
  return cudaGetLastError();
}

}  // namespace sgd
