// saga_dense_cluster_generic.cu — dense SAGA epochs on a thread-block cluster (reference: src/saga-dense.h:147-212):
// the any-shape form (run-time slice count, coefficient state in shared memory when it fits, otherwise in HBM / L2) that
// takes the wide designs saga_dense_cluster.cu has no compile-time instantiation for (p > 8192, or K * p too large for
// the eight CTAs' shared memory).
//
// One cluster of 8 CTAs x 256 threads per fit = 2048 lanes. Lane L = 256 * cta + tid owns the features j = L, L + 2048,
// ... for ALL classes: their W and g_sum live in that CTA's shared memory for the whole launch and are touched by no
// other thread, so the dense sweeps of the reference's update (the K x p matrix-vector product, the coefficient step,
// the prox over all p features and the g_sum update, src/saga-dense.h:154, 176-183) are split 2048 ways. What crosses
// threads per update is only the K partial dot products:
//   lane: running sum over its features (ascending j)  ->  xor-butterfly inside each warp  ->  the CTA's 8 warp sums
//   added in ascending order  ->  the 8 CTA sums exchanged all-to-all through DISTRIBUTED SHARED MEMORY (each CTA
//   stores its K values into the other CTAs' shared memory and arrives on their mbarrier with release.cluster)  ->
//   every CTA adds the 8 CTA sums in ascending order and runs the (K-value) gradient step redundantly, so no second
//   exchange is needed; CTA 0 alone stores the gradient memory.
// That association is the arithmetic specification of a wide dense dot product (include/sgdnet_arith.h, item 2) and is
// what the oracle's portable mode computes (oracle/sgdnet_oracle.cpp dot_dense_wide).
// A CTA streams only ITS 256-feature slices of each sampled row: 1-D bulk copies (cp.async.bulk -> UBLKCP) into a
// shared-memory ring, kRing - 1 rows ahead, driven by the sampling sequence.
// Algorithmic HBM bytes per update: 8*p (row) + 4 (index) + 8*K_y (y) + 16*K (gradient memory read + write).
#include <cooperative_groups.h>

#include "common.cuh"
#include "kernels.h"

namespace sgd {

namespace {

constexpr int kCT = 256;        // feature lanes per CTA (warps 1..8); warp 0 is the control warp
constexpr int kCBlock = kCT + 32;
constexpr int kCluster = 8;     // CTAs per fit
constexpr int kLanes = kCT * kCluster;
constexpr int kCRing = 4;       // row ring depth

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void st_cluster_f64(uint32_t addr, double v) {
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
// remote store that completes bytes on a (remote) mbarrier: data and signal travel together, no fence on either side
// (an mbarrier.arrive.release.cluster compiles to MEMBAR.ALL.GPU and its acquire side to an L1 invalidation, CCTL.IVALL -
// together 45 % of the control warp's time when the exchange used them)
__device__ __forceinline__ void st_async_f64(uint32_t addr, double v, uint32_t mbar_addr) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];" ::"r"(addr), "d"(v), "r"(mbar_addr)
               : "memory");
}
__device__ __forceinline__ void arrive_cluster(uint32_t bar_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ void wait_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}


struct ClusterSmem {
  double* ring;      // [kCRing][nch][kCT]
  double* W;         // [K][nch * kCT] (or null: state stays in HBM / L2)
  double* G;
  double* part;      // [2][kCluster + 1][32]: CTA sums of both parities from every CTA; row kCluster: gradient memory from CTA 0
  double* red;       // [8 warps][32]
  double* gch;       // [32]
  double* conv;      // [2][kCluster][2]: epoch-end maxima from every CTA
  uint64_t* full;    // [kCRing]
  uint64_t* pbar;    // [2] partial sums arrived (count 1 + transaction bytes: (kCluster + 1) * K doubles per update)
  uint64_t* cbar;    // [2] epoch-end maxima arrived (count kCluster)
};

__host__ __device__ inline size_t cluster_fixed_bytes(int nch) {
  return sizeof(double) * (size_t(kCRing) * nch * kCT + 2 * (kCluster + 1) * 32 + 8 * 32 + 32 + 2 * kCluster * 2) +
         sizeof(uint64_t) * (kCRing + 4);
}

__device__ __forceinline__ ClusterSmem carve_cluster(unsigned char* base, int K, int nch, bool state_in_smem) {
  ClusterSmem s;
  size_t off = 0;
  s.ring = reinterpret_cast<double*>(base + off); off += sizeof(double) * size_t(kCRing) * nch * kCT;
  s.part = reinterpret_cast<double*>(base + off); off += sizeof(double) * 2 * (kCluster + 1) * 32;
  s.red = reinterpret_cast<double*>(base + off); off += sizeof(double) * 8 * 32;
  s.gch = reinterpret_cast<double*>(base + off); off += sizeof(double) * 32;
  s.conv = reinterpret_cast<double*>(base + off); off += sizeof(double) * 2 * kCluster * 2;
  s.full = reinterpret_cast<uint64_t*>(base + off); off += sizeof(uint64_t) * kCRing;
  s.pbar = reinterpret_cast<uint64_t*>(base + off); off += sizeof(uint64_t) * 2;
  s.cbar = reinterpret_cast<uint64_t*>(base + off); off += sizeof(uint64_t) * 2;
  off = (off + 15) & ~size_t(15);
  if (state_in_smem) {
    s.W = reinterpret_cast<double*>(base + off); off += sizeof(double) * size_t(K) * nch * kCT;
    s.G = reinterpret_cast<double*>(base + off);
  } else {
    s.W = nullptr;
    s.G = nullptr;
  }
  return s;
}

}  // namespace

size_t dense_cluster_generic_smem_bytes(int K, int p) {
  const int nch = (p + kLanes - 1) / kLanes;
  const size_t fixed = (cluster_fixed_bytes(nch) + 15) & ~size_t(15);
  const size_t state = sizeof(double) * 2 * size_t(K) * nch * kCT;
  return (fixed + state <= dense_smem_budget()) ? fixed + state : fixed;
}

// KT: class-count bucket (1, 4, 8, 16, 32); PEN: penalty functor. Both compile-time so that the per-class loops unroll.
template <int KT, int PEN>
__global__ void __launch_bounds__(kCBlock, 1)
saga_dense_cluster_generic_kernel(FitDev* __restrict__ fit, Progress* __restrict__ prog, const RoundArgs ra) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr bool kScalar = (KT == 1);
  Progress& pg = *prog;
  const FitDev& f = *fit;
  const uint32_t cta = cluster_ctarank();
  if (ra.n_epochs <= 0 || pg.status != kRunning) {       // uniform over the cluster
    if (cta == 0 && threadIdx.x == 0) {
      pg.epochs_last_launch = 0;
      publish_progress(f.mirror, pg, ra.round_id);
    }
    return;
  }
  const bool free_run = (ra.flags & 1) != 0;
  const uint64_t t_start = globaltimer_ns();

  // warp 0: control warp (exchange, gradient step, row copies); warps 1..8: the CTA's 256 feature lanes. The control
  // warp owns no features, so the feature warps can prepare the update's step constants (three FP64 divisions) while
  // the control warp is in the exchange, and nothing but the K-value chain sits between the two block barriers.
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool control = warp == 0;
  const int tid = static_cast<int>(threadIdx.x) - 32;      // feature lane of this CTA (negative in the control warp)
  const int fwarp = warp - 1;
  const int K = kScalar ? 1 : f.K, p = f.p, ld = f.ld, Ky = f.Ky;
  const int64_t n = f.n;
  const double nd = static_cast<double>(static_cast<uint32_t>(n));
  const double rn = 1.0 / nd;
  const int family = f.family;
  const bool fit_intercept = f.fit_intercept != 0;
  const int nch = (p + kLanes - 1) / kLanes;             // 256-feature slices per CTA and row
  const int nf = nch * kCT;                               // feature slots of this CTA (some beyond p)

  uint32_t dyn_bytes;
  asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_bytes));
  const size_t fixed = (cluster_fixed_bytes(nch) + 15) & ~size_t(15);
  const bool state_in_smem = fixed + sizeof(double) * 2 * size_t(K) * nf <= size_t(dyn_bytes);
  ClusterSmem sm = carve_cluster(smem_raw, K, nch, state_in_smem);

  const int li = pg.lambda_ind;
  const double gamma = f.gamma[li], alpha = f.alpha[li], beta = f.beta[li];
  const double r = 1.0 - alpha * gamma;     // wscale_update

  // feature slot q = i * 256 + tid  <->  feature j = 2048 * i + 256 * cta + tid
  auto feature_of = [&](int i) { return kLanes * i + kCT * static_cast<int>(cta) + tid; };
  // state addressing: shared memory [k][slot] or global class-major [k][j]
  double* Wg = f.W;
  double* Gg = f.gsum;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kCRing; ++i) mbar_init(&sm.full[i], 1);
    mbar_init(&sm.pbar[0], 1);
    mbar_init(&sm.pbar[1], 1);
    mbar_init(&sm.cbar[0], kCluster);
    mbar_init(&sm.cbar[1], kCluster);
    fence_barrier_init();
  }
  if (state_in_smem && !control) {
    for (int i = 0; i < nch; ++i) {
      const int j = feature_of(i);
      for (int k = 0; k < K; ++k) {
        sm.W[size_t(k) * nf + i * kCT + tid] = j < p ? Wg[size_t(k) * p + j] : 0.0;
        sm.G[size_t(k) * nf + i * kCT + tid] = j < p ? Gg[size_t(k) * p + j] : 0.0;
      }
    }
  }
  __syncthreads();
  cluster_sync_all();       // every CTA's barriers exist before anybody arrives on them remotely

  // this CTA's slices of a row: slice i covers features [2048 i + 256 cta, +256) clipped to the row's padded length
  const uint32_t* __restrict__ seq = ra.seq;
  const int64_t total = n * ra.n_epochs;
  auto slice_bytes = [&](int i) {
    const int j0 = kLanes * i + kCT * static_cast<int>(cta);
    const int len = (j0 >= ld) ? 0 : ((ld - j0 < kCT) ? ld - j0 : kCT);
    return static_cast<uint32_t>(len) * 8u;
  };
  uint32_t row_bytes = 0;
  for (int i = 0; i < nch; ++i) row_bytes += slice_bytes(i);
  auto issue_row = [&](int64_t q, uint32_t sq) {      // one thread; sq = seq[q]
    const int slot = static_cast<int>(q % kCRing);
    const double* src = f.xd + size_t(sq) * ld;
    if (row_bytes == 0) {
      mbar_arrive(&sm.full[slot]);
      return;
    }
    mbar_expect_tx(&sm.full[slot], row_bytes);
    for (int i = 0; i < nch; ++i) {
      const uint32_t b = slice_bytes(i);
      if (b) bulk_g2s(sm.ring + (size_t(slot) * nch + i) * kCT, src + kLanes * i + kCT * cta, b, &sm.full[slot]);
    }
  };
  // row copies are issued by lane 31 of the control warp (not a class lane), one row per update, with the sample index
  // fetched one update ahead
  const bool issuer = control && lane == 31;
  int64_t issued = 0;
  uint32_t s_refill = 0;
  if (issuer) {
    for (; issued < kCRing - 1 && issued < total; ++issued) issue_row(issued, seq[issued]);
    if (issued < total) s_refill = seq[issued];
  }

  // intercept state: class k in lane k of the control warp of EVERY CTA (identical, redundant)
  const bool owner = control && lane < K;
  double b_reg = 0.0, gsi_reg = 0.0;
  if (owner) {
    b_reg = f.b[lane];
    gsi_reg = f.gsi[lane];
  }

  double wscale = 1.0;
  uint32_t it_outer = pg.it_outer;
  uint32_t epochs_done = 0;
  int64_t tg = 0;
  uint32_t prev_s = 0xffffffffu, prev2_s = 0xffffffffu;
  double prev_g = 0.0, prev2_g = 0.0;
  bool finished = false;

  // per-sample operands one update ahead. The gradient memory is read and written by CTA 0 ALONE (same lanes, so program
  // order is all the coherence it needs; the last two samples are forwarded from registers because their stores may
  // still be in flight when the next value is prefetched) and travels to the other CTAs with CTA 0's partial sums.
  auto fetch_y = [&](uint32_t sx) { return f.yt[size_t(sx) * Ky + (Ky == 1 ? 0 : lane)]; };
  auto fetch_gm = [&](uint32_t sx) { return cta == 0 ? f.gmem[size_t(sx) * K + lane] : 0.0; };
  uint32_t s_cur = seq[0], s_nxt = (total > 1) ? seq[1] : 0u;
  double y_cur = 0.0, gm_cur = 0.0;
  if (owner) {
    y_cur = fetch_y(s_cur);
    gm_cur = fetch_gm(s_cur);
  }

  const uint32_t a_part = smem_u32(sm.part), a_pbar = smem_u32(sm.pbar);

  for (int ep = 0; ep < ra.n_epochs && !finished; ++ep) {
    for (int64_t t = 0; t < n; ++t, ++tg) {
      const uint32_t s = s_cur;
      const int slot = static_cast<int>(tg % kCRing);
      const uint32_t parity = static_cast<uint32_t>((tg / kCRing) & 1);
      const uint32_t xpar = static_cast<uint32_t>(tg & 1);                // exchange buffer of this update
      const uint32_t xphase = static_cast<uint32_t>((tg >> 1) & 1);       // phase of pbar[xpar]
      const double y_val = y_cur;
      const double gm_mine = (s == prev_s) ? prev_g : ((s == prev2_s) ? prev2_g : gm_cur);     // meaningful in CTA 0
      s_cur = s_nxt;
      if (tg + 2 < total) s_nxt = seq[tg + 2];
      if (owner && tg + 1 < total) {
        y_cur = fetch_y(s_cur);
        gm_cur = fetch_gm(s_cur);
      }

      const double* __restrict__ xr = sm.ring + size_t(slot) * nch * kCT;

      // ---- A: partial dot products of this lane (ascending j), warp butterfly, warp sums
      if (!control) {
        mbar_wait(&sm.full[slot], parity);
        double acc[KT];
#pragma unroll
        for (int k = 0; k < KT; ++k) acc[k] = 0.0;
        for (int i = 0; i < nch; ++i) {
          const int j = feature_of(i);
          if (j < p) {
            const double xj = xr[i * kCT + tid];
#pragma unroll
            for (int k = 0; k < KT; ++k)
              if (kScalar || k < K)
                acc[k] += (state_in_smem ? sm.W[size_t(k) * nf + i * kCT + tid] : Wg[size_t(k) * p + j]) * xj;
          }
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
            if (kScalar || k < K) sm.red[fwarp * 32 + k] = acc[k];
        }
      }
      __syncthreads();   // (1) warp sums visible; every thread is past step C of the previous update

      // this update's step constants (functions of the deterministic wscale track), in the feature warps while the
      // control warp exchanges: gamma / wscale, (beta gamma) / wscale with wscale as it will be after this step
      double gw = 0.0, step = 0.0, thr = 0.0, ws_c = 0.0;
      const double bgs = beta * gamma * 1.0;
      if (!control) {
        ws_c = ((wscale < kSmall) ? 1.0 : wscale) * r;
        gw = gamma / ws_c;
        step = gamma / ws_c * 1.0;
        thr = bgs / ws_c;
      }

      if (issuer && issued < total) {      // refill the slot the previous update released
        issue_row(issued, s_refill);
        ++issued;
        if (issued < total) s_refill = seq[issued];
      }

      // ---- B: CTA sum (8 warps ascending) -> every CTA; then the 8 CTA sums ascending; gradient (redundant per CTA)
      if (control) {
        if (lane < K) {
          double tsum = 0.0;
          for (int w = 0; w < 8; ++w) tsum += sm.red[w * 32 + lane];
          const uint32_t my_slot = a_part + ((xpar * (kCluster + 1) + cta) * 32u + static_cast<uint32_t>(lane)) * 8u;
          const uint32_t gm_slot = a_part + ((xpar * (kCluster + 1) + kCluster) * 32u + static_cast<uint32_t>(lane)) * 8u;
#pragma unroll
          for (uint32_t c = 0; c < kCluster; ++c) {
            const uint32_t bar_c = map_to_cta(a_pbar + xpar * 8u, c);
            st_async_f64(map_to_cta(my_slot, c), tsum, bar_c);
            if (cta == 0) st_async_f64(map_to_cta(gm_slot, c), gm_mine, bar_c);
          }
        }
        // this CTA expects (kCluster + 1) * K doubles per update on its own barrier: one local arrival arms the phase
        if (lane == 0) mbar_expect_tx(&sm.pbar[xpar], static_cast<uint32_t>((kCluster + 1) * K * 8));
        mbar_wait(&sm.pbar[xpar], xphase);
        const double gm_val = (lane < K) ? sm.part[(xpar * (kCluster + 1) + kCluster) * 32 + lane] : 0.0;
        const bool valid = lane < K;
        double lp = 0.0;
        if (valid) {
          double dot = 0.0;
          for (int c = 0; c < kCluster; ++c) dot += sm.part[(xpar * (kCluster + 1) + c) * 32 + lane];
          lp = dot * wscale + b_reg;
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
          const double gch = g - gm_val;
          if (cta == 0) f.gmem[size_t(s) * K + lane] = g;
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
        for (int i = 0; i < nch && !control; ++i) {
          const int j = feature_of(i);
          if (j < p)
            for (int k = 0; k < K; ++k) {
              if (state_in_smem) sm.W[size_t(k) * nf + i * kCT + tid] *= wscale;
              else Wg[size_t(k) * p + j] *= wscale;
            }
        }
        wscale = 1.0;
      }
      wscale *= r;
      __syncthreads();   // (2) g_change visible

      // ---- C: fused coefficient step, prox, gradient-average update on the owned features
      // (src/saga-dense.h:176-183; penalty functors src/penalties.h:27-79 with scaling = 1)
      double gch[KT];
#pragma unroll
      for (int k = 0; k < KT; ++k) gch[k] = (kScalar || k < K) ? sm.gch[k] : 0.0;
      for (int i = 0; i < nch && !control; ++i) {
        const int j = feature_of(i);
        if (j >= p) continue;
        const double xj = xr[i * kCT + tid];
        double w[KT], gs[KT];
        double sq = 0.0;
#pragma unroll
        for (int k = 0; k < KT; ++k) {
          if (kScalar || k < K) {
            double* Wp = state_in_smem ? &sm.W[size_t(k) * nf + i * kCT + tid] : &Wg[size_t(k) * p + j];
            double* Gp = state_in_smem ? &sm.G[size_t(k) * nf + i * kCT + tid] : &Gg[size_t(k) * p + j];
            gs[k] = *Gp;
            const double gx = gch[k] * xj;
            const double v = (*Wp - gx * gw) - step * gs[k];
            w[k] = (PEN == kElasticNet) ? soft_threshold(v, thr) : v;
            if (PEN == kGroupLasso) sq += v * v;
            *Gp = gs[k] + div_by_n(gx, nd, rn);
          }
        }
        if (PEN == kGroupLasso) {
          const double factor = bgs / sqrt(sq);
          const double mult = 1.0 - factor / ws_c;
#pragma unroll
          for (int k = 0; k < KT; ++k)
            if (kScalar || k < K) w[k] = (factor < 1.0) ? w[k] * mult : 0.0;
        }
#pragma unroll
        for (int k = 0; k < KT; ++k)
          if (kScalar || k < K) {
            if (state_in_smem) sm.W[size_t(k) * nf + i * kCT + tid] = w[k];
            else Wg[size_t(k) * p + j] = w[k];
          }
      }
    }

    // ---- epoch end: unscale, convergence over the whole cluster (src/saga-dense.h:188-208, src/utils.h:240-262)
    double mc = 0.0, ms = 0.0;
    for (int i = 0; i < nch && !control; ++i) {
      const int j = feature_of(i);
      if (j >= p) continue;
      for (int k = 0; k < K; ++k) {
        const size_t e = size_t(k) * p + j;
        double* Wp = state_in_smem ? &sm.W[size_t(k) * nf + i * kCT + tid] : &Wg[e];
        const double w = *Wp * wscale;
        *Wp = w;
        mc = fmax(mc, fabs(w - f.Wprev[e]));
        ms = fmax(ms, fabs(w));
        f.Wprev[e] = w;
      }
    }
    wscale = 1.0;
    mc = warp_max(mc);
    ms = warp_max(ms);
    __syncthreads();       // red[] is free again
    if (lane == 0 && !control) {
      sm.red[fwarp * 32] = mc;
      sm.red[fwarp * 32 + 1] = ms;
    }
    __syncthreads();
    const uint32_t epar = static_cast<uint32_t>(ep & 1), ephase = static_cast<uint32_t>((ep >> 1) & 1);
    if (threadIdx.x == 0) {
      double mc_c = 0.0, ms_c = 0.0;
      for (int w = 0; w < 8; ++w) {
        mc_c = fmax(mc_c, sm.red[w * 32]);
        ms_c = fmax(ms_c, sm.red[w * 32 + 1]);
      }
      const uint32_t a_conv = smem_u32(sm.conv) + ((epar * kCluster + cta) * 2u) * 8u;
      for (uint32_t c = 0; c < kCluster; ++c) {
        st_cluster_f64(map_to_cta(a_conv, c), mc_c);
        st_cluster_f64(map_to_cta(a_conv + 8u, c), ms_c);
        arrive_cluster(map_to_cta(smem_u32(sm.cbar) + epar * 8u, c));
      }
    }
    wait_cluster(&sm.cbar[epar], ephase);
    double mc_all = 0.0, ms_all = 0.0;
    for (int c = 0; c < kCluster; ++c) {
      mc_all = fmax(mc_all, sm.conv[(epar * kCluster + c) * 2]);
      ms_all = fmax(ms_all, sm.conv[(epar * kCluster + c) * 2 + 1]);
    }
    __syncthreads();
    const bool all_zero = (ms_all == 0.0) && (mc_all == 0.0);
    const bool no_change = (ms_all != 0.0) && (mc_all / ms_all <= f.tol);
    ++it_outer;
    ++epochs_done;
    finished = !free_run && ((all_zero || no_change) || !(it_outer < f.max_iter));
  }

  // drain copies that were issued but never consumed (early stop) before the shared memory goes away
  if (issuer)
    for (int64_t q = tg; q < issued; ++q)
      mbar_wait(&sm.full[static_cast<int>(q % kCRing)], static_cast<uint32_t>((q / kCRing) & 1));
  __syncthreads();

  if (state_in_smem && !control) {
    for (int i = 0; i < nch; ++i) {
      const int j = feature_of(i);
      if (j >= p) continue;
      for (int k = 0; k < K; ++k) {
        Wg[size_t(k) * p + j] = sm.W[size_t(k) * nf + i * kCT + tid];
        Gg[size_t(k) * p + j] = sm.G[size_t(k) * nf + i * kCT + tid];
      }
    }
  }
  if (cta == 0 && owner) {
    f.b[lane] = b_reg;
    f.gsi[lane] = gsi_reg;
  }
  cluster_sync_all();      // nobody leaves while another CTA may still write into its shared memory
  if (cta == 0 && threadIdx.x == 0) {
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

template <int KT, int PEN>
static cudaError_t launch_cluster_variant(size_t smem, FitDev* fit, Progress* prog, const RoundArgs& ra, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(saga_dense_cluster_generic_kernel<KT, PEN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(kCluster, 1, 1);
  cfg.blockDim = dim3(kCBlock, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, saga_dense_cluster_generic_kernel<KT, PEN>, fit, prog, ra);
}

template <int KT>
static cudaError_t launch_cluster_kt(int pen, size_t smem, FitDev* fit, Progress* prog, const RoundArgs& ra, cudaStream_t st) {
  switch (pen) {
    case kRidge: return launch_cluster_variant<KT, kRidge>(smem, fit, prog, ra, st);
    case kElasticNet: return launch_cluster_variant<KT, kElasticNet>(smem, fit, prog, ra, st);
    default: return launch_cluster_variant<KT, kGroupLasso>(smem, fit, prog, ra, st);
  }
}

cudaError_t launch_saga_dense_cluster_generic(int K, int pen, size_t smem, FitDev* fit, Progress* prog, const RoundArgs& ra, cudaStream_t st) {
  switch (dense_kt_bucket(K)) {
    case 1: return launch_cluster_kt<1>(pen, smem, fit, prog, ra, st);
    case 4: return launch_cluster_kt<4>(pen, smem, fit, prog, ra, st);
    case 8: return launch_cluster_kt<8>(pen, smem, fit, prog, ra, st);
    case 16: return launch_cluster_kt<16>(pen, smem, fit, prog, ra, st);
    default: return launch_cluster_kt<32>(pen, smem, fit, prog, ra, st);
  }
}

}  // namespace sgd
