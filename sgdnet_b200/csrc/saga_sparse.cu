// saga_sparse.cu — sparse SAGA epochs with just-in-time lagged prox (reference: src/saga-sparse.h:76-155, 256-371).
//
// One persistent CTA per fit. Kernels:
//
//  saga_sparse_wave_kernel<S>  K == 1 (gaussian / binomial), no virtual centring: the headline path (BASELINE configs
//     2 and 5). S worker warps take the sampled rows round-robin and overlap every update that shares no feature with
//     a row still in flight; one chain warp runs the strictly serial scalar recurrence (intercept -> lp -> gradient);
//     one producer warp streams each row's padded-CSR index / value runs (and, for rows that have conflicts, their
//     conflict codes) into a shared-memory ring with 1-D bulk copies (cp.async.bulk, complete_tx on an mbarrier per
//     slot). Coefficient state is one 32-byte record {w, g_sum, lag} per feature, L2-resident, read and written
//     once per nonzero per update with 256-bit accesses. The exactness argument is in the comment block below.
//     Algorithmic HBM bytes per update: 12*nnz_row + 16 (row info) + 4 (index) + 8 (y) + 16 (gradient memory).
//
//  wave_deps_kernel  which features of a row are also held by one of the S-1 rows before it, from the host-drawn
//     sampling sequence alone (so it can run ahead of the solver).
//
//  saga_sparse_generic_kernel  any K <= 32 and/or standardize = TRUE (the reference's O(p*K) virtual-centring sweeps,
//     src/saga-sparse.h:127-128, 276-277, reproduced as block-wide passes). Phases are separated by block barriers.
//
// Epoch end (all): Reset(n) over all features by the whole CTA, W *= wscale, lag = 0, convergence test
// (src/saga-sparse.h:340-348, 367; src/utils.h:240-262).
#include <algorithm>
#include <cstddef>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace sgd {

constexpr int kCap = 128;       // entries per ring slot; longer rows are read in place
constexpr int kChunks = kCap / 32;

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

// ============================================================================================ K == 1 wavefront path
//
// Exactness argument. The reference's update t reads and writes only (i) the coefficient state (w, g_sum, lag) of the
// features of row s_t, (ii) gmem[s_t], (iii) the scalars (intercept, g_sum_intercept, wscale). wscale is a
// deterministic function of t (wscale *= r every step, reset to 1 whenever it drops below SMALL), so every warp can
// track it on its own with the same multiplications. Two updates whose rows share no feature and no sample therefore
// interact only through the scalar chain  intercept -> lp -> gradient -> intercept.  The kernel splits the update
// along exactly that line:
//
//   chain warp (1)    owns the intercept state in registers and walks the rows in order: lp = dot + b, gradient,
//                     g_change (published to the row's worker), gradient memory, intercept step. This is the only
//                     strictly serial part of SAGA and nothing else is on its instruction stream.
//   worker warps (S)  take the rows round-robin (row t -> worker t mod S): gather the row's coefficient state, catch
//                     it up (LaggedUpdate k = t), reduce the dot product, hand it to the chain warp, and once g_change
//                     is back do AddWeighted / LaggedUpdate(k = t+1) / AddWeighted(g_sum) and scatter. Rows in flight
//                     overlap unless they share a feature; which of a row's features ARE touched by one of the
//                     previous S-1 rows is computed exactly, ahead of the launch, by wave_deps_kernel from the known
//                     sampling sequence, and only those gathers wait (fdone of the conflicting row).
//   producer warp (1) streams rows (index run, value run, conflict codes) into a shared-memory ring with bulk copies.
//
// Each element of the state still sees exactly the reference's sequence of floating point operations, in the
// reference's order; only independent work is overlapped.
//
// Coefficient state lives in HBM/L2 as ONE 32-byte record per feature {w, g_sum, lag}: a row's gather and scatter
// are then a single 256-bit access per nonzero (LDG.E.ENL2.256 / STG.E.ENL2.256). Measured on B200 from one SM
// (scripts/microbench2.cu): three arrays in L2 630 cycles/row of 100 features, packed record 250, packed record in
// the distributed shared memory of a 16-CTA cluster 564 - which is why the state is not held in DSMEM.
//
//  - rdy[q]    completes when row q's dot product and per-sample operands are in the chain queue.
//  - gok[q]    completes when the chain warp has published g_change of row q (and stored the gradient memory).
//  - fdone[q]  completes when row q's scatter is visible in HBM (only rows behind a serial row wait for it; a
//              conflict with an ordinary row in flight is resolved through the forwarding buffer after gok).
//  - done[q]   like fdone but chained in row order: "done(q)" means every row <= q is complete, which bounds the rows
//              in flight to S consecutive ones (the window wave_deps_kernel looked at).
//  - a row whose nonzeros do not fit a ring slot, or at which the wscale reset (src/saga-sparse.h:285-295) fires,
//    is run serially: it waits for every earlier row, and the rows after it wait for it.
constexpr int kWSlots = 32;     // row ring depth: S rows held by the workers + rows in flight from HBM
constexpr int kSeq = 16;        // per-row barrier / queue / forwarding rings (indexed by row sequence number; at most S <= 12 rows are in flight)
static_assert(kChunks == 4, "conflict codes pack four 16-bit entries per lane");

struct WaveSlotMeta {
  uint32_t s;
  int32_t nnz;
  int64_t start;
  double y;
  uint32_t dup;     // distance to the most recent in-window row with the same sample (0 = none)
  uint32_t has_code;   // the row shares a feature with a row of its window: its conflict codes were copied
};

struct __align__(128) WaveSmem {
  double val[kWSlots][kCap];
  int32_t idx[kWSlots][kCap];
  uint64_t code[kWSlots][32];   // per lane: four 16-bit conflict entries (wave_deps_kernel)
  double fw_w[kSeq][kCap];      // forwarding: a row's caught-up (w, g_sum) and its x values by nonzero position, for
  double fw_g[kSeq][kCap];      //   the rows that touch the same feature while it is still in flight: they redo
  double fw_x[kSeq][kCap];      //   the row's coefficient step themselves as soon as its g_change is known
  double c_sc[kSeq], c_step1[kSeq], c_thr1[kSeq];   // the row's step constants (functions of its wscale)
  WaveSlotMeta meta[kWSlots];
  uint64_t full[kWSlots];
  uint64_t empty[kWSlots];
  uint64_t rdy[kSeq];
  uint64_t gok[kSeq];
  uint64_t fdone[kSeq];
  uint64_t done[kSeq];
  double q_dot[kSeq];     // worker -> chain: (W.x_s) * wscale
  double q_ya[kSeq];      //                  y (gaussian) or 1 - y (binomial)
  double q_gm[kSeq];      //                  gradient memory of the sample
  double q_gch[kSeq];     // chain -> worker: g_change
  uint32_t q_s[kSeq];     // worker -> chain: sample id
  double red[2 * 32];
  double wscale_s;
};

__device__ __forceinline__ void ld_state(const FeatState* p, double& w, double& g, uint32_t& lag) {
  unsigned long long a, b, c;
  [[maybe_unused]] unsigned long long pad;      // the record's padding word
  asm volatile("ld.global.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(pad) : "l"(p) : "memory");
  w = __longlong_as_double(static_cast<long long>(a));
  g = __longlong_as_double(static_cast<long long>(b));
  lag = static_cast<uint32_t>(c);
}
__device__ __forceinline__ void st_state(FeatState* p, double w, double g, uint32_t lag) {
  asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(__double_as_longlong(w)), "l"(__double_as_longlong(g)),
               "l"(static_cast<unsigned long long>(lag)), "l"(0ull)
               : "memory");
}

#ifndef SGD_SLEEP_NS
#define SGD_SLEEP_NS 100
#endif
__device__ __forceinline__ void wait_row(uint64_t* ring, uint32_t q) {
  mbar_wait(&ring[q % kSeq], (q / kSeq) & 1u);
}
// for waits that are long by design (a worker ahead of the chain): leave the issue slots to the other warps
__device__ __forceinline__ void wait_row_relaxed(uint64_t* ring, uint32_t q) {
  uint64_t* bar = &ring[q % kSeq];
  const uint32_t par = (q / kSeq) & 1u;
  while (!mbar_test_wait(bar, par)) __nanosleep(SGD_SLEEP_NS);
}

// ---- conflict codes: for every row instance q = epoch*n + t of the staged sequence and every nonzero position e of
// its row, a 16-bit entry about the most recent earlier row OF THE SAME EPOCH within `window` rows that holds the same
// feature: bits 0-3 its distance d (0 = no such row), bits 4-10 the feature's position in that row (where its
// forwarded state sits in shared memory), bit 11 "read it from HBM instead" (that row was too long for a ring slot and
// ran serially). Lane l of a worker reads one 64-bit word holding the entries of positions l, l+32, l+64, l+96.
constexpr uint32_t kCodeGlobal = 1u << 11;

constexpr int kDepSlots = 32;       // row instances staged per block iteration: the `window` rows before the block's rows + its own
static_assert(kDepSlots <= 32, "one bit per staged row; distances are 4-bit (window <= 15)");
constexpr int kDepBuckets = 8192;   // hashed feature buckets; a bucket holds one bit per staged row
__device__ __forceinline__ uint32_t dep_hash1(int32_t j) { return (static_cast<uint32_t>(j) * 2654435761u) >> 19; }   // 13 bits
__device__ __forceinline__ uint32_t dep_hash2(int32_t j) { return (static_cast<uint32_t>(j) * 0x85ebca6bu + 0x27d4eb2fu) >> 19; }

// Which features of a row are also held by one of the `window` rows before it. Per block iteration 32 consecutive row
// instances (the last 32 - window of them are the block's own) are staged in shared memory: their index runs, and ONE
// table of 8192 hashed feature buckets x 32 bits - bit r of bucket h says "staged row r holds a feature that hashes to
// h" (two hash functions into the same table). A feature of row r is then tested against ALL its predecessors with two
// loads: table[h1] & table[h2], masked to the rows of its window, is the set of rows that MAY hold it (each wrongly
// with probability (200 / 8192)^2 = 0.06 %); the candidates, nearest first, are confirmed by binary search in their
// sorted index runs, which also yields the feature's position there. This is exact, and a quarter of the shared-memory
// traffic of testing seven per-row filters one after the other (the first form of this kernel): the pass costs the
// many-fits regime as much SM time as it saves the solver, so its cost per row is what bounds a batch.
constexpr int kDepWarps = 8;
constexpr int kDepPerWarp = kDepSlots / kDepWarps;      // staged rows per warp and block iteration
__global__ void __launch_bounds__(kDepWarps * 32, 4)
wave_deps_kernel(const FitDev* __restrict__ fit, const RoundArgs ra, int window) {
  extern __shared__ __align__(16) unsigned char deps_smem[];
  int32_t (*sidx)[kCap] = reinterpret_cast<int32_t (*)[kCap]>(deps_smem);
  uint32_t* table = reinterpret_cast<uint32_t*>(deps_smem + sizeof(int32_t) * kDepSlots * kCap);
  int32_t* snnz = reinterpret_cast<int32_t*>(table + kDepBuckets);
  uint32_t* ssamp = reinterpret_cast<uint32_t*>(snnz + kDepSlots);
  uint32_t& slong = *(ssamp + kDepSlots);   // staged rows too long for a ring slot (they run serially; their features come from HBM)
  // depends on the sampling sequence alone (not on the fit's progress), so it can run ahead of the solver
  if (ra.n_epochs <= 0 || ra.dep == nullptr) return;
  const FitDev& f = *fit;
  const int64_t n = f.n;
  const int64_t total = n * ra.n_epochs;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int32_t* __restrict__ ci = f.ci;
  const int own = kDepSlots - window;                 // rows of its own per block iteration
  const int64_t n_chunks = (total + own - 1) / own;

  // Staging is a chain of three dependent global loads (sequence -> row descriptor -> index run). It is software-
  // pipelined over the block iterations so that no warp ever waits for it: while iteration i is searched, the index
  // runs of iteration i + 1 and the descriptors of iteration i + 2 are in flight. Warp w stages slots w, w + 8, ...
  uint32_t s_a[kDepPerWarp], s_b[kDepPerWarp];          // samples of the next / the one after next iteration's slots
  RowInfo ri_a[kDepPerWarp], ri_b[kDepPerWarp];
  int32_t jj[kDepPerWarp][kChunks];                     // index values of the next iteration's slots (lane e, e + 32, ...)
  auto fetch_desc = [&](int64_t chunk, uint32_t (&sv)[kDepPerWarp], RowInfo (&rv)[kDepPerWarp]) {
#pragma unroll
    for (int u = 0; u < kDepPerWarp; ++u) {
      const int64_t q = chunk * own - window + (warp + kDepWarps * u);
      sv[u] = 0xffffffffu;
      rv[u] = RowInfo{};
      if (chunk < n_chunks && q >= 0 && q < total) {
        sv[u] = ra.seq[q];
        rv[u] = f.rows[sv[u]];
      }
    }
  };
  auto fetch_idx = [&](const RowInfo (&rv)[kDepPerWarp]) {
#pragma unroll
    for (int u = 0; u < kDepPerWarp; ++u)
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        const int e = c * 32 + lane;
        jj[u][c] = (rv[u].nnz <= kCap && e < rv[u].nnz) ? ci[rv[u].start + e] : -1;
      }
  };
  fetch_desc(blockIdx.x, s_a, ri_a);
  fetch_idx(ri_a);
  fetch_desc(int64_t(blockIdx.x) + gridDim.x, s_b, ri_b);

  for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    const int64_t q0 = chunk * own;
    __syncthreads();                                   // the previous iteration's searches are done
    for (int i = threadIdx.x; i < kDepBuckets; i += blockDim.x) table[i] = 0u;
    if (threadIdx.x == 0) slong = 0u;
    __syncthreads();
    // ---- this iteration's rows from the registers into shared memory
#pragma unroll
    for (int u = 0; u < kDepPerWarp; ++u) {
      const int r = warp + kDepWarps * u;              // slot r holds row instance q0 - window + r
      const int32_t nnz = ri_a[u].nnz;
      if (nnz <= kCap) {
        const uint32_t bit = 1u << r;
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          const int32_t j = jj[u][c];
          if (j >= 0) {
            sidx[r][c * 32 + lane] = j;
            atomicOr(&table[dep_hash1(j)], bit);
            atomicOr(&table[dep_hash2(j)], bit);
          }
        }
      } else if (lane == 0) {
        atomicOr(&slong, 1u << r);
      }
      if (lane == 0) {
        snnz[r] = nnz;
        ssamp[r] = s_a[u];
      }
    }
    // ---- next iteration's index runs (their descriptors arrived an iteration ago) and the descriptors after them
#pragma unroll
    for (int u = 0; u < kDepPerWarp; ++u) {
      s_a[u] = s_b[u];
      ri_a[u] = ri_b[u];
    }
    fetch_idx(ri_a);
    fetch_desc(chunk + 2 * int64_t(gridDim.x), s_b, ri_b);
    __syncthreads();
    const uint32_t long_rows = slong;
    for (int rl = warp; rl < own; rl += kDepWarps) {
      const int64_t q = q0 + rl;
      if (q >= total) break;
      const int64_t t = q % n;
      const int me = window + rl;
      const int32_t nnz = snnz[me];
      const int dmax = static_cast<int>(t < window ? t : window);      // only rows of the same epoch
      // staged rows me - dmax .. me - 1
      const uint32_t wmask = ((1u << me) - 1u) & ~((1u << (me - dmax)) - 1u);
      // nearest predecessor with the same sample: lane d - 1 looks at distance d
      const uint32_t same = __ballot_sync(0xffffffffu, lane < dmax && ssamp[me - 1 - (lane < dmax ? lane : 0)] == ssamp[me]);
      const uint32_t dupd = same ? static_cast<uint32_t>(__ffs(same)) : 0u;
      uint32_t ent[kChunks];
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        const int e = c * 32 + lane;
        ent[c] = 0;
        if (nnz > kCap || e >= nnz) continue;
        const int32_t j = sidx[me][e];
        uint32_t cand = ((table[dep_hash1(j)] & table[dep_hash2(j)]) | long_rows) & wmask;
        while (cand != 0u) {                            // nearest predecessor first
          const int pr = 31 - __clz(cand);
          cand &= ~(1u << pr);
          const uint32_t d = static_cast<uint32_t>(me - pr);
          if ((long_rows >> pr) & 1u) {
            ent[c] = d | kCodeGlobal;
            break;
          }
          const int32_t nnz2 = snnz[pr];
          const int32_t* __restrict__ c2 = sidx[pr];
          int lo = 0, hi = nnz2;                        // first position with c2[pos] >= j
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (c2[mid] < j) lo = mid + 1; else hi = mid;
          }
          if (lo < nnz2 && c2[lo] == j) {
            ent[c] = d | (static_cast<uint32_t>(lo) << 4);
            break;
          }
        }
      }
      // most rows share nothing with their window (82 % at config 2's density): their 256 B of codes are neither
      // written here nor copied by the solver's producer; bit 7 of dup[] says which rows have codes
      const bool any = __any_sync(0xffffffffu, (ent[0] | ent[1] | ent[2] | ent[3]) != 0u);
      if (any)
        ra.dep[q * 32 + lane] = uint64_t(ent[0]) | (uint64_t(ent[1]) << 16) | (uint64_t(ent[2]) << 32) | (uint64_t(ent[3]) << 48);
      if (lane == 0) ra.dup[q] = static_cast<uint8_t>(dupd | (any ? 0x80u : 0u));
    }
  }
}

#ifdef SGD_WAVE_TRACE
// Timeline trace (measurement build only): clock64 of eight events per row for rows [kTraceFrom, kTraceFrom + kTraceRows)
// of the first epoch of a launch, written straight to global memory by one lane (fire-and-forget stores).
//  0 chain: row taken up   1 chain: gok published   2 chain: next row's operands in hand
//  3 worker: row in the ring (full)   4 worker: state gathered and caught up   5 worker: rdy published
//  6 worker: gok seen   7 worker: row complete (done)          plus [8] nearest conflict distance, [9] next_ready
constexpr uint32_t kTraceFrom = 50000, kTraceRows = 4096;
__device__ long long g_wave_trace[kTraceRows][10];
#define TRACE(ev, t_) do { if ((t_) - kTraceFrom < kTraceRows && ep_trace) g_wave_trace[(t_) - kTraceFrom][ev] = clock64(); } while (0)
#define TRACEV(ev, t_, v_) do { if ((t_) - kTraceFrom < kTraceRows && ep_trace) g_wave_trace[(t_) - kTraceFrom][ev] = (v_); } while (0)
extern "C" void sgdnet_debug_wave_trace(long long* out) { cudaMemcpyFromSymbol(out, g_wave_trace, sizeof(g_wave_trace)); }
#else
#define TRACE(ev, t_)
#define TRACEV(ev, t_, v_)
#endif


// A taken branch costs a lone warp about 23 cycles on sm_100a (scripts/microbench.cu: a dependent DFMA is 8), and the
// solver is a handful of lone warps, so the hot paths below are written to compile to straight-line predicated code:
// loads use clamped indices instead of guards, stores carry their predicate into the instruction, "skip when the lag
// is zero" is a select. None of this changes a single floating point operation that reaches the state.
template <int PEN>
__device__ __forceinline__ double penalty_k1(double w, double gs, double step, double thr) {
  const double v = w - step * gs;
  return (PEN == kElasticNet) ? soft_threshold(v, thr) : v;   // K == 1 never uses the group penalty
}

__device__ __forceinline__ void st_state_if(bool pred, FeatState* p, double w, double g, uint32_t lag) {
  asm volatile(
      "{\n"
      ".reg .pred pp;\n"
      "setp.ne.b32 pp, %5, 0;\n"
      "@pp st.global.v4.b64 [%0], {%1,%2,%3,%4};\n"
      "}\n" ::"l"(p),
      "l"(__double_as_longlong(w)), "l"(__double_as_longlong(g)), "l"(static_cast<unsigned long long>(lag)), "l"(0ull),
      "r"(static_cast<int>(pred))
      : "memory");
}

struct WaveConst {
  uint32_t n;
  int p, family;
  bool fit_intercept;
  double nd, gamma, r, sc2, bg;
  const double* ls_table;
  FeatState* st;
};

// ------------------------------------------------------------------ producer: lane l feeds ring slot l
// Warp-synchronous: the lanes load the sample index, row descriptor, response and duplicate code of their next row
// together (32 independent requests in flight), one batch ahead of the batch being issued, then the warp polls the
// slots' "empty" barriers and issues each row's bulk copies as soon as its slot is released.
__device__ __noinline__ void wave_producer(WaveSmem& sm, const FitDev& f, const RoundArgs& ra, int ep, uint32_t q_base,
                                           uint32_t n, int lane) {
  const uint32_t* __restrict__ eseq = ra.seq + size_t(ep) * n;
  const uint8_t* __restrict__ edup = ra.dup + size_t(ep) * n;
  const uint64_t* __restrict__ edep = ra.dep + size_t(ep) * n * 32;
  uint32_t t = (static_cast<uint32_t>(lane) - q_base) % kWSlots;   // q = q_base + t lands on slot `lane`
  bool have = t < n;
  uint32_t s = 0, dv = 0;
  RowInfo ri{};
  double y = 0.0;
  if (have) {
    s = eseq[t];
    dv = edup[t];
    ri = f.rows[s];
    y = f.yt[s];
  }
  while (__any_sync(0xffffffffu, have)) {
    const uint32_t tn = t + kWSlots;
    const bool have_n = have && tn < n;
    uint32_t sn = 0, dvn = 0;
    RowInfo rin{};
    double yn = 0.0;
    if (have_n) {
      sn = eseq[tn];
      dvn = edup[tn];
      rin = f.rows[sn];
      yn = f.yt[sn];
    }
    bool pending = have;
    const uint32_t q = q_base + t;
    const uint32_t par = ((q / kWSlots) & 1u) ^ 1u;
    while (__any_sync(0xffffffffu, pending)) {
      if (pending && mbar_try_wait(&sm.empty[lane], par)) {
        WaveSlotMeta m;
        m.s = s;
        m.nnz = ri.nnz;
        m.start = ri.start;
        m.y = y;
        m.dup = dv & 0x7fu;
        m.has_code = dv >> 7;
        sm.meta[lane] = m;
        if (ri.nnz > 0 && ri.nnz <= kCap) {
          const uint32_t bi = static_cast<uint32_t>((ri.nnz + 3) / 4) * 16u;
          const uint32_t bv = static_cast<uint32_t>((ri.nnz + 1) / 2) * 16u;
          const uint32_t bc = m.has_code ? 256u : 0u;
          mbar_expect_tx(&sm.full[lane], bi + bv + bc);
          bulk_g2s(sm.idx[lane], f.ci + ri.start, bi, &sm.full[lane]);
          bulk_g2s(sm.val[lane], f.cv + ri.start, bv, &sm.full[lane]);
          if (bc) bulk_g2s(sm.code[lane], edep + size_t(t) * 32, 256u, &sm.full[lane]);
        } else {
          mbar_arrive(&sm.full[lane]);
        }
        pending = false;
      }
      if (__any_sync(0xffffffffu, pending)) __nanosleep(200);   // the ring is 32 rows deep: no hurry
    }
    t = tn;
    have = have_n;
    s = sn;
    dv = dvn;
    ri = rin;
    y = yn;
  }
}

// ------------------------------------------------------------------ chain warp: the serial scalar recurrence
// lp = dot + b; Gradient (src/families.h:89-96, 161-168); g_change; intercept step (src/saga-sparse.h:300-304).
template <int FAMILY, bool INTERCEPT>
__device__ __noinline__ void wave_chain(WaveSmem& sm, const FitDev& f, const WaveConst& k, uint32_t q_base, int lane,
                                        double& b_io, double& gsi_io) {
  const uint32_t n = k.n;
  const double nd = k.nd, rn = 1.0 / k.nd, gamma = k.gamma;
#ifdef SGD_WAVE_TRACE
  const bool ep_trace = q_base == 0u && lane == 0;
#endif
  double b_reg = b_io, gsi_reg = gsi_io;
  double* __restrict__ gmem = f.gmem;
  const bool lead = lane == 0;
  const uint32_t rt_zero = static_cast<uint32_t>(f.pad0_);   // always 0, but only at run time (see dep_on)
  // shared-window addresses of the queue rings (8-byte entries; q_s has 4-byte entries)
  const uint32_t sb = pin_u32(smem_u32(&sm));
  const uint32_t a_rdy = sb + offsetof(WaveSmem, rdy), a_gok = sb + offsetof(WaveSmem, gok);
  const uint32_t a_dot = sb + offsetof(WaveSmem, q_dot), a_ya = sb + offsetof(WaveSmem, q_ya);
  const uint32_t a_gm = sb + offsetof(WaveSmem, q_gm), a_gch = sb + offsetof(WaveSmem, q_gch);
  const uint32_t a_s = sb + offsetof(WaveSmem, q_s);
  wait_row(sm.rdy, q_base);
  uint32_t o8 = (q_base % kSeq) * 8u;
  double dot = lds_f64(a_dot + o8), ya = lds_f64(a_ya + o8), gm = lds_f64(a_gm + o8);
  uint32_t s = lds_u32(a_s + (o8 >> 1));
#pragma unroll 4
  for (uint32_t t = 0; t < n; ++t) {
    const uint32_t q = q_base + t;
    const uint32_t n8 = ((q + 1u) % kSeq) * 8u, par1 = ((q + 1u) / kSeq) & 1u;
    TRACE(0, t);
    // probe the next row's operands while this row's arithmetic runs (non-blocking; the ring has a spare barrier
    // phase, so probing one row past the epoch's end is harmless)
    const bool next_ready = mbar_test_wait_a(a_rdy + n8, par1);
    // ... and fetch them behind the probe. The loads' address is made to depend on the probe's result: the barrier
    // unit and the load/store unit do not order an independent load behind the probe, and a load that overtakes it
    // can return the previous occupant of the ring slot although the probe then reports "ready".
    const uint32_t n8d = dep_on(n8, next_ready, rt_zero);
    double dot_n = lds_f64(a_dot + n8d), ya_n = lds_f64(a_ya + n8d), gm_n = lds_f64(a_gm + n8d);
    uint32_t s_n = lds_u32(a_s + (n8d >> 1));
    const double lp = dot + b_reg;
    double g;
    if (FAMILY == kBinomial) g = ya - 1.0 / (1.0 + sgd_exp_inrange(lp));   // ya = 1 - y
    else g = lp - ya;
    const double gch = g - gm;
    // every lane holds the same values: the stores need no branch, the arrive carries its predicate
    sts_f64(a_gch + o8, gch);
    gmem[s] = g;
    mbar_arrive_if(lead, a_gok + o8);
    TRACE(1, t);
    TRACEV(9, t, next_ready ? 1 : 0);
    if (INTERCEPT) {
      const double gn = div_by_n(gch, nd, rn);
      gsi_reg += gn;
      b_reg -= gamma * (gsi_reg * 0.01 + gn);
    }
    if (__builtin_expect(!next_ready && t + 1u < n, 0)) {
      mbar_wait_a(a_rdy + n8, par1);
      dot_n = lds_f64(a_dot + n8);
      ya_n = lds_f64(a_ya + n8);
      gm_n = lds_f64(a_gm + n8);
      s_n = lds_u32(a_s + (n8 >> 1));
    }
    o8 = n8;
    dot = dot_n;
    ya = ya_n;
    gm = gm_n;
    s = s_n;
#ifdef SGD_WAVE_TRACE
    if ((t) - kTraceFrom < kTraceRows && ep_trace) g_wave_trace[(t) - kTraceFrom][2] = (dot != 12345.678) ? clock64() : 0;
#endif
  }
  b_io = b_reg;
  gsi_io = gsi_reg;
}

// ------------------------------------------------------------------ worker warp: rows t = warp, warp+S, ...
// PEN = kRidge | kElasticNet; IDENT: alpha*gamma == 0 (lasso or lambda == 0), so wscale stays exactly 1,
// lag_scaling[m] is exactly m, and every "/ wscale" is a division by 1.0 (exact, skipped).
template <int S, int PEN, bool IDENT>
__device__ __noinline__ void wave_worker(WaveSmem& sm, const FitDev& f, const WaveConst& k, uint32_t q_base, int warp,
                                         int lane) {
  const uint32_t n = k.n;
  const double gamma = k.gamma, r = k.r, sc2 = k.sc2, bg = k.bg;
  const double* __restrict__ ls_table = k.ls_table;
  FeatState* __restrict__ st = k.st;
  const int p = k.p;

#ifdef SGD_WAVE_TRACE
  const bool ep_trace = q_base == 0u && lane == 0;
#endif
  double ws = 1.0;        // wscale at the start of step t_sim (before that step's reset test)
  uint32_t t_sim = 0;
  for (uint32_t t = warp; t < n; t += S) {
    const uint32_t q = q_base + t;
    const int slot = static_cast<int>(q % kWSlots);
    const int sq = static_cast<int>(q % kSeq);

    // deterministic wscale track: steps t_sim .. t-1 belong to other workers (all within the window)
    uint32_t force = 0;   // distance to the most recent reset row among them (0 = none)
    bool reset_here = false;
    double ws_next = 1.0, step0 = gamma, sc = -gamma, step1 = gamma, thr1 = bg;   // IDENT: x / 1.0 == x exactly
    if (!IDENT) {
      while (t_sim < t) {
        if (ws < kSmall) {
          ws = 1.0;
          force = t - t_sim;
        }
        ws *= r;
        ++t_sim;
      }
      reset_here = ws < kSmall;
      ws_next = (reset_here ? 1.0 : ws) * r;   // wscale after this step
      step0 = gamma / ws;
      sc = -gamma / ws_next;
      step1 = gamma / ws_next * 1.0;
      thr1 = bg / ws_next;
    }

    mbar_wait(&sm.full[slot], (q / kWSlots) & 1u);
    TRACE(3, t);
    const WaveSlotMeta m = sm.meta[slot];
    const bool serial = reset_here || m.nnz > kCap;
    const double ya = (k.family == kBinomial) ? 1.0 - m.y : m.y;

    if (!serial) {
      // ---- the row and its conflict codes into registers, then give the slot back
      int jr[kChunks];
      bool valid[kChunks];
      double vr[kChunks], wr[kChunks], gr[kChunks];
      uint32_t lr[kChunks], dr[kChunks];   // dr: conflict entry (distance | position << 4 | kCodeGlobal)
      const uint64_t code_raw = sm.code[slot][lane];
      const uint64_t code = m.has_code ? code_raw : 0ull;   // stale ring contents when the row has no codes
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        const int e = c * 32 + lane;
        valid[c] = e < m.nnz;
        jr[c] = valid[c] ? sm.idx[slot][e] : 0;          // clamped: invalid positions gather feature 0 and drop it
        vr[c] = valid[c] ? sm.val[slot][e] : 0.0;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.empty[slot]);
      // ---- gather every position now; the few that a row in flight touches are re-read below
#pragma unroll
      for (int c = 0; c < kChunks; ++c) ld_state(st + jr[c], wr[c], gr[c], lr[c]);
      const double gm_early = f.gmem[m.s];
      // distance to the nearest row in flight that touches the feature; a reset row in the window touches all
      bool late = false;
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        uint32_t d = static_cast<uint32_t>(code >> (16 * c)) & 0xffffu;
        if (!IDENT && force != 0 && (d == 0 || (d & 15u) >= force)) d = force | kCodeGlobal;
        d = valid[c] ? d : 0u;
        dr[c] = d;
        late = late || d != 0;
      }
      // LaggedUpdate(k = t) on one gathered feature (src/saga-sparse.h:76-100, src/penalties.h); the result is only
      // taken when the lag is non-zero, exactly like the reference's `if (lagged_amount != 0)`
      auto caught_up = [&](double w, double g, uint32_t lg) {
        const uint32_t lagged = t - lg;
        const double scal = IDENT ? static_cast<double>(lagged) : ls_table[lagged < n ? lagged : 0u];
        const double bgs = bg * scal;
        const double w_new = penalty_k1<PEN>(w, g, step0 * scal, IDENT ? bgs : bgs / ws);
        return (lagged != 0) ? w_new : w;
      };
#pragma unroll
      for (int c = 0; c < kChunks; ++c) wr[c] = caught_up(wr[c], gr[c], lr[c]);
#ifdef SGD_WAVE_TRACE
      if ((t) - kTraceFrom < kTraceRows && ep_trace) {
        g_wave_trace[t - kTraceFrom][4] = (wr[0] + wr[1] + wr[2] + wr[3] != 12345.678) ? clock64() : 0;
      }
#endif
      double gm = gm_early;
      // ---- rows in flight. A feature that row t-d (d < S) also holds is not read from HBM: as soon as the chain
      // warp has published that row's g_change, this warp repeats the row's own coefficient step on the state it
      // forwarded (same operands, same operations, so the same bits it scatters) and catches the result up.
      // Kept warp-uniform (a warp that diverges around a wait pays hundreds of cycles per later shuffle): the warp
      // waits for every row any lane needs, then each lane swaps the state in by select.
      uint32_t need_g = 0, need_f = 0;   // bit d: g_change of / HBM scatter of row t-d is needed
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        const uint32_t bit = (dr[c] != 0) ? (1u << (dr[c] & 15u)) : 0u;
        need_g |= (dr[c] & kCodeGlobal) ? 0u : bit;
        need_f |= (dr[c] & kCodeGlobal) ? bit : 0u;
      }
      need_g = __reduce_or_sync(0xffffffffu, need_g) | ((m.dup != 0) ? (1u << m.dup) : 0u);
      need_f = __reduce_or_sync(0xffffffffu, need_f);
      TRACEV(8, t, static_cast<long long>(need_g | (need_f << 16)));
      auto forward_stores = [&]() {
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          sm.fw_w[sq][c * 32 + lane] = wr[c];
          sm.fw_g[sq][c * 32 + lane] = gr[c];
          sm.fw_x[sq][c * 32 + lane] = vr[c];
        }
        if (!IDENT && lane == 0) {
          sm.c_sc[sq] = sc;
          sm.c_step1[sq] = step1;
          sm.c_thr1[sq] = thr1;
        }
      };
      // the coefficient step of row t-d redone on its forwarded state for this lane's entries (all of them, or all
      // but those at distance `skip_d`), caught up to this row
      auto apply_forward = [&](uint32_t skip_d) {
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          const uint32_t d = dr[c] & 15u, sqd = (q - d) % kSeq, pos = (dr[c] >> 4) & 127u;
          const bool fwd = dr[c] != 0 && (dr[c] & kCodeGlobal) == 0 && d != skip_d;
          const double gx = sm.fw_x[sqd][pos] * sm.q_gch[sqd];
          const double fg = sm.fw_g[sqd][pos];
          const double w_new = penalty_k1<PEN>(sm.fw_w[sqd][pos] + gx * (IDENT ? sc : sm.c_sc[sqd]), fg,
                                               IDENT ? step1 : sm.c_step1[sqd], IDENT ? thr1 : sm.c_thr1[sqd]);
          const double g_new = fg + gx * sc2;
          wr[c] = fwd ? caught_up(w_new, g_new, t - d + 1u) : wr[c];   // row t-d left lag = t-d+1
          gr[c] = fwd ? g_new : gr[c];
        }
      };
      bool published = false;
#ifndef SGD_NO_FAST_CONFLICT
      if (need_g != 0 && need_f == 0 && m.dup == 0) {
        // ---- fast conflict path. Timeline traces of this kernel show the serial cost of a conflict: from the moment
        // the chain warp publishes the g_change a row waits for to the moment that row's dot product is published
        // took about 1000 cycles (forward step on all four chunks, catch-up, twelve stores, the products and a full
        // butterfly), all of it with the chain warp idle. When ONE position of the row is all that depends on the
        // nearest unfinished row, everything else is done before the wait: the other rows' forwards, the forwarding
        // stores, the running sums, and the butterfly with the late lane contributing nothing - what that lane
        // RECEIVES in the five levels are sums that never include its own value, i.e. exactly the addends of
        // ((((a + r1) + r2) + r3) + r4) + r5, the value every lane of the full butterfly ends with (each level adds
        // the same two numbers, and addition commutes). After the wait only the late lane works: one forward step,
        // at most four additions to close its running sum, five to close the butterfly, and it publishes.
        uint32_t my_near = 16u;
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          const uint32_t d = dr[c] & 15u;
          my_near = (dr[c] != 0 && d < my_near) ? d : my_near;
        }
        const uint32_t dnear = __reduce_min_sync(0xffffffffu, my_near);
        int cnt = 0, c1 = 0;
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          const bool hit = dr[c] != 0 && (dr[c] & 15u) == dnear;
          cnt += hit ? 1 : 0;
          c1 = hit ? c : c1;
        }
        if (__reduce_add_sync(0xffffffffu, static_cast<uint32_t>(cnt)) == 1u) {
          const bool late_lane = cnt == 1;
          for (uint32_t rest = need_g & ~(1u << dnear); rest != 0;) {     // older rows: complete already
            const uint32_t d = 31u - static_cast<uint32_t>(__clz(rest));
            rest &= ~(1u << d);
            wait_row(sm.gok, q - d);
          }
          apply_forward(dnear);
          forward_stores();                      // the late position is stored again below
          double acc = 0.0, prod[kChunks];
#pragma unroll
          for (int c = 0; c < kChunks; ++c) {
            prod[c] = vr[c] * wr[c];
            const double a2 = acc + prod[c];
            acc = (valid[c] && !(late_lane && c >= c1)) ? a2 : acc;
          }
          double v = late_lane ? 0.0 : acc;
          double rc[5];
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            const double x = __shfl_xor_sync(0xffffffffu, v, 16 >> i);
            rc[i] = x;
            v = v + x;
          }
          // the late position's own operands
          double x_late = vr[0];
          uint32_t code1 = dr[0];
#pragma unroll
          for (int c = 0; c < kChunks; ++c) {
            x_late = (c == c1) ? vr[c] : x_late;
            code1 = (c == c1) ? dr[c] : code1;
          }
          const uint32_t sqd = (q - dnear) % kSeq, pos = (code1 >> 4) & 127u;
          __syncwarp();                          // every lane's forwarding stores precede the late lane's arrive
          wait_row(sm.gok, q - dnear);
          if (late_lane) {
            // (that row's forwarded state is only guaranteed once its gok is: it is stored ahead of its rdy)
            const double fwx = sm.fw_x[sqd][pos], fwg = sm.fw_g[sqd][pos], fww = sm.fw_w[sqd][pos];
            const double sc_p = IDENT ? sc : sm.c_sc[sqd];
            const double step1_p = IDENT ? step1 : sm.c_step1[sqd];
            const double thr1_p = IDENT ? thr1 : sm.c_thr1[sqd];
            const double gx = fwx * sm.q_gch[sqd];
            const double w_new = penalty_k1<PEN>(fww + gx * sc_p, fwg, step1_p, thr1_p);
            const double g_new = fwg + gx * sc2;
            // the row right before this one left lag = t: nothing to catch up (caught_up would return w_new itself)
            const double w_c = (dnear == 1u) ? w_new : caught_up(w_new, g_new, t - dnear + 1u);
#pragma unroll
            for (int c = 0; c < kChunks; ++c) {
              wr[c] = (c == c1) ? w_c : wr[c];
              gr[c] = (c == c1) ? g_new : gr[c];
            }
            sm.fw_w[sq][c1 * 32 + lane] = w_c;
            sm.fw_g[sq][c1 * 32 + lane] = g_new;
            double a = acc + x_late * w_c;
#pragma unroll
            for (int c = 1; c < kChunks; ++c) {
              const double a2 = a + prod[c];
              a = (c > c1 && valid[c]) ? a2 : a;
            }
            a = a + rc[0];
            a = a + rc[1];
            a = a + rc[2];
            a = a + rc[3];
            a = a + rc[4];
            sm.q_dot[sq] = IDENT ? a : a * ws;
            sm.q_ya[sq] = ya;
            sm.q_gm[sq] = gm;
            sm.q_s[sq] = m.s;
            mbar_arrive(&sm.rdy[sq]);
          }
          __syncwarp();
          published = true;
        }
      }
#endif
      if (!published) {
        if ((need_g | need_f) != 0) {
          // oldest first: those are complete already, the nearest row is the one worth sleeping on
          for (uint32_t rest = need_f; rest != 0;) {
            const uint32_t d = 31u - static_cast<uint32_t>(__clz(rest));
            rest &= ~(1u << d);
            wait_row(sm.fdone, q - d);
          }
          for (uint32_t rest = need_g; rest != 0;) {
            const uint32_t d = 31u - static_cast<uint32_t>(__clz(rest));
            rest &= ~(1u << d);
            wait_row(sm.gok, q - d);
          }
          if (m.dup != 0) gm = f.gmem[m.s];
          apply_forward(16u);
          if (need_f != 0) {                          // behind a serial row (rare): re-read HBM
#pragma unroll
            for (int c = 0; c < kChunks; ++c) {
              if (dr[c] & kCodeGlobal) {
                ld_state(st + jr[c], wr[c], gr[c], lr[c]);
                wr[c] = caught_up(wr[c], gr[c], lr[c]);
              }
            }
            __syncwarp();
          }
        }
        // ---- forward this row's caught-up state (read by the rows that conflict with it, after its gok)
        forward_stores();
        // ---- the sparse dot product: position e -> running sum e mod 32, then the butterfly (sgdnet_arith.h)
        double acc = 0.0;
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          const double a2 = acc + vr[c] * wr[c];
          acc = valid[c] ? a2 : acc;
        }
        double dot = warp_sum(acc);
        if (!IDENT) dot = dot * ws;
        __syncwarp();                                 // the forwarded state of every lane precedes the arrive
        if (lane == 0) {
          sm.q_dot[sq] = dot;
          sm.q_ya[sq] = ya;
          sm.q_gm[sq] = gm;
          sm.q_s[sq] = m.s;
          mbar_arrive(&sm.rdy[sq]);
        }
      }
      TRACE(5, t);
      wait_row(sm.gok, q);
      const double gch = sm.q_gch[sq];
      TRACE(6, t);
      // ---- write-after-write order in HBM: a feature this row shares with a row still in flight is scattered by
      // both, and the earlier row's record has to land first. `done` is chained in row order, so the nearest needed
      // row's `done` covers every needed row; it has normally completed long before (one probe, behind the publish, off
      // the chain warp's path). Found with the fast conflict path: without this wait the order only held because a
      // conflicting row used to publish its dot product ~1000 cycles after the g_change of the row it waited for.
      if (need_g != 0) wait_row(sm.done, q - static_cast<uint32_t>(__ffs(static_cast<int>(need_g)) - 1));
      // ---- AddWeighted(w), LaggedUpdate(k = t+1, lag 1), AddWeighted(g_sum), scatter
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        const double gx = vr[c] * gch;
        const double w = penalty_k1<PEN>(wr[c] + gx * sc, gr[c], step1, thr1);
        st_state_if(valid[c], st + jr[c], w, gr[c] + gx * sc2, t + 1u);
      }
    } else {
      // ---- serial row: every earlier row is complete before anything is read; operands are used in place
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.empty[slot]);
      if (t > 0) wait_row(sm.done, q - 1u);
      const double gm = f.gmem[m.s];
      const int32_t* __restrict__ ci = f.ci + m.start;
      const double* __restrict__ cv = f.cv + m.start;
      double acc = 0.0;
      for (int e = lane; e < m.nnz; e += 32) {
        const int j = ci[e];
        double w, gs;
        uint32_t lg;
        ld_state(st + j, w, gs, lg);
        const uint32_t lagged = t - lg;
        if (lagged != 0) {
          const double scal = IDENT ? static_cast<double>(lagged) : ls_table[lagged];
          const double bgs = bg * scal;
          w = penalty_k1<PEN>(w, gs, step0 * scal, IDENT ? bgs : bgs / ws);
          st_state(st + j, w, gs, t);
        }
        acc += cv[e] * w;
      }
      double dot = warp_sum(acc);
      if (!IDENT) dot = dot * ws;
      if (lane == 0) {
        sm.q_dot[sq] = dot;
        sm.q_ya[sq] = ya;
        sm.q_gm[sq] = gm;
        sm.q_s[sq] = m.s;
        mbar_arrive(&sm.rdy[sq]);
      }
      if (reset_here) {
        // Reset(t) over all features, lag = t (src/saga-sparse.h:285-295)
        __syncwarp();
        for (int j = lane; j < p; j += 32) {
          double w, gs;
          uint32_t lg;
          ld_state(st + j, w, gs, lg);
          const uint32_t lagged = t - lg;
          if (lagged != 0) {
            const double scal = ls_table[lagged];
            const double bgs = bg * scal;
            w = penalty_k1<PEN>(w, gs, step0 * scal, bgs / ws);
          }
          st_state(st + j, w * ws, gs, t);
        }
        __syncwarp();
      }
      wait_row(sm.gok, q);
      const double gch = sm.q_gch[sq];
      for (int e = lane; e < m.nnz; e += 32) {
        const int j = ci[e];
        const double gx = cv[e] * gch;
        double w, gs;
        uint32_t lg;
        ld_state(st + j, w, gs, lg);
        w = penalty_k1<PEN>(w + gx * sc, gs, step1, thr1);
        st_state(st + j, w, gs + gx * sc2, t + 1u);
      }
    }
    // ---- completion chained in row order: the row's HBM scatter is visible and so is every earlier row's
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(&sm.fdone[sq]);
      if (t > 0) wait_row(sm.done, q - 1u);
      mbar_arrive(&sm.done[sq]);
    }
    TRACE(7, t);
    ws = ws_next;
    t_sim = t + 1u;
  }
  // wscale after the epoch's last step (identical in every worker)
  if (!IDENT) {
    while (t_sim < n) {
      if (ws < kSmall) ws = 1.0;
      ws *= r;
      ++t_sim;
    }
  }
  if (warp == 0 && lane == 0) sm.wscale_s = IDENT ? 1.0 : ws;
}

// One CTA per fit. The fits of a batch may differ in family, penalty and in whether wscale moves (cv alpha grids), so
// the specialisations are chosen per CTA (and per role) at run time.
// Warp roles. A warp's scheduler (SM sub-partition) is its id mod 4. The chain warp's instruction stream is the
// critical path of the whole fit, so it gets a scheduler to itself: it is warp 3 and the other warps of that
// sub-partition (7, 11, ...) are idle placeholders that only take part in the block barriers and the epoch-end
// sweep (measured: 648 -> 634 cycles per update on config 2's shape, 565 -> 539 without row conflicts).
// Role index r = rank among the remaining warps: r < S workers, r == S the producer.
constexpr int wave_block_warps(int S) { return (S + 1) + (S + 1 + 2) / 3; }   // S+1 role warps on 3 of every 4 ids
__device__ __forceinline__ int wave_role(int warp, int S) {   // -1 chain, -2 idle, else role index
  if ((warp & 3) == 3) return warp == 3 ? -1 : -2;
  const int r = warp - (warp >> 2);
  return r <= S ? r : -2;
}

template <int S>
__global__ void __launch_bounds__(wave_block_warps(S) * 32, 1)
saga_sparse_wave_kernel(FitDev* __restrict__ fit, Progress* __restrict__ prog, const RoundArgs ra) {
  extern __shared__ __align__(128) unsigned char wave_smem_raw[];
  WaveSmem& sm = *reinterpret_cast<WaveSmem*>(wave_smem_raw);

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

  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int role = wave_role(warp, S);
  const int li = pg.lambda_ind;
  WaveConst k;
  k.n = static_cast<uint32_t>(f.n);
  k.p = f.p;
  k.family = f.family;
  k.fit_intercept = f.fit_intercept != 0;
  k.nd = static_cast<double>(k.n);
  k.gamma = f.gamma[li];
  const double alpha = f.alpha[li], beta = f.beta[li];
  k.r = 1.0 - alpha * k.gamma;               // wscale_update
  k.sc2 = 1.0 / k.nd;
  k.bg = beta * k.gamma;                     // (beta*gamma)*1.0
  k.ls_table = f.lag_scaling;
  k.st = f.st;
  const bool ident = (k.r == 1.0);
  const int pen = f.penalty;
  const uint32_t n = k.n;
  const int p = k.p;

  if (tid == 0) {
    for (int i = 0; i < kWSlots; ++i) {
      mbar_init(&sm.full[i], 1);
      mbar_init(&sm.empty[i], 1);
    }
    for (int i = 0; i < kSeq; ++i) {
      mbar_init(&sm.rdy[i], 1);
      mbar_init(&sm.gok[i], 1);
      mbar_init(&sm.fdone[i], 1);
      mbar_init(&sm.done[i], 1);
    }
    fence_barrier_init();
  }
  __syncthreads();

  double b_reg = f.b[0], gsi_reg = f.gsi[0];   // live in the chain warp

  uint32_t it_outer = pg.it_outer, epochs_done = 0;
  bool finished = false;
  uint32_t q_base = 0;      // sequence number (mod 2^32) of the first row of the current epoch

  for (int ep = 0; ep < ra.n_epochs && !finished; ++ep, q_base += n) {
    if (role == S) {
      wave_producer(sm, f, ra, ep, q_base, n, lane);
    } else if (role == -2) {
      // idle placeholder of the chain warp's scheduler
    } else if (role == -1) {
      if (k.family == kBinomial) {
        if (k.fit_intercept) wave_chain<kBinomial, true>(sm, f, k, q_base, lane, b_reg, gsi_reg);
        else wave_chain<kBinomial, false>(sm, f, k, q_base, lane, b_reg, gsi_reg);
      } else {
        if (k.fit_intercept) wave_chain<kGaussian, true>(sm, f, k, q_base, lane, b_reg, gsi_reg);
        else wave_chain<kGaussian, false>(sm, f, k, q_base, lane, b_reg, gsi_reg);
      }
    } else {
      if (pen == kRidge) {
        if (ident) wave_worker<S, kRidge, true>(sm, f, k, q_base, role, lane);
        else wave_worker<S, kRidge, false>(sm, f, k, q_base, role, lane);
      } else {
        if (ident) wave_worker<S, kElasticNet, true>(sm, f, k, q_base, role, lane);
        else wave_worker<S, kElasticNet, false>(sm, f, k, q_base, role, lane);
      }
    }
    __syncthreads();

    // ---- epoch end: Reset(n), unscale, lag = 0, convergence (src/saga-sparse.h:340-348, 367)
    const double wscale = sm.wscale_s;
    const double step_e = k.gamma / wscale;
    double mc = 0.0, ms = 0.0;
    for (int j = tid; j < p; j += T) {
      double w, gs;
      uint32_t lg;
      ld_state(k.st + j, w, gs, lg);
      const uint32_t lagged = n - lg;
      if (lagged != 0) {
        const double scal = ident ? static_cast<double>(lagged) : k.ls_table[lagged];
        const double bgs = k.bg * scal;
        const double v = w - (step_e * scal) * gs;
        w = (pen == kElasticNet) ? soft_threshold(v, bgs / wscale) : v;
      }
      w = w * wscale;
      st_state(k.st + j, w, gs, 0u);
      f.W[j] = w;
      mc = fmax(mc, fabs(w - f.Wprev[j]));
      ms = fmax(ms, fabs(w));
      f.Wprev[j] = w;
    }
    const bool conv = block_converged(mc, ms, sm.red, f.tol);
    ++it_outer;
    ++epochs_done;
    finished = !free_run && (conv || !(it_outer < f.max_iter));
  }

  if (role == -1 && lane == 0) {
    f.b[0] = b_reg;
    f.gsi[0] = gsi_reg;
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

// ============================================================================================ generic path
constexpr int kGenThreads = 256;

__global__ void __launch_bounds__(kGenThreads, 1)
saga_sparse_generic_kernel(FitDev* __restrict__ fit, Progress* __restrict__ prog, const RoundArgs ra) {
  __shared__ double red[2][2 * 32 * (kGenThreads / 32)];   // [parity][warp][2K]  (K <= 32)
  __shared__ double gch_s[32];
  __shared__ double cred[2 * (kGenThreads / 32)];

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
    pg.solver_ns += globaltimer_ns() - t_start;
    publish_progress(f.mirror, pg, ra.round_id);
  }
}


int wave_warps() {
  static int s = [] {
    int v = 8;
    if (const char* env = std::getenv("SGDNET_WAVE_WARPS")) v = std::atoi(env);
    return (v == 4 || v == 8 || v == 12) ? v : 8;
  }();
  return s;
}

template <int S>
static cudaError_t launch_wave(FitDev* fit, Progress* prog, const RoundArgs& ra, cudaStream_t st) {
  // per device and cheap: set on every launch (the ABI lets one process move between devices, sgdnet_set_device)
  cudaError_t e = cudaFuncSetAttribute(saga_sparse_wave_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(sizeof(WaveSmem)));
  if (e != cudaSuccess) return e;
  saga_sparse_wave_kernel<S><<<1, wave_block_warps(S) * 32, sizeof(WaveSmem), st>>>(fit, prog, ra);
  return cudaGetLastError();
}

// `ctas`: how many CTAs this fit's conflict-code pass may use (the caller shares the GPU among the fits in flight)
cudaError_t launch_wave_deps(const FitDev* fit, const RoundArgs& ra, int64_t rows, int ctas, cudaStream_t st) {
  const int window = wave_warps() - 1;
  const int own = kDepSlots - window;
  const int64_t want = (rows + own - 1) / own;
  dim3 grid(static_cast<unsigned>(std::max<int64_t>(1, std::min<int64_t>(want, ctas))));
  const size_t smem = sizeof(int32_t) * kDepSlots * kCap + sizeof(uint32_t) * (kDepBuckets + 2 * kDepSlots + 4);   // 48.3 KB
  cudaError_t e = cudaFuncSetAttribute(wave_deps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  if (e != cudaSuccess) return e;
  wave_deps_kernel<<<grid, kDepWarps * 32, smem, st>>>(fit, ra, window);
  return cudaGetLastError();
}

cudaError_t launch_saga_sparse(bool fast_k1, FitDev* fit, Progress* prog, const RoundArgs& ra, cudaStream_t st) {
  if (fast_k1) {
    switch (wave_warps()) {
      case 4: return launch_wave<4>(fit, prog, ra, st);
      case 12: return launch_wave<12>(fit, prog, ra, st);
      default: return launch_wave<8>(fit, prog, ra, st);
    }
  }
  saga_sparse_generic_kernel<<<1, kGenThreads, 0, st>>>(fit, prog, ra);
  return cudaGetLastError();
}

}  // namespace sgd
