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
#include <algorithm>
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
//  - fdone[q]  completes when row q's scatter is visible (what a conflicting later row waits for).
//  - done[q]   like fdone but chained in row order: "done(q)" means every row <= q is complete, which bounds the rows
//              in flight to S consecutive ones (the window wave_deps_kernel looked at).
//  - a row whose nonzeros do not fit a ring slot, or at which the wscale reset (src/saga-sparse.h:285-295) fires,
//    is run serially: it waits for every earlier row, and the rows after it wait for it.
constexpr int kWSlots = 32;     // row ring depth: S rows held by the workers + rows in flight from HBM
constexpr int kSeq = 32;        // per-row barrier / queue rings (indexed by row sequence number)
static_assert(kChunks == 4, "conflict codes pack four 4-bit distances per lane");

struct WaveSlotMeta {
  uint32_t s;
  int32_t nnz;
  int64_t start;
  double y;
  uint32_t dup;     // distance to the most recent in-window row with the same sample (0 = none)
  uint32_t pad_;
};

struct __align__(128) WaveSmem {
  double val[kWSlots][kCap];
  int32_t idx[kWSlots][kCap];
  uint16_t code[kWSlots][32];
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
  unsigned long long a, b, c, d;
  asm volatile("ld.global.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p) : "memory");
  w = __longlong_as_double(static_cast<long long>(a));
  g = __longlong_as_double(static_cast<long long>(b));
  lag = static_cast<uint32_t>(c);
}
__device__ __forceinline__ void st_state(FeatState* p, double w, double g, uint32_t lag) {
  asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(__double_as_longlong(w)), "l"(__double_as_longlong(g)),
               "l"(static_cast<unsigned long long>(lag)), "l"(0ull)
               : "memory");
}

__device__ __forceinline__ void wait_row(uint64_t* ring, int64_t q) {
  mbar_wait(&ring[q % kSeq], static_cast<uint32_t>((q / kSeq) & 1));
}

// gch / n with n an integer below 2^32: q = a*(1/n) corrected once with the exact residual is the correctly rounded
// quotient (the residual step leaves a relative error of 2^-104 while a/n stays 2^-86 away from every rounding
// boundary because n has at most 32 significant bits), so this returns the same bits as the IEEE division the
// reference performs. Operands outside the safe exponent range (and zero, to keep its sign) take the division.
__device__ __forceinline__ double div_by_n(double a, double nd, double rn) {
  const double aa = fabs(a);
  if (aa > 1e-270 && aa < 1e270) {
    const double q0 = a * rn;
    const double r0 = fma(-q0, nd, a);
    return fma(r0, rn, q0);
  }
  return a / nd;
}

// ---- conflict codes: for every row instance q = epoch*n + t of the staged sequence and every nonzero position e of
// its row, the distance d in 1..window to the most recent earlier row OF THE SAME EPOCH that holds the same feature
// (0 = none within the window). A row too long for a ring slot conflicts with everything. Lane l of the solver reads
// one 16-bit word holding the codes of positions l, l+32, l+64, l+96.
__global__ void __launch_bounds__(256)
wave_deps_kernel(const FitDev* __restrict__ fits, const Progress* __restrict__ prog, const RoundArgs* __restrict__ args,
                 int window) {
  const int fit_id = blockIdx.y;
  const RoundArgs ra = args[fit_id];
  if (ra.n_epochs <= 0 || prog[fit_id].status != kRunning || ra.dep == nullptr) return;
  const FitDev& f = fits[fit_id];
  const int64_t n = f.n;
  const int64_t total = n * ra.n_epochs;
  const int lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int64_t warp_global = int64_t(blockIdx.x) * nwarps + (threadIdx.x >> 5);
  const int64_t warp_stride = int64_t(gridDim.x) * nwarps;
  const int32_t* __restrict__ ci = f.ci;
  for (int64_t q = warp_global; q < total; q += warp_stride) {
    const int64_t t = q % n;
    const uint32_t* __restrict__ eseq = ra.seq + (q - t);
    const uint32_t s = eseq[t];
    const RowInfo ri = f.rows[s];
    int j[kChunks];
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const int e = c * 32 + lane;
      j[c] = (ri.nnz <= kCap && e < ri.nnz) ? ci[ri.start + e] : -1;
    }
    uint32_t code = 0, dupd = 0;
    const int dmax = static_cast<int>(t < window ? t : window);
    for (int d = 1; d <= dmax; ++d) {
      const uint32_t s2 = eseq[t - d];
      if (s2 == s && dupd == 0) dupd = static_cast<uint32_t>(d);
      const RowInfo r2 = f.rows[s2];
      if (r2.nnz == 0) continue;
      const int32_t* __restrict__ c2 = ci + r2.start;
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        if (j[c] < 0 || ((code >> (4 * c)) & 15u) != 0) continue;
        bool hit = r2.nnz > kCap;
        if (!hit) {
          int lo = 0, hi = r2.nnz;           // first position with c2[pos] >= j[c]
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (c2[mid] < j[c]) lo = mid + 1; else hi = mid;
          }
          hit = lo < r2.nnz && c2[lo] == j[c];
        }
        if (hit) code |= static_cast<uint32_t>(d) << (4 * c);
      }
    }
    ra.dep[q * 32 + lane] = static_cast<uint16_t>(code);
    if (lane == 0) ra.dup[q] = static_cast<uint8_t>(dupd);
  }
}

template <int S>
__global__ void __launch_bounds__((S + 2) * 32, 1)
saga_sparse_wave_kernel(FitDev* __restrict__ fits, Progress* __restrict__ prog, const RoundArgs* __restrict__ args) {
  extern __shared__ __align__(128) unsigned char wave_smem_raw[];
  WaveSmem& sm = *reinterpret_cast<WaveSmem*>(wave_smem_raw);
  constexpr int kProducer = S, kChain = S + 1;   // warp roles; the chain warp has the highest warp id of its scheduler

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
  const bool identity = (r == 1.0);          // lasso: wscale == 1 and lag_scaling[m] == m exactly
  const double sc2 = 1.0 / nd;
  const double bg = beta * gamma;            // (beta*gamma)*1.0

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

  FeatState* __restrict__ st = f.st;
  double b_reg = f.b[0], gsi_reg = f.gsi[0];   // live in the chain warp

  uint32_t it_outer = pg.it_outer, epochs_done = 0;
  bool finished = false;
  int64_t q_base = 0;       // sequence number of the first row of the current epoch

  for (int ep = 0; ep < ra.n_epochs && !finished; ++ep, q_base += n) {
    const uint32_t* __restrict__ eseq = ra.seq + size_t(ep) * n;

    if (warp == kProducer) {
      // ------------------------------------------------------------------ producer: lane l feeds ring slot l
      // Warp-synchronous: the lanes load the sample index, row descriptor, response and duplicate code of their next
      // row together (32 independent requests in flight), one batch ahead of the batch being issued, then the warp
      // polls the slots' "empty" barriers and issues each row's bulk copies as soon as its slot is released.
      const uint8_t* __restrict__ edup = ra.dup + size_t(ep) * n;
      const uint16_t* __restrict__ edep = ra.dep + size_t(ep) * n * 32;
      int64_t t = (lane - static_cast<int>(q_base % kWSlots) + kWSlots) % kWSlots;   // q = q_base + t lands on slot `lane`
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
        const int64_t tn = t + kWSlots;
        const bool have_n = tn < n;
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
        const int64_t q = q_base + t;
        const uint32_t par = static_cast<uint32_t>(((q / kWSlots) & 1) ^ 1);
        while (__any_sync(0xffffffffu, pending)) {
          if (pending && mbar_try_wait(&sm.empty[lane], par)) {
            WaveSlotMeta m;
            m.s = s;
            m.nnz = ri.nnz;
            m.start = ri.start;
            m.y = y;
            m.dup = dv;
            m.pad_ = 0;
            sm.meta[lane] = m;
            if (ri.nnz > 0 && ri.nnz <= kCap) {
              const uint32_t bi = static_cast<uint32_t>((ri.nnz + 3) / 4) * 16u;
              const uint32_t bv = static_cast<uint32_t>((ri.nnz + 1) / 2) * 16u;
              mbar_expect_tx(&sm.full[lane], bi + bv + 64u);
              bulk_g2s(sm.idx[lane], f.ci + ri.start, bi, &sm.full[lane]);
              bulk_g2s(sm.val[lane], f.cv + ri.start, bv, &sm.full[lane]);
              bulk_g2s(sm.code[lane], edep + size_t(t) * 32, 64u, &sm.full[lane]);
            } else {
              mbar_arrive(&sm.full[lane]);
            }
            pending = false;
          }
        }
        t = tn;
        have = have_n;
        s = sn;
        dv = dvn;
        ri = rin;
        y = yn;
      }
    } else if (warp == kChain) {
      // ------------------------------------------------------------------ chain warp: the serial scalar recurrence
      const double rn = 1.0 / nd;
      for (int64_t t = 0; t < n; ++t) {
        const int64_t q = q_base + t;
        const int sq = static_cast<int>(q % kSeq);
        wait_row(sm.rdy, q);
        const double dot = sm.q_dot[sq];
        const double ya = sm.q_ya[sq];
        const double gm = sm.q_gm[sq];
        const uint32_t s = sm.q_s[sq];
        const double lp = dot + b_reg;
        // Gradient (src/families.h:89-96, 161-168); ya = 1 - y for the binomial family
        const double g = (family == kBinomial) ? ya - 1.0 / (1.0 + sgd_exp(lp)) : lp - ya;
        const double gch = g - gm;
        if (lane == 0) {
          sm.q_gch[sq] = gch;
          f.gmem[s] = g;
          mbar_arrive(&sm.gok[sq]);
        }
        if (fit_intercept) {
          const double gn = div_by_n(gch, nd, rn);
          gsi_reg += gn;
          b_reg -= gamma * (gsi_reg * 0.01 + gn);
        }
      }
    } else {
      // ------------------------------------------------------------------ worker warp: rows t = warp, warp+S, ...
      double ws = 1.0;        // wscale at the start of step t_sim (before that step's reset test)
      int64_t t_sim = 0;
      for (int64_t t = warp; t < n; t += S) {
        const int64_t q = q_base + t;
        const int slot = static_cast<int>(q % kWSlots);
        const int sq = static_cast<int>(q % kSeq);
        const uint32_t t32 = static_cast<uint32_t>(t);

        // deterministic wscale track: steps t_sim .. t-1 belong to other workers (all within the window)
        int force = 0;        // distance to the most recent reset row among them (0 = none)
        if (!identity) {
          while (t_sim < t) {
            if (ws < kSmall) {
              ws = 1.0;
              force = static_cast<int>(t - t_sim);
            }
            ws *= r;
            ++t_sim;
          }
        }
        const bool reset_here = !identity && ws < kSmall;

        mbar_wait(&sm.full[slot], static_cast<uint32_t>((q / kWSlots) & 1));
        const WaveSlotMeta m = sm.meta[slot];
        const bool serial = reset_here || m.nnz > kCap;
        const double ya = (family == kBinomial) ? 1.0 - m.y : m.y;
        double ws_next;       // wscale after this step

        if (!serial) {
          // ---- the row and its conflict codes into registers, then give the slot back
          int jr[kChunks];
          double vr[kChunks], wr[kChunks], gr[kChunks];
          uint32_t lr[kChunks];
          uint32_t code = (m.nnz > 0) ? sm.code[slot][lane] : 0u;
#pragma unroll
          for (int c = 0; c < kChunks; ++c) {
            const int e = c * 32 + lane;
            jr[c] = (e < m.nnz) ? sm.idx[slot][e] : -1;
            vr[c] = (e < m.nnz) ? sm.val[slot][e] : 0.0;
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.empty[slot]);
          // a reset row in the window is a conflict on every feature
          bool any_late = false;
#pragma unroll
          for (int c = 0; c < kChunks; ++c) {
            uint32_t d = (code >> (4 * c)) & 15u;
            if (force != 0 && (d == 0 || d > static_cast<uint32_t>(force))) d = static_cast<uint32_t>(force);
            if (jr[c] < 0) d = 0;
            code = (code & ~(15u << (4 * c))) | (d << (4 * c));
            any_late = any_late || d != 0;
          }
          // ---- early gathers: features no row in flight touches
#pragma unroll
          for (int c = 0; c < kChunks; ++c)
            if (jr[c] >= 0 && ((code >> (4 * c)) & 15u) == 0) ld_state(st + jr[c], wr[c], gr[c], lr[c]);
          if (m.dup != 0) wait_row(sm.fdone, q - m.dup);
          const double gm = f.gmem[m.s];
          // ---- late gathers: each waits for the nearest row that touches its feature
          if (any_late) {
#pragma unroll
            for (int c = 0; c < kChunks; ++c) {
              const uint32_t d = (code >> (4 * c)) & 15u;
              if (d != 0) {
                wait_row(sm.fdone, q - d);
                ld_state(st + jr[c], wr[c], gr[c], lr[c]);
              }
            }
          }
          // ---- LaggedUpdate(k = t) and the sparse dot product
          const double step0 = gamma / ws;
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
                pc.thr = pc.bgs / ws;
                pc.w_scale = ws;
                wr[c] = penalty_scalar(pen, wr[c], gr[c], pc);
              }
              acc += vr[c] * wr[c];
            }
          }
          const double dot = warp_sum(acc) * ws;
          if (lane == 0) {
            sm.q_dot[sq] = dot;
            sm.q_ya[sq] = ya;
            sm.q_gm[sq] = gm;
            sm.q_s[sq] = m.s;
            mbar_arrive(&sm.rdy[sq]);
          }
          // everything of the coefficient step that does not depend on the gradient
          ws_next = ws * r;
          const double sc = -gamma / ws_next;
          PenCoef pc1;
          pc1.step = gamma / ws_next * 1.0;
          pc1.bgs = bg;
          pc1.thr = bg / ws_next;
          pc1.w_scale = ws_next;
          wait_row(sm.gok, q);
          const double gch = sm.q_gch[sq];
          // ---- AddWeighted(w), LaggedUpdate(k = t+1, lag 1), AddWeighted(g_sum), scatter
#pragma unroll
          for (int c = 0; c < kChunks; ++c) {
            if (jr[c] >= 0) {
              const double gx = vr[c] * gch;
              double w = wr[c] + gx * sc;
              w = penalty_scalar(pen, w, gr[c], pc1);
              st_state(st + jr[c], w, gr[c] + gx * sc2, t32 + 1u);
            }
          }
        } else {
          // ---- serial row: every earlier row is complete before anything is read; operands are used in place
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.empty[slot]);
          if (q > 0) wait_row(sm.done, q - 1);
          const double gm = f.gmem[m.s];
          const int32_t* __restrict__ ci = f.ci + m.start;
          const double* __restrict__ cv = f.cv + m.start;
          double acc = 0.0;
          for (int e = lane; e < m.nnz; e += 32) {
            const int j = ci[e];
            double w, gs;
            uint32_t lg;
            ld_state(st + j, w, gs, lg);
            const uint32_t lagged = t32 - lg;
            if (lagged != 0) {
              w = penalty_scalar(pen, w, gs, pen_coef(gamma, beta, ws, lag_scale(identity, f.lag_scaling, lagged)));
              st_state(st + j, w, gs, t32);
            }
            acc += cv[e] * w;
          }
          const double dot = warp_sum(acc) * ws;
          if (lane == 0) {
            sm.q_dot[sq] = dot;
            sm.q_ya[sq] = ya;
            sm.q_gm[sq] = gm;
            sm.q_s[sq] = m.s;
            mbar_arrive(&sm.rdy[sq]);
          }
          double wcur = ws;
          if (reset_here) {
            // Reset(t) over all features, lag = t (src/saga-sparse.h:285-295)
            __syncwarp();
            for (int j = lane; j < p; j += 32) {
              double w, gs;
              uint32_t lg;
              ld_state(st + j, w, gs, lg);
              const uint32_t lagged = t32 - lg;
              if (lagged != 0)
                w = penalty_scalar(pen, w, gs, pen_coef(gamma, beta, wcur, lag_scale(identity, f.lag_scaling, lagged)));
              st_state(st + j, w * wcur, gs, t32);
            }
            __syncwarp();
            wcur = 1.0;
          }
          ws_next = wcur * r;
          const double sc = -gamma / ws_next;
          const PenCoef pc1 = pen_coef(gamma, beta, ws_next, 1.0);
          wait_row(sm.gok, q);
          const double gch = sm.q_gch[sq];
          for (int e = lane; e < m.nnz; e += 32) {
            const int j = ci[e];
            const double gx = cv[e] * gch;
            double w, gs;
            uint32_t lg;
            ld_state(st + j, w, gs, lg);
            w = w + gx * sc;
            w = penalty_scalar(pen, w, gs, pc1);
            st_state(st + j, w, gs + gx * sc2, t32 + 1u);
          }
        }
        // ---- completion: this row's scatter is visible; then the same, chained in row order
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&sm.fdone[sq]);
          if (q > 0) wait_row(sm.done, q - 1);
          mbar_arrive(&sm.done[sq]);
        }
        ws = ws_next;
        t_sim = t + 1;
      }
      // wscale after the epoch's last step (identical in every worker)
      if (!identity) {
        while (t_sim < n) {
          if (ws < kSmall) ws = 1.0;
          ws *= r;
          ++t_sim;
        }
      }
      if (tid == 0) sm.wscale_s = ws;
    }
    __syncthreads();

    // ---- epoch end: Reset(n), unscale, lag = 0, convergence (src/saga-sparse.h:340-348, 367)
    const double wscale = sm.wscale_s;
    double mc = 0.0, ms = 0.0;
    for (int j = tid; j < p; j += T) {
      double w, gs;
      uint32_t lg;
      ld_state(st + j, w, gs, lg);
      const uint32_t lagged = n32 - lg;
      if (lagged != 0)
        w = penalty_scalar(pen, w, gs, pen_coef(gamma, beta, wscale, lag_scale(identity, f.lag_scaling, lagged)));
      w = w * wscale;
      st_state(st + j, w, gs, 0u);
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

  if (warp == kChain && lane == 0) {
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

int wave_warps() {
  static int s = [] {
    int v = 8;
    if (const char* env = std::getenv("SGDNET_WAVE_WARPS")) v = std::atoi(env);
    return (v == 4 || v == 6 || v == 8 || v == 12) ? v : 8;
  }();
  return s;
}

template <int S>
static cudaError_t launch_wave(int n_fits, FitDev* fits, Progress* prog, const RoundArgs* args, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(saga_sparse_wave_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(sizeof(WaveSmem)));
    if (e != cudaSuccess) return e;
    configured = true;
  }
  saga_sparse_wave_kernel<S><<<n_fits, (S + 2) * 32, sizeof(WaveSmem), st>>>(fits, prog, args);
  return cudaGetLastError();
}

cudaError_t launch_wave_deps(int n_fits, const FitDev* fits, const Progress* prog, const RoundArgs* args,
                             int64_t max_rows, int sms, cudaStream_t st) {
  const int64_t want = (max_rows + 7) / 8;
  const int64_t cap = std::max<int64_t>(1, int64_t(sms) * 8 / std::max(1, n_fits));
  dim3 grid(static_cast<unsigned>(std::max<int64_t>(1, std::min(want, cap))), n_fits);
  wave_deps_kernel<<<grid, 256, 0, st>>>(fits, prog, args, wave_warps() - 1);
  return cudaGetLastError();
}

cudaError_t launch_saga_sparse(int n_fits, bool fast_k1, FitDev* fits, Progress* prog, const RoundArgs* args,
                               cudaStream_t st) {
  if (fast_k1) {
    switch (wave_warps()) {
      case 4: return launch_wave<4>(n_fits, fits, prog, args, st);
      case 6: return launch_wave<6>(n_fits, fits, prog, args, st);
      case 12: return launch_wave<12>(n_fits, fits, prog, args, st);
      default: return launch_wave<8>(n_fits, fits, prog, args, st);
    }
  }
  saga_sparse_generic_kernel<<<n_fits, kGenThreads, 0, st>>>(fits, prog, args);
  return cudaGetLastError();
}

}  // namespace sgd
