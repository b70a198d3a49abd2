// saga_dense_cluster.cu — dense SAGA epochs on a thread-block cluster (reference: src/saga-dense.h:147-212), for
// designs with p >= SGD_WIDE_P features (BASELINE configs 3 and 4).
//
// One cluster per fit: one CTA per 256-feature block of the design (2, 4 or 8 of them; blocks a narrow design does not
// have are exact zeros in the specified sum and are simply absent), times the class split described at the kernel. A CTA
// is 8 feature warps (256 lanes) + 1 control warp. Feature lane L = 256 * block + tid owns the features j = L, L + 2048,
// ... for the CTA's classes. Their W and g_sum stay in that lane's REGISTERS for the whole launch when slices x classes
// <= 16 (configs 3 and 4), otherwise in the CTA's shared memory; nobody else touches them, so the dense sweeps of the
// reference's update (the K x p matrix-vector product, the coefficient step, the prox over all p features and the
// g_sum update, src/saga-dense.h:154, 176-183) are split over up to 2048 lanes. What crosses threads per update is the
// K partial dot products, with the association of include/sgdnet_arith.h item 2 (= the oracle's dot_dense_wide):
//   lane: running sum over its features (ascending j)
//   warp: the xor-butterfly 16, 8, 4, 2, 1 - computed by recursive halving: at offset o a lane keeps one half of its
//         class sums and hands the other half to lane ^ o, so a level moves K/2, K/4, ... values instead of K (the
//         SHFL unit, one warp-instruction per cycle per SM, is what bounded the plain butterfly); each pair sum
//         a_i + a_(i^o) is formed once, by either partner - same operands, same bits (tests/test_arith_cpu.py)
//   CTA:  the 8 warp sums added in ascending order by the control warp (class k in lane k)
//   cluster: the CTA sums exchanged all-to-all through DISTRIBUTED SHARED MEMORY - st.async into the other CTAs'
//         shared memory, completing bytes on their mbarrier, so data and signal travel together and neither side
//         fences; the (class, destination) pairs are spread over the control warp's lanes - and added in ascending
//         block order by every control warp, which then runs the K-value gradient step redundantly: no second
//         exchange. CTA 0 alone reads and writes the gradient memory; its value rides along.
// Inside a CTA the two hand-overs per update are named barriers used one way (bar.arrive by the producer side,
// bar.sync by the consumer side): the feature warps never wait for "warp sums taken", the control warp never waits
// for "g_change taken". Everything that is not on the path warp sums -> g_change (intercept step, gradient-memory
// store, next sample's operands, the next row's bulk copies, the step constants' two divisions) runs in the slack of
// the control warp; the probe of the next row's mbarrier runs in the slack of the feature warps.
// A CTA streams only ITS 256-feature slices of each sampled row: 1-D bulk copies (cp.async.bulk -> UBLKCP) into an
// 8-deep shared-memory ring driven by the sampling sequence.
// Algorithmic HBM bytes per update: 8*p (row) + 4 (index) + 8*K_y (y) + 16*K (gradient memory read + write).
#include "common.cuh"
#include "kernels.h"

namespace sgd {

size_t dense_cluster_generic_smem_bytes(int K, int p);
cudaError_t launch_saga_dense_cluster_generic(int K, int pen, size_t smem, FitDev* fit, Progress* prog, const RoundArgs& ra, cudaStream_t st);

namespace {

constexpr int kCT = 256;        // feature lanes per CTA (warps 1..8); warp 0 is the control warp
constexpr int kCBlock = kCT + 32;
constexpr int kCluster = 8;     // CTAs per fit
constexpr int kLanes = kCT * kCluster;
constexpr int kCRing = 8;       // row ring depth (power of two)
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
  return r;
}
// remote store that completes bytes on a (remote) mbarrier
__device__ __forceinline__ void st_async_f64(uint32_t addr, double v, uint32_t mbar_addr) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];" ::"r"(addr), "d"(v), "r"(mbar_addr)
               : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// named barriers used one way: producers arrive, consumers sync (both name the full block as the expected count)
__device__ __forceinline__ void bar_arrive_named(int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(kCBlock) : "memory"); }
__device__ __forceinline__ void bar_sync_named(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(kCBlock) : "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

#ifdef SGD_CL_TRACE
// Timeline trace (measurement build only): clock64 of eight events per update for updates [kClFrom, kClFrom + kClRows) of a
// launch, CTA 0 only.  control warp: 0 warp sums in hand (past barrier 1), 1 exchange stores issued, 2 all CTA sums
// arrived, 3 g_change stored.  feature warp 0: 5 row in the ring, 6 warp sums stored, 8 g_change in hand (past barrier 2),
// 9 coefficient step done.
constexpr int64_t kClFrom = 20000, kClRows = 4096;
__device__ long long g_cl_trace[kClRows][10];
__device__ __forceinline__ long long clock_ordered() {      // ordered with the barriers and memory operations around it
  long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
  return t;
}
#define CLTRACE(ev, cond) do { if ((cond) && cta == 0 && lane == 0 && tg - kClFrom >= 0 && tg - kClFrom < kClRows) g_cl_trace[tg - kClFrom][ev] = clock_ordered(); } while (0)
#else
#define CLTRACE(ev, cond)
#endif

template <int NCH>
struct __align__(128) ClusterFixed {
  double ring[kCRing][NCH][kCT];
  double part[2][kCluster + 1][32];   // CTA sums of both parities from every CTA; row kCluster: gradient memory from CTA 0
  double red[8][32];                  // warp sums, class-indexed
  double gch[32];
  double conv[2][kCluster][2];        // epoch-end maxima from every CTA
  double cred[8][2];
  double exp_tab[64];                 // sgd_exp's 2^(j/32) table (hi, lo): per-lane indices would serialise in the constant cache
  double consts[2][4];                // step constants of the update (parity): gamma / wscale, step, threshold, wscale after the step
  uint64_t full[kCRing];
  uint64_t pbar[2];                   // partial sums arrived (count 1 + transaction bytes: (kCluster + 1) * K doubles per update)
  uint64_t cbar[2];                   // epoch-end maxima arrived (count 1 + transaction bytes)
};

constexpr bool cluster_reg_state(int KT, int NCH) { return KT * NCH <= 16; }
constexpr int ilog2(int v) { return v <= 1 ? 0 : 1 + ilog2(v / 2); }

// Recursive-halving form of the 32-lane xor-butterfly over N class sums per lane (N a power of two <= 32). Returns, in
// every lane, the butterfly total of class (lane >> (5 - log2 N)).
template <int N, int O>
struct Halving {
  static __device__ __forceinline__ double run(double (&v)[N], int lane) {
    if constexpr (O == 0) {
      return v[0];
    } else if constexpr (N == 1) {
      double h[1] = {v[0] + __shfl_xor_sync(kFull, v[0], O)};
      return Halving<1, O / 2>::run(h, lane);
    } else {
      const bool up = (lane & O) != 0;
      double h[N / 2];
#pragma unroll
      for (int i = 0; i < N / 2; ++i) {
        const double keep = up ? v[N / 2 + i] : v[i];
        const double send = up ? v[i] : v[N / 2 + i];
        h[i] = keep + __shfl_xor_sync(kFull, send, O);
      }
      return Halving<N / 2, O / 2>::run(h, lane);
    }
  }
};

// sgd_log (include/sgdnet_arith.h) for arguments in [1, 64] - a sum of at most 32 exponentials of non-positive numbers
// of which one is exp(0): the same operations, hence the same bits, without the special-case exits in front of them;
// anything else (a NaN from a diverged fit) takes sgd_log itself.
static __device__ __noinline__ double sgd_log_rare(double x) { return sgd_log(x); }
__device__ __forceinline__ double sgd_log_sum(double x) {
  const uint32_t hi = static_cast<uint32_t>(__double2hiint(x)), lo = static_cast<uint32_t>(__double2loint(x));
  int32_t k = static_cast<int32_t>(hi >> 20) - 1023;
  const uint32_t mant_hi = hi & 0x000fffffu;
  // mantissa > 0x6a09e667f3bcc  <=>  m >= sqrt(2): use m/2 and k+1
  const bool big = mant_hi > 0x6a09eu || (mant_hi == 0x6a09eu && lo > 0x667f3bccu);
  k += big ? 1 : 0;
  const double m = __hiloint2double(static_cast<int>(mant_hi | (big ? 0x3fe00000u : 0x3ff00000u)), static_cast<int>(lo));
  const double f = m - 1.0;
  const double dk = static_cast<double>(k);
  const double kLn2Hi = 6.93147180369123816490e-01, kLn2Lo = 1.90821492927058770002e-10;
  const double s = f / (2.0 + f);
  const double z = s * s;
  const double w = z * z;
  const double t1 = w * (3.999999999940941908e-01 + w * (2.222219843214978396e-01 + w * 1.531383769920937332e-01));
  const double t2 = z * (6.666666666666735130e-01 +
                         w * (2.857142874366239149e-01 + w * (1.818357216161805012e-01 + w * 1.479819860511658591e-01)));
  const double R = t2 + t1;
  const double hfsq = 0.5 * f * f;
  const double out = dk * kLn2Hi - ((hfsq - (s * (hfsq + R) + dk * kLn2Lo)) - f);
  if (__builtin_expect(!(x >= 1.0 && x <= 64.0), 0)) return sgd_log_rare(x);
  return out;
}

// sgd_exp_inrange (common.cuh) reading the 2^(j/32) table from shared memory: the same operations and bits
__device__ __forceinline__ double sgd_exp_inrange_tab(double x, const double* __restrict__ tab) {
  const double kInvStep = 46.16624130844683, kStepHi = 0.021660849392446835, kStepLo = 5.145609244655338e-14;
  const double kShift = 6755399441055744.0;
  const double ts = fma(x, kInvStep, kShift);
  const double kd = ts - kShift;
  const int32_t k = __double2loint(ts);
  double r = fma(-kd, kStepHi, x);
  r = fma(-kd, kStepLo, r);
  double p = 1.0 / 720.0;
  p = fma(p, r, 1.0 / 120.0);
  p = fma(p, r, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  const double q = fma(r * r, p, r);
  const int32_t j = k & 31, m = k >> 5;
  const double2 t = *reinterpret_cast<const double2*>(tab + 2 * j);
  const double res = t.x + fma(t.x, q, t.y);
  const double out = res * __hiloint2double((m + 1023) << 20, 0);
  if (__builtin_expect(!(x >= -707.0 && x <= 709.0), 0)) return sgd_exp_rare(x);
  return out;
}

// LogSumExp (src/math.h:25-33) over the classes held one per lane (lanes that hold no class pass valid = false), the
// bits of lse_warp (common.cuh), for at most KT classes in lanes 0 .. KT-1. The maximum is exact in any order: two
// integer warp reductions over an order-preserving key replace five shuffle-and-compare rounds. The sum is the
// 32-slot butterfly padded with zeros (include/sgdnet_arith.h, item 2): the levels whose partner lanes all hold the
// pad (offsets >= KT) add +0.0 to a non-negative number and are skipped.
template <int KT>
__device__ __forceinline__ double lse_classes(double lp, bool valid, int lane, const double* __restrict__ exp_tab) {
  const long long b = __double_as_longlong(valid ? lp : -INFINITY);
  const unsigned long long key = b < 0 ? ~static_cast<unsigned long long>(b) : (static_cast<unsigned long long>(b) | 0x8000000000000000ull);
  const unsigned hi = static_cast<unsigned>(key >> 32), lo = static_cast<unsigned>(key);
  const unsigned hmax = __reduce_max_sync(kFull, hi);
  const unsigned lmax = __reduce_max_sync(kFull, hi == hmax ? lo : 0u);
  const unsigned long long kmax = (static_cast<unsigned long long>(hmax) << 32) | lmax;
  const double mx = __longlong_as_double((kmax >> 63) ? static_cast<long long>(kmax & 0x7fffffffffffffffull) : static_cast<long long>(~kmax));
  double e = valid ? sgd_exp_inrange_tab(lp - mx, exp_tab) : 0.0;
#pragma unroll
  for (int o = (KT >= 32 ? 16 : KT / 2); o > 0; o >>= 1) e += __shfl_xor_sync(kFull, e, o);
  const double sum = (lane < KT) ? e : 1.0;      // lanes beyond the first group hold no class and no sum
  return sgd_log_sum(sum) + mx;
}

// div_by_n (common.cuh) with the way out deferred: instead of calling the IEEE division for an operand outside the safe
// exponent range, it raises `rare`, and the caller redoes its quotients with the division once, after the straight-line
// code. Keeps a sweep over several classes free of branches, so the classes' dependent chains overlap.
__device__ __forceinline__ double div_by_n_flag(double a, double nd, double rn, bool& rare) {
  const uint32_t hi = static_cast<uint32_t>(__double2hiint(a));
  const uint32_t ex = (hi >> 20) & 0x7ffu;
  const bool ok = (ex - 127u) <= 1792u;
  const double q0 = a * rn;
  const double r0 = fma(-q0, nd, a);
  const double q1 = fma(r0, rn, q0);
  const bool zero = ((hi << 1) | static_cast<uint32_t>(__double2loint(a))) == 0u;
  rare |= !ok && !zero;
  return ok ? q1 : a;       // a == +-0 when !ok && zero
}

}  // namespace

#ifdef SGD_CL_TRACE
}  // namespace sgd
extern "C" void sgdnet_debug_cluster_trace(long long* out) { cudaMemcpyFromSymbol(out, sgd::g_cl_trace, sizeof(sgd::g_cl_trace)); }
namespace sgd {
#endif

namespace {

// KT: class-count bucket (1, 4, 8, 16, 32); PEN: penalty functor; NCH: 256-feature slices per CTA and row (1, 2, 4);
// CS: class split (1, 2, 4) - a design with at most 8 / CS feature blocks (p <= 2048 / CS) leaves CTAs of the cluster
// without features, so the classes of a feature block are split over CS CTAs instead: CTA c owns feature block c % nfb
// and the KT / CS classes from (c / nfb) * KT / CS on, which divides the feature warps' work per update (dot products,
// butterfly, coefficient step) by CS. Per class the dot product's association does not change: its 256-lane block sums
// still come one from each feature block. Not for the group lasso, whose prox couples the classes of a feature.
template <int KT, int PEN, int NCH, int CS>
__global__ void __launch_bounds__(kCBlock, 1)
saga_dense_cluster_kernel(FitDev* __restrict__ fit, Progress* __restrict__ prog, const RoundArgs ra) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int KL = KT / CS;                              // classes owned by this CTA's feature lanes
  static_assert(CS == 1 || (NCH == 1 && PEN != kGroupLasso && KL >= 1), "class split: one slice, separable prox");
  constexpr bool kReg = cluster_reg_state(KL, NCH);
  constexpr int kNF = NCH * kCT;
  Progress& pg = *prog;
  const FitDev& f = *fit;
  const uint32_t cta = cluster_ctarank();
  const uint32_t nct = cluster_nctarank();     // 2, 4 or 8 CTAs: feature blocks of the design (at most 8) x class split
  const uint32_t nfb = nct / CS;               // 256-feature blocks
  const uint32_t fb = cta % nfb;               // this CTA's feature block ...
  const int kbase = static_cast<int>(cta / nfb) * KL;      // ... and first class
  if (ra.n_epochs <= 0 || pg.status != kRunning) {       // uniform over the cluster
    if (cta == 0 && threadIdx.x == 0) {
      pg.epochs_last_launch = 0;
      publish_progress(f.mirror, pg, ra.round_id);
    }
    return;
  }
  const bool free_run = (ra.flags & 1) != 0;
  const uint64_t t_start = globaltimer_ns();

  ClusterFixed<NCH>& sm = *reinterpret_cast<ClusterFixed<NCH>*>(smem_raw);
  double* const Ws = reinterpret_cast<double*>(smem_raw + sizeof(ClusterFixed<NCH>));   // [KL][NCH * 256] when !kReg
  double* const Gs = Ws + (kReg ? 0 : KL * kNF);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool control = warp == 0;
  const int tid = static_cast<int>(threadIdx.x) - 32;      // feature lane of this CTA (negative in the control warp)
  const int fwarp = warp - 1;
  const int K = f.K, p = f.p, ld = f.ld, Ky = f.Ky;
  const int64_t n = f.n;
  const double nd = static_cast<double>(static_cast<uint32_t>(n));
  const double rn = 1.0 / nd;
  const int family = f.family;
  const bool fit_intercept = f.fit_intercept != 0;

  const int li = pg.lambda_ind;
  const double gamma = f.gamma[li], alpha = f.alpha[li], beta = f.beta[li];
  const double r = 1.0 - alpha * gamma;     // wscale_update
  const double bgs = beta * gamma * 1.0;

  double* const Wg = f.W;
  double* const Gg = f.gsum;
  const uint32_t* __restrict__ seq = ra.seq;
  const int64_t total = n * ra.n_epochs;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kCRing; ++i) mbar_init(&sm.full[i], 1);
    mbar_init(&sm.pbar[0], 1);
    mbar_init(&sm.pbar[1], 1);
    mbar_init(&sm.cbar[0], 1);
    mbar_init(&sm.cbar[1], 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 2 * (kCluster + 1) * 32; i += kCBlock) (&sm.part[0][0][0])[i] = 0.0;
  for (int i = threadIdx.x; i < 8 * 32; i += kCBlock) (&sm.red[0][0])[i] = 0.0;
  for (int i = threadIdx.x; i < 64; i += kCBlock) sm.exp_tab[i] = sgd_exp_tab_dev[i];
  if (threadIdx.x == 0) {       // step constants of the first update (wscale == 1)
    sm.consts[0][0] = gamma / (1.0 * r);
    sm.consts[0][1] = bgs / (1.0 * r);
  }

  // feature slot (i, tid)  <->  feature j = 2048 * i + 256 * cta + tid
  const int jbase = kCT * static_cast<int>(fb) + tid;
  double Wr[kReg ? NCH * KL : 1], Gr[kReg ? NCH * KL : 1];
  if (!control) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int j = kLanes * i + jbase;
#pragma unroll
      for (int k = 0; k < KL; ++k) {
        const bool have = j < p && kbase + k < K;
        const double w0 = have ? Wg[size_t(kbase + k) * p + j] : 0.0;
        const double g0 = have ? Gg[size_t(kbase + k) * p + j] : 0.0;
        if constexpr (kReg) {
          Wr[i * KL + k] = w0;
          Gr[i * KL + k] = g0;
        } else {
          Ws[(k * NCH + i) * kCT + tid] = w0;
          Gs[(k * NCH + i) * kCT + tid] = g0;
        }
      }
    }
  }
  __syncthreads();
  cluster_sync_all();       // every CTA's barriers exist before anybody arrives on them remotely

  double wscale = 1.0;
  uint32_t it_outer = pg.it_outer;
  uint32_t epochs_done = 0;
  int64_t tg = 0;
  bool finished = false;

  // ------------------------------------------------------------------ control warp state
  // this CTA's slices of a row: slice i covers features [2048 i + 256 cta, +256) clipped to the row's padded length
  uint32_t slice_b[NCH];
  uint32_t row_bytes = 0;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int j0 = kLanes * i + kCT * static_cast<int>(fb);
    const int len = (j0 >= ld) ? 0 : ((ld - j0 < kCT) ? ld - j0 : kCT);
    slice_b[i] = static_cast<uint32_t>(len) * 8u;
    row_bytes += slice_b[i];
  }
  auto issue_row = [&](int64_t q, uint32_t sq) {      // one thread; sq = seq[q]
    const int slot = static_cast<int>(q & (kCRing - 1));
    const double* src = f.xd + size_t(sq) * ld + kCT * fb;
    if (row_bytes == 0) {
      mbar_arrive(&sm.full[slot]);
      return;
    }
    mbar_expect_tx(&sm.full[slot], row_bytes);
#pragma unroll
    for (int i = 0; i < NCH; ++i)
      if (slice_b[i]) bulk_g2s(&sm.ring[slot][i][0], src + kLanes * i, slice_b[i], &sm.full[slot]);
  };
  const bool issuer = control && lane == 31;
  uint32_t s_refill = 0;
  if (issuer) {
    for (int64_t q = 0; q < kCRing && q < total; ++q) issue_row(q, seq[q]);
    if (kCRing < total) s_refill = seq[kCRing];
  }
  // intercept state: class k in lane k of the control warp of EVERY CTA (identical, redundant)
  const bool valid = control && lane < K;
  double b_reg = 0.0, gsi_reg = 0.0;
  // per-sample operands one update ahead. The gradient memory is read and written by CTA 0 ALONE, by the same lanes and
  // with the store of update t ahead of the load for update t + 1 in program order, which is all the coherence it needs;
  // it travels to the other CTAs with CTA 0's partial sums.
  auto fetch_y = [&](uint32_t sx) { return f.yt[size_t(sx) * Ky + (Ky == 1 ? 0 : lane)]; };
  uint32_t s_cur = 0, s_n1 = 0, s_n2 = 0;
  double y_cur = 0.0, gm_cur = 0.0;
  // exchange role of a control lane: local class xk to destination CTA xd (+ 32 / KL per round); CTA 0 sends the
  // gradient memory of all classes the same way (class gk, destination gd + 32 / KT per round)
  const uint32_t xk = static_cast<uint32_t>(lane) % KL, xd = static_cast<uint32_t>(lane) / KL;
  const uint32_t gk = static_cast<uint32_t>(lane) % KT, gd = static_cast<uint32_t>(lane) / KT;
  uint32_t a_gm = 0;
  uint32_t a_mine = 0, a_pbar = 0;      // slot of class xk in part[0][cta][] and pbar[0], as shared-window addresses
  if (control) {
    s_cur = seq[0];
    s_n1 = (total > 1) ? seq[1] : 0u;
    s_n2 = (total > 2) ? seq[2] : 0u;
    if (valid) {
      b_reg = f.b[lane];
      gsi_reg = f.gsi[lane];
      y_cur = fetch_y(s_cur);
      if (cta == 0) gm_cur = f.gmem[size_t(s_cur) * K + lane];
    }
    a_mine = smem_u32(&sm.part[0][fb][kbase + xk]);
    a_gm = smem_u32(&sm.part[0][kCluster][gk]);
    a_pbar = smem_u32(&sm.pbar[0]);
  }
  constexpr uint32_t kParStride = (kCluster + 1) * 32 * 8;                 // bytes between the two parities of part[]

  for (int ep = 0; ep < ra.n_epochs && !finished; ++ep) {
    if (control) {
      // ================================================================ control warp
      for (int64_t t = 0; t < n; ++t, ++tg) {
        const uint32_t xpar = static_cast<uint32_t>(tg & 1);                // exchange buffer of this update
        const uint32_t xphase = static_cast<uint32_t>((tg >> 1) & 1);       // phase of pbar[xpar]
        bar_sync_named(1);                                                  // warp sums are in red[]
        CLTRACE(0, true);
        double tsum = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) tsum += sm.red[w][lane];
        // this CTA's sums (its classes, its feature block) go to all nct CTAs: (class, destination) pairs are spread
        // over the 32 lanes, 32 / KL destinations per round, so the control warp issues one or two remote stores per
        // lane instead of a chain of them
        {
          constexpr uint32_t kDestPerRound = 32 / KL;
          const uint32_t a_sum = a_mine + xpar * kParStride, a_bar = a_pbar + xpar * 8u;
          const double v_sum = __shfl_sync(kFull, tsum, xk);
#pragma unroll
          for (uint32_t c0 = 0; c0 < kCluster; c0 += kDestPerRound) {
            const uint32_t c = c0 + xd;
            if (c0 < nct && c < nct && kbase + static_cast<int>(xk) < K) st_async_f64(map_to_cta(a_sum, c), v_sum, map_to_cta(a_bar, c));
          }
          if (cta == 0) {
            constexpr uint32_t kGmPerRound = 32 / KT;
            const double v_gm = __shfl_sync(kFull, gm_cur, gk);
#pragma unroll
            for (uint32_t c0 = 0; c0 < kCluster; c0 += kGmPerRound) {
              const uint32_t c = c0 + gd;
              if (c0 < nct && c < nct && static_cast<int>(gk) < K)
                st_async_f64(map_to_cta(a_gm + xpar * kParStride, c), v_gm, map_to_cta(a_bar, c));
            }
          }
        }
        // this CTA expects (nfb + 1) * K doubles per update on its own barrier (per class one sum from every feature block,
        // and the gradient memory): one local arrival arms the phase
        if (lane == 0) mbar_expect_tx(&sm.pbar[xpar], (nfb + 1u) * static_cast<uint32_t>(K) * 8u);
        CLTRACE(1, true);
        mbar_wait(&sm.pbar[xpar], xphase);
        CLTRACE(2, true);
        double dot = 0.0;
#pragma unroll
        for (int c = 0; c < kCluster; ++c)
          if (c < static_cast<int>(nfb)) dot += sm.part[xpar][c][lane];       // blocks beyond the design's features are exact zeros
        const double gm_val = sm.part[xpar][kCluster][lane];
        const double lp = dot * wscale + b_reg;
        double g;
        if (family == kMultinomial) {
          const double yc = __shfl_sync(kFull, y_cur, 0);
          const double lse = lse_classes<KT>(lp, valid, lane, sm.exp_tab);
          g = sgd_exp_inrange_tab(lp - lse, sm.exp_tab);
          if (static_cast<unsigned>(lane) == static_cast<unsigned>(yc + 0.5)) g -= 1.0;
        } else if (family == kBinomial) {
          g = 1.0 - y_cur - 1.0 / (1.0 + sgd_exp_inrange_tab(lp, sm.exp_tab));
        } else {
          g = lp - y_cur;
        }
        const double gch = g - gm_val;
        sm.gch[lane] = valid ? gch : 0.0;
        CLTRACE(3, true);
        bar_arrive_named(2);                                                // g_change is in gch[]
        // ---- slack: the feature warps are in their coefficient step and next dot product
        if (valid) {
          if (cta == 0) f.gmem[size_t(s_cur) * K + lane] = g;
          if (fit_intercept) {
            const double gn = div_by_n(gch, nd, rn);
            gsi_reg += gn;
            b_reg -= gamma * (gsi_reg + gn);
          }
        }
        if (wscale < kSmall) wscale = 1.0;
        wscale *= r;
        {      // step constants of the next update (the coefficient scale is back at 1 when an epoch begins)
          const double w_next = (t + 1 == n) ? 1.0 : wscale;
          const double ws_n = ((w_next < kSmall) ? 1.0 : w_next) * r;
          if (lane == 0) {
            sm.consts[(tg + 1) & 1][0] = gamma / ws_n;
            sm.consts[(tg + 1) & 1][1] = bgs / ws_n;
          }
        }
        if (issuer && tg + kCRing < total) {      // the slot this update's row sat in is free: every lane read it before barrier 1
          issue_row(tg + kCRing, s_refill);
          if (tg + kCRing + 1 < total) s_refill = seq[tg + kCRing + 1];
        }
        s_cur = s_n1;
        s_n1 = s_n2;
        if (tg + 3 < total) s_n2 = seq[tg + 3];
        if (valid && tg + 1 < total) {
          y_cur = fetch_y(s_cur);
          if (cta == 0) {
            gm_cur = f.gmem[size_t(s_cur) * K + lane];
            prefetch_l2(&f.gmem[size_t(s_n1) * K + lane]);
          }
          prefetch_l2(&f.yt[size_t(s_n1) * Ky]);
        }
      }
    } else {
      // ================================================================ feature warps
      for (int64_t t = 0; t < n; ++t, ++tg) {
        const int slot = static_cast<int>(tg & (kCRing - 1));
        CLTRACE(4, fwarp == 0);
        if (tg == 0) mbar_wait(&sm.full[0], 0u);      // later rows are waited for in the shadow of the previous exchange
        CLTRACE(5, fwarp == 0);
        // ---- A: partial dot products of this lane (ascending j), warp butterfly by recursive halving, warp sums
        double x[NCH];
#pragma unroll
        for (int i = 0; i < NCH; ++i) x[i] = (kLanes * i + jbase < p) ? sm.ring[slot][i][tid] : 0.0;
        double acc[KL];
#pragma unroll
        for (int k = 0; k < KL; ++k) acc[k] = 0.0;
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
#pragma unroll
          for (int k = 0; k < KL; ++k) {
            const double w = kReg ? Wr[i * KL + k] : Ws[(k * NCH + i) * kCT + tid];
            acc[k] += w * x[i];
          }
        }
        const double tot = Halving<KL, 16>::run(acc, lane);
        constexpr int kShift = 5 - ilog2(KL);
        if ((lane & ((1 << kShift) - 1)) == 0) sm.red[fwarp][lane >> kShift] = tot;
        CLTRACE(6, fwarp == 0);
        bar_arrive_named(1);
        // this update's step constants are functions of the deterministic wscale track: the two divisions (gamma / wscale,
        // (beta gamma) / wscale with wscale as it will be after this step) were done by the control warp in its slack
        const bool reset = wscale < kSmall;
        const double ws_c = (reset ? 1.0 : wscale) * r;
        // the next update's row (its copy was issued kCRing - 1 updates ago): a completed mbarrier still costs a probe
        // of about 90 cycles, spent here while the control warp exchanges instead of at the head of the next update
        if (tg + 1 < total) mbar_wait(&sm.full[(tg + 1) & (kCRing - 1)], static_cast<uint32_t>(((tg + 1) / kCRing) & 1));
        bar_sync_named(2);                                                  // g_change is in gch[]
        CLTRACE(8, fwarp == 0);
        const double gw = sm.consts[tg & 1][0];
        const double step = gw * 1.0;
        const double thr = sm.consts[tg & 1][1];
        double gch[KL];
#pragma unroll
        for (int k = 0; k < KL; ++k) gch[k] = sm.gch[kbase + k];
        if (reset) {      // src/saga-dense.h:166-170
#pragma unroll
          for (int i = 0; i < NCH; ++i)
#pragma unroll
            for (int k = 0; k < KL; ++k) {
              if constexpr (kReg) Wr[i * KL + k] *= wscale;
              else Ws[(k * NCH + i) * kCT + tid] *= wscale;
            }
        }
        wscale = ws_c;
        // ---- C: fused coefficient step, prox, gradient-average update on the owned features
        // (src/saga-dense.h:176-183; penalty functors src/penalties.h:27-79 with scaling = 1). Slots beyond p and classes
        // beyond K hold zeros and stay zeros under these operations.
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
          double w[KL], gs[KL], gnew[KL];
          double sq = 0.0;
          bool rare = false;
#pragma unroll
          for (int k = 0; k < KL; ++k) {
            const double w_in = kReg ? Wr[i * KL + k] : Ws[(k * NCH + i) * kCT + tid];
            gs[k] = kReg ? Gr[i * KL + k] : Gs[(k * NCH + i) * kCT + tid];
            const double gx = gch[k] * x[i];
            const double v = (w_in - gx * gw) - step * gs[k];
            w[k] = (PEN == kElasticNet) ? soft_threshold(v, thr) : v;
            if (PEN == kGroupLasso) sq += v * v;
            gnew[k] = gs[k] + div_by_n_flag(gx, nd, rn, rare);
          }
          if (__builtin_expect(rare, 0)) {
#pragma unroll
            for (int k = 0; k < KL; ++k) gnew[k] = gs[k] + (gch[k] * x[i]) / nd;
          }
          if (PEN == kGroupLasso) {
            const double factor = bgs / sqrt(sq);
            const double mult = 1.0 - factor / ws_c;
#pragma unroll
            for (int k = 0; k < KL; ++k) w[k] = (factor < 1.0) ? w[k] * mult : 0.0;
          }
#pragma unroll
          for (int k = 0; k < KL; ++k) {
            if constexpr (kReg) {
              Wr[i * KL + k] = w[k];
              Gr[i * KL + k] = gnew[k];
            } else {
              Ws[(k * NCH + i) * kCT + tid] = w[k];
              Gs[(k * NCH + i) * kCT + tid] = gnew[k];
            }
          }
        }
        CLTRACE(9, fwarp == 0);
      }
    }

    // ---- epoch end: unscale, convergence over the whole cluster (src/saga-dense.h:188-208, src/utils.h:240-262)
    double mc = 0.0, ms = 0.0;
    if (!control) {
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int j = kLanes * i + jbase;
#pragma unroll
        for (int k = 0; k < KL; ++k) {
          if (j < p && kbase + k < K) {
            const size_t e = size_t(kbase + k) * p + j;
            const double w_in = kReg ? Wr[i * KL + k] : Ws[(k * NCH + i) * kCT + tid];
            const double w = w_in * wscale;
            if constexpr (kReg) Wr[i * KL + k] = w;
            else Ws[(k * NCH + i) * kCT + tid] = w;
            mc = fmax(mc, fabs(w - f.Wprev[e]));
            ms = fmax(ms, fabs(w));
            f.Wprev[e] = w;
          }
        }
      }
      mc = warp_max(mc);
      ms = warp_max(ms);
      if (lane == 0) {
        sm.cred[fwarp][0] = mc;
        sm.cred[fwarp][1] = ms;
      }
    }
    wscale = 1.0;
    __syncthreads();
    const uint32_t epar = static_cast<uint32_t>(ep & 1), ephase = static_cast<uint32_t>((ep >> 1) & 1);
    if (threadIdx.x == 0) {
      double mc_c = 0.0, ms_c = 0.0;
      for (int w = 0; w < 8; ++w) {
        mc_c = fmax(mc_c, sm.cred[w][0]);
        ms_c = fmax(ms_c, sm.cred[w][1]);
      }
      const uint32_t a_conv = smem_u32(&sm.conv[epar][cta][0]);
      const uint32_t a_cbar = smem_u32(&sm.cbar[epar]);
      for (uint32_t c = 0; c < nct; ++c) {
        const uint32_t bar_c = map_to_cta(a_cbar, c);
        st_async_f64(map_to_cta(a_conv, c), mc_c, bar_c);
        st_async_f64(map_to_cta(a_conv + 8u, c), ms_c, bar_c);
      }
      mbar_expect_tx(&sm.cbar[epar], nct * 16u);
    }
    mbar_wait(&sm.cbar[epar], ephase);
    double mc_all = 0.0, ms_all = 0.0;
    for (int c = 0; c < static_cast<int>(nct); ++c) {
      mc_all = fmax(mc_all, sm.conv[epar][c][0]);
      ms_all = fmax(ms_all, sm.conv[epar][c][1]);
    }
    const bool all_zero = (ms_all == 0.0) && (mc_all == 0.0);
    const bool no_change = (ms_all != 0.0) && (mc_all / ms_all <= f.tol);
    ++it_outer;
    ++epochs_done;
    finished = !free_run && ((all_zero || no_change) || !(it_outer < f.max_iter));
  }

  // drain copies that were issued but never consumed (early stop) before the shared memory goes away
  if (issuer)
    for (int64_t q = tg; q < tg + kCRing && q < total; ++q)
      mbar_wait(&sm.full[static_cast<int>(q & (kCRing - 1))], static_cast<uint32_t>((q / kCRing) & 1));
  __syncthreads();

  if (!control) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int j = kLanes * i + jbase;
#pragma unroll
      for (int k = 0; k < KL; ++k) {
        if (j < p && kbase + k < K) {
          Wg[size_t(kbase + k) * p + j] = kReg ? Wr[i * KL + k] : Ws[(k * NCH + i) * kCT + tid];
          Gg[size_t(kbase + k) * p + j] = kReg ? Gr[i * KL + k] : Gs[(k * NCH + i) * kCT + tid];
        }
      }
    }
  }
  if (cta == 0 && valid) {
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

template <int KT, int NCH, int CS>
constexpr size_t cluster_smem(void) {
  return sizeof(ClusterFixed<NCH>) + (cluster_reg_state(KT / CS, NCH) ? 0 : sizeof(double) * 2 * size_t(KT / CS) * NCH * kCT);
}

inline int nch_bucket(int p) {
  const int nch = (p + kLanes - 1) / kLanes;
  return nch <= 1 ? 1 : (nch <= 2 ? 2 : (nch <= 4 ? 4 : 0));
}
// class split of a shape (see the kernel): only with one slice, a separable prox and at least CS classes in the bucket
inline int class_split(int kt, int p, int pen) {
  if (p > 4 * kCT || pen == kGroupLasso) return 1;
  const int cs = p <= 2 * kCT ? 4 : 2;
  return kt >= cs ? cs : (kt >= 2 ? 2 : 1);
}
inline size_t fast_smem_bytes(int kt, int nch, int cs) {
  const size_t fixed = nch == 1 ? sizeof(ClusterFixed<1>) : (nch == 2 ? sizeof(ClusterFixed<2>) : sizeof(ClusterFixed<4>));
  return fixed + (cluster_reg_state(kt / cs, nch) ? 0 : sizeof(double) * 2 * size_t(kt / cs) * nch * kCT);
}
// the compile-time instantiations cover slice counts 1, 2, 4 whose state fits the registers or the shared memory
inline bool fast_eligible(int K, int p, int pen) {
  const int nch = nch_bucket(p);
  return nch != 0 && fast_smem_bytes(dense_kt_bucket(K), nch, class_split(dense_kt_bucket(K), p, pen)) <= dense_smem_budget();
}

template <int KT, int PEN, int NCH, int CS>
cudaError_t launch_cluster_variant(int nct, FitDev* fit, Progress* prog, const RoundArgs& ra, cudaStream_t st) {
  constexpr size_t smem = cluster_smem<KT, NCH, CS>();
  if constexpr (smem > 227 * 1024) {
    return cudaErrorInvalidConfiguration;       // not reachable: fast_eligible() sends these shapes to the generic kernel
  } else {
    cudaError_t e = cudaFuncSetAttribute(saga_dense_cluster_kernel<KT, PEN, NCH, CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(nct, 1, 1);
    cfg.blockDim = dim3(kCBlock, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = nct;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, saga_dense_cluster_kernel<KT, PEN, NCH, CS>, fit, prog, ra);
  }
}

template <int KT, int PEN>
cudaError_t launch_cluster_nch(int nch, int cs, int nct, FitDev* fit, Progress* prog, const RoundArgs& ra, cudaStream_t st) {
  if constexpr (PEN != kGroupLasso && KT >= 2) {
    if (nch == 1 && cs == 2) return launch_cluster_variant<KT, PEN, 1, 2>(nct, fit, prog, ra, st);
    if constexpr (KT >= 4) {
      if (nch == 1 && cs == 4) return launch_cluster_variant<KT, PEN, 1, 4>(nct, fit, prog, ra, st);
    }
  }
  switch (nch) {
    case 1: return launch_cluster_variant<KT, PEN, 1, 1>(nct, fit, prog, ra, st);
    case 2: return launch_cluster_variant<KT, PEN, 2, 1>(nct, fit, prog, ra, st);
    default: return launch_cluster_variant<KT, PEN, 4, 1>(nct, fit, prog, ra, st);
  }
}

template <int KT>
cudaError_t launch_cluster_kt(int pen, int nch, int cs, int nct, FitDev* fit, Progress* prog, const RoundArgs& ra, cudaStream_t st) {
  switch (pen) {
    case kRidge: return launch_cluster_nch<KT, kRidge>(nch, cs, nct, fit, prog, ra, st);
    case kElasticNet: return launch_cluster_nch<KT, kElasticNet>(nch, cs, nct, fit, prog, ra, st);
    default: return launch_cluster_nch<KT, kGroupLasso>(nch, cs, nct, fit, prog, ra, st);
  }
}

}  // namespace

size_t dense_cluster_smem_bytes(int K, int p, int pen) {
  const int kt = dense_kt_bucket(K);
  return fast_eligible(K, p, pen) ? fast_smem_bytes(kt, nch_bucket(p), class_split(kt, p, pen)) : dense_cluster_generic_smem_bytes(K, p);
}

cudaError_t launch_saga_dense_cluster(int K, int p, int pen, size_t smem, FitDev* fit, Progress* prog, const RoundArgs& ra, cudaStream_t st) {
  if (!fast_eligible(K, p, pen)) return launch_saga_dense_cluster_generic(K, pen, smem, fit, prog, ra, st);
  const int nch = nch_bucket(p);
  const int cs = class_split(dense_kt_bucket(K), p, pen);
  // CTAs: one per 256-feature block of the design (absent blocks are exact zeros) x class split
  const int nct = (p <= 2 * kCT ? 2 : (p <= 4 * kCT ? 4 : kCluster)) * cs;
  switch (dense_kt_bucket(K)) {
    case 1: return launch_cluster_kt<1>(pen, nch, cs, nct, fit, prog, ra, st);
    case 4: return launch_cluster_kt<4>(pen, nch, cs, nct, fit, prog, ra, st);
    case 8: return launch_cluster_kt<8>(pen, nch, cs, nct, fit, prog, ra, st);
    case 16: return launch_cluster_kt<16>(pen, nch, cs, nct, fit, prog, ra, st);
    default: return launch_cluster_kt<32>(pen, nch, cs, nct, fit, prog, ra, st);
  }
}

}  // namespace sgd
