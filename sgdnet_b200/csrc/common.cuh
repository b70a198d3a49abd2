// common.cuh — device-side vocabulary shared by the sgdnet_b200 kernels (sm_100a, FP64 throughout).
//
// The arithmetic below restates, operation for operation, the reference's functors so that the per-element
// order of floating point operations is the reference's:
//   SoftThreshold            src/prox.h:32-39
//   Ridge/ElasticNet/Group   src/penalties.h:27-79
//   Gradient / Loss          src/families.h:81-96, 152-168, 235-260, 350-365
//   LogSumExp                src/math.h:25-33
// The library is compiled with -fmad=false: no multiply-add contraction anywhere. Explicit fma()
// calls (used only where they are provably exact) are unaffected by that flag.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sgdnet_arith.h"

namespace sgd {

constexpr double kSmall = 100.0 * 2.220446049250313e-16;   // src/constants.h:22

enum Family : int { kGaussian = 0, kBinomial = 1, kMultinomial = 2, kMGaussian = 3 };
enum Penalty : int { kRidge = 0, kElasticNet = 1, kGroupLasso = 2 };
enum Status : int { kRunning = 0, kLambdaDone = 1, kFitDone = 2 };

constexpr int kMaxClasses = 64;   // K handled by one warp in the gradient step
constexpr int kRescaleBlocks = 64; // CTAs of the per-lambda rescale + archive pass

// One padded CSR row: `start` is a multiple of 4 entries so that the index run (int32) and the value run (double)
// both begin on 16-byte boundaries, which is what cp.async.bulk needs. Pad entries are never read as data.
struct __align__(16) RowInfo {
  int64_t start;
  int32_t nnz;
  int32_t pad_;
};

// Sparse K == 1 coefficient state, one 32-byte record per feature so that a row's gather / scatter is a single
// 256-bit access per nonzero (saga_sparse.cu, wavefront kernel).
struct __align__(32) FeatState {
  double w, g;
  uint32_t lag, pad0_;
  uint64_t pad1_;
};

// Everything one fit needs on the device. One of these per fit lives in HBM; CTA `blockIdx.x` (or blockIdx.y for
// the grid-wide passes) works on fit `blockIdx.x`. The host mirrors the struct and re-reads only `Progress`.
struct Progress;
struct FitDev {
  // ---- problem shape
  int32_t sparse, family, penalty, fit_intercept;
  int32_t standardize;       // sparse + standardize: virtual centring through `c` (src/saga-sparse.h:127-128, 276-277)
  int32_t K, Ky, p, ld;      // ld = dense row stride in doubles (p rounded up to 2)
  int64_t n;
  // ---- design (read-only, possibly shared by several fits)
  const double*  xd;         // dense: [n][ld] row-major, standardised
  const RowInfo* rows;       // sparse: padded CSR
  const int32_t* ci;
  const double*  cv;
  const double*  c;          // x_center_scaled [p] (zeros unless sparse&&standardize)
  const double*  yt;         // [n][Ky]
  // ---- warm-start state (src/sgdnet.cpp:186-198), class-major: W[k*p + j]
  double *W, *gsum, *Wprev, *b, *gsi, *gmem;   // gmem [n][K]
  uint32_t* lag;             // [p]
  FeatState* st;             // [p] sparse K == 1: packed {W, gsum, lag} (then W mirrors st.w at epoch ends only)
  double*   lag_scaling;     // [n+1] (unused when ls_identity)
  // ---- path
  const double *gamma, *alpha, *beta;   // per lambda
  int32_t  n_lambda;
  uint32_t max_iter;
  double   tol;
  double   null_deviance_scaled;
  // ---- rescale inputs and archives (src/utils.h:352-378)
  const double *x_center, *x_scale, *y_center, *y_scale;
  double *beta_arch;         // [n_lambda][p][K]
  double *a0_arch;           // [n_lambda][K]
  double *dev_ratio;         // [n_lambda]
  uint32_t *epochs, *codes;  // [n_lambda]
  double *losses;            // debug: [n_lambda * max_iter] or null
  // ---- scratch
  double *partials;          // per-block partial sums for the grid-wide passes
  double *xb_partials;       // [kRescaleBlocks][K] per-block sums of x_center_j * beta_j (Rescale)
  uint32_t *nz_mask;         // [ceil(p/32)] sparse K == 1: bit j set iff W[j] != 0 at the last rescale (deviance pass)
  int32_t debug;
  int32_t pad0_;             // always 0 (read as a run-time zero, see dep_on)
  struct Progress* mirror;   // pinned host copy of the fit's Progress, written by the kernels that change it
};

// Device-updated progress of one fit; read back by the host after every round.
struct Progress {
  int32_t  lambda_ind;
  int32_t  status;           // Status
  uint32_t it_outer;         // epochs run so far at lambda_ind
  uint32_t npasses;          // accumulated over the path
  uint32_t epochs_last_launch;
  uint32_t round_seen;       // id of the last round (RoundArgs::round_id) whose result this is; the host polls it
  double   wscale;           // carried only inside an epoch; 1.0 between epochs
  uint64_t solver_ns;        // device time spent in the SAGA epoch kernels so far (%globaltimer around each launch)
};

// The host does not synchronise streams to learn that a round is over: the one thread that updates a fit's Progress
// copies it to pinned host memory, payload first, then (system-scope fence in between) the round id the host is
// polling for. Stream order still governs everything on the device.
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void publish_progress(Progress* mirror, const Progress& pg, uint32_t round_id) {
  if (mirror == nullptr) return;
  volatile Progress* m = mirror;
  m->lambda_ind = pg.lambda_ind;
  m->status = pg.status;
  m->it_outer = pg.it_outer;
  m->npasses = pg.npasses;
  m->epochs_last_launch = pg.epochs_last_launch;
  m->wscale = pg.wscale;
  m->solver_ns = pg.solver_ns;
  __threadfence_system();
  m->round_seen = round_id;
}

// ------------------------------------------------------------------------------------------ small helpers
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// SoftThreshold (src/prox.h:32-39): max(x - s, 0) - max(-x - s, 0) for a threshold s >= 0. Evaluated as
// t = |x| - s; t <= 0 ? +0 : copysign(t, x) - the same value bit for bit (IEEE rounding is symmetric, so
// RN(|x| - s) = |RN(x -+ s)|; the side that is clamped contributes an exact +0; a NaN stays a NaN as with
// std::max), in a third of the instructions (an FP64 max is a compare-and-select sequence).
// tests/test_arith_cpu.py checks the identity against the reference form.
__device__ __forceinline__ double soft_threshold(double x, double s) {
  const double t = fabs(x) - s;
  return (t <= 0.0) ? 0.0 : copysign(t, x);
}

// The three scalars a penalty functor derives from (gamma, beta, w_scale, scaling); computed once per call site with
// the reference's operation order: step = gamma/w_scale*scaling, thr = beta*gamma*scaling/w_scale (ElasticNet),
// bgs = beta*gamma*scaling (GroupLasso numerator).
struct PenCoef {
  double step, thr, bgs, w_scale;
};
__device__ __forceinline__ PenCoef pen_coef(double gamma, double beta, double w_scale, double scaling) {
  PenCoef c;
  c.step = gamma / w_scale * scaling;
  c.bgs = beta * gamma * scaling;
  c.thr = c.bgs / w_scale;
  c.w_scale = w_scale;
  return c;
}

// penalty(w, j, w_scale, scaling, g_sum) on the K values of one feature, strided by `stride` doubles.
template <typename WPtr, typename GPtr>
__device__ __forceinline__ void apply_penalty(int pen, WPtr w, GPtr gs, int K, int stride, const PenCoef& c) {
  if (pen == kRidge) {
    for (int k = 0; k < K; ++k) w[k * stride] -= c.step * gs[k * stride];
  } else if (pen == kElasticNet) {
    for (int k = 0; k < K; ++k) {
      double v = w[k * stride] - c.step * gs[k * stride];
      w[k * stride] = soft_threshold(v, c.thr);
    }
  } else {
    double sq = 0.0;
    for (int k = 0; k < K; ++k) {
      double v = w[k * stride] - c.step * gs[k * stride];
      w[k * stride] = v;
      sq += v * v;
    }
    const double factor = c.bgs / sqrt(sq);
    if (factor < 1.0) {
      const double mult = 1.0 - factor / c.w_scale;
      for (int k = 0; k < K; ++k) w[k * stride] *= mult;
    } else {
      for (int k = 0; k < K; ++k) w[k * stride] = 0.0;
    }
  }
}

// K == 1 form used by the hot loops (same operations, no loops)
__device__ __forceinline__ double penalty_scalar(int pen, double w, double gs, const PenCoef& c) {
  double v = w - c.step * gs;
  if (pen == kRidge) return v;
  if (pen == kElasticNet) return soft_threshold(v, c.thr);
  const double factor = c.bgs / sqrt(v * v);
  return (factor < 1.0) ? v * (1.0 - factor / c.w_scale) : 0.0;
}

// sgd_exp for the solver's serial chain: the same operations and the same bits as sgd_exp (include/sgdnet_arith.h),
// arranged so that arguments whose result is a normal number run straight through (one never-taken branch at the
// end instead of five early exits); everything else goes to sgd_exp itself.
static __device__ __noinline__ double sgd_exp_rare(double x) { return sgd_exp(x); }
// keeps a value in its register: ptxas would otherwise re-derive shared-window addresses (S2UR + ULEA) at every use
__device__ __forceinline__ uint32_t pin_u32(uint32_t v) {
  uint32_t o;
  asm volatile("mov.u32 %0, %1;" : "=r"(o) : "r"(v));
  return o;
}
// returns v through a select on `flag`; `zero` is a run-time zero (FitDev::pad0_) neither nvcc nor ptxas can fold,
// so the result carries a true data dependence on `flag` that the hardware has to honour
__device__ __forceinline__ uint32_t dep_on(uint32_t v, bool flag, uint32_t zero) { return flag ? v : v + zero; }
__device__ __forceinline__ double sgd_exp_inrange(double x) {
  const double kInvStep = 46.16624130844683, kStepHi = 0.021660849392446835, kStepLo = 5.145609244655338e-14;
  const double kShift = 6755399441055744.0;
  const double ts = fma(x, kInvStep, kShift);
  const double kd = ts - kShift;
  // kd is an integer of magnitude < 2^31 here, and kShift = 1.5 * 2^52 puts it in the low word of ts: the same
  // value as (int32_t)kd without a conversion on the dependent path
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
  const double thi = sgd_exp_tab_dev[2 * j], tlo = sgd_exp_tab_dev[2 * j + 1];
  const double res = thi + fma(thi, q, tlo);
  const double out = res * __hiloint2double((m + 1023) << 20, 0);
  if (__builtin_expect(!(x >= -707.0 && x <= 709.0), 0)) return sgd_exp_rare(x);   // also NaN; m + 1023 may leave [2, 2046]
  return out;
}

// ---- families, scalar (K == 1)
__device__ __forceinline__ double gradient_scalar(int family, double lp, double y) {
  if (family == kBinomial) return 1.0 - y - 1.0 / (1.0 + sgd_exp(lp));
  return lp - y;   // gaussian
}
__device__ __forceinline__ double loss_scalar(int family, double lp, double y) {
  if (family == kBinomial) return sgd_log(1.0 + sgd_exp(lp)) - y * lp;
  return 0.5 * (lp - y) * (lp - y);
}

// a / n with n an integer below 2^32: q = a*(1/n) corrected once with the exact residual is the correctly rounded
// quotient (the residual step leaves a relative error of 2^-104 while a/n stays 2^-86 away from every rounding
// boundary because n has at most 32 significant bits), so this returns the same bits as the IEEE division the
// reference performs. Zero is returned as it is (keeps its sign); operands outside the safe exponent range take
// the division.
static __device__ __noinline__ double div_by_n_rare(double a, double nd) { return a / nd; }
__device__ __forceinline__ double div_by_n(double a, double nd, double rn) {
  // safe when the biased exponent is in [127, 1919] (|a| in [2^-896, 2^897)): integer test on the high word
  const uint32_t hi = static_cast<uint32_t>(__double2hiint(a));
  const uint32_t ex = (hi >> 20) & 0x7ffu;
  const bool ok = (ex - 127u) <= 1792u;
  // the quotient first, the (never taken) way out behind it: the branch then resolves under the latency of the three
  // dependent operations instead of in front of them (634 -> 618 cycles per update in the wavefront kernel's chain);
  // "a is a zero" is an integer test as well
  const double q0 = a * rn;
  const double r0 = fma(-q0, nd, a);
  const double q1 = fma(r0, rn, q0);
  const bool zero = ((hi << 1) | static_cast<uint32_t>(__double2loint(a))) == 0u;
  double res = ok ? q1 : a;       // a == +-0 when !ok && zero
  if (__builtin_expect(!ok && !zero, 0)) res = div_by_n_rare(a, nd);
  return res;
}

// ---- families, K-vector held one class per lane of a warp (K <= 32 per call chunk is handled by callers through
// shared memory for larger K). lp/g are per-lane values for class `lane`; lanes >= K pass lp = -inf.
__device__ __forceinline__ double lse_warp(double lp_lane, bool valid) {
  double mx = warp_max(valid ? lp_lane : -INFINITY);
  double e = valid ? sgd_exp(lp_lane - mx) : 0.0;
  double s = warp_sum(e);          // butterfly over 32 slots padded with zeros (sgdnet_arith.h, item 2)
  return sgd_log(s) + mx;
}

// ---- mbarrier / bulk-copy (TMA 1-D) primitives
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// one non-blocking probe of the barrier phase (the hardware may suspend the thread for a bounded time inside)
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking probe (never suspends)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- the same primitives on 32-bit shared-window addresses (computed once per role, so that the hot loops do not
// re-derive the shared window base around every access) and with the predicate carried into the instruction
__device__ __forceinline__ double lds_f64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void mbar_arrive_if(bool pred, uint32_t a) {
  asm volatile(
      "{\n"
      ".reg .pred pp;\n"
      "setp.ne.b32 pp, %1, 0;\n"
      "@pp mbarrier.arrive.shared::cta.b64 _, [%0];\n"
      "}\n" ::"r"(a),
      "r"(static_cast<int>(pred))
      : "memory");
}
__device__ __forceinline__ bool mbar_test_wait_a(uint32_t a, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(a), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_a(uint32_t a, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(a),
      "r"(parity)
      : "memory");
}

// global -> shared bulk copy (UBLKCP in SASS); bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

}  // namespace sgd
