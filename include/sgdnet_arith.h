/*
 * sgdnet_arith.h — the arithmetic specification of libsgdnet_b200.so (host + device, header only).
 *
 * Why this exists. sgdnet's convergence test (reference src/utils.h:240-262) is decided, on every binomial or
 * multinomial path, by coefficients that decay geometrically until the soft threshold (src/prox.h:32-39) sets them to
 * exactly zero: the epoch at which that happens depends on rounding at the 1e-16 level, and every later lambda then
 * sees a shifted sampling sequence. The reference is therefore only reproducible for one (libm, Eigen, CPU) triple.
 * To make "same supports, same path lengths" a checkable statement, the library pins the two things the reference
 * leaves to its platform:
 *
 *  1. exp and log inside the solver loop are the functions below, built from IEEE-754 +,-,*,/ and fma only, so a CPU
 *     and a GPU evaluate them bit-identically. Both are faithful (< 1 ulp; tests/test_arith.py measures it against
 *     200-bit references), i.e. they differ from glibc's exp/log in the last bit at most.
 *  2. reductions inside the solver loop have a fixed association order (documented where they are used:
 *     sgdnet_b200/csrc/saga_*.cu, DESIGN.md section "arithmetic"): a dense dot product over p features is 256
 *     interleaved running sums (feature j -> sum j mod 256, ascending j), each group of 32 consecutive sums is
 *     combined by the xor-butterfly 16,8,4,2,1, the 8 group results are added in ascending order; from
 *     SGD_WIDE_P = 512 features on the same scheme runs eight blocks wide: 2048 interleaved running sums, butterfly
 *     per 32, the 8 group results of each block of 256 sums added in ascending order, then the 8 block sums added
 *     in ascending order (one block per CTA of the solver's thread-block cluster); a sparse row dot
 *     product is 32 interleaved running sums over the row's nonzero positions, combined by the same butterfly; a
 *     sum over K <= 32 classes is the butterfly over 32 slots padded with zeros.
 *
 * Everything else in the loop is elementwise IEEE arithmetic in the reference's own operation order, compiled
 * without multiply-add contraction (-fmad=false / -ffp-contract=off).
 */
#ifndef SGDNET_ARITH_H_
#define SGDNET_ARITH_H_

#include <math.h>
#include <stdint.h>
#include <string.h>

#define SGD_WIDE_P 512   /* dense designs with at least this many features use the eight-block dot product */

#if defined(__CUDACC__)
#define SGD_HD __host__ __device__ __forceinline__
#else
#define SGD_HD static inline
#endif

SGD_HD uint64_t sgd_bits(double x) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u;
  memcpy(&u, &x, sizeof(u));
  return u;
#endif
}
SGD_HD double sgd_from_bits(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double x;
  memcpy(&x, &u, sizeof(x));
  return x;
#endif
}

/* 2^(j/32), j = 0..31, as an unevaluated sum hi + lo (hi = nearest double, lo = the remainder). */
#define SGD_EXP_TABLE_INIT \
  1.0, 0.0, \
  1.0218971486541166, 5.109225028973444e-17, \
  1.0442737824274138, 8.551889705537965e-17, \
  1.0671404006768237, -7.899853966841582e-17, \
  1.0905077326652577, -3.046782079812471e-17, \
  1.1143867425958924, 1.0410278456845571e-16, \
  1.1387886347566916, 8.912812676025408e-17, \
  1.1637248587775775, 3.8292048369240935e-17, \
  1.189207115002721, 3.982015231465646e-17, \
  1.215247359980469, -7.712630692681488e-17, \
  1.241857812073484, 4.658027591836937e-17, \
  1.2690509571917332, 2.667932131342186e-18, \
  1.2968395546510096, 2.5382502794888315e-17, \
  1.3252366431597413, -2.8587312100388614e-17, \
  1.3542555469368927, 7.70094837980299e-17, \
  1.383909881963832, -6.770511658794786e-17, \
  1.4142135623730951, -9.667293313452913e-17, \
  1.4451808069770467, -3.0237581349939873e-17, \
  1.4768261459394993, -3.483994556892796e-17, \
  1.5091644275934228, -1.016455327754295e-16, \
  1.5422108254079407, 7.949834809697621e-17, \
  1.5759808451078865, -1.0136916471278304e-17, \
  1.6104903319492543, 2.4707192569797888e-17, \
  1.645755478153965, -1.0125679913674773e-16, \
  1.681792830507429, 8.199010020581497e-17, \
  1.718619298122478, -1.851380418263111e-17, \
  1.7562521603732995, 2.960140695448873e-17, \
  1.7947090750031072, 1.8227458427912087e-17, \
  1.8340080864093424, 3.283107224245627e-17, \
  1.8741676341103, -6.122763413004143e-17, \
  1.9152065613971474, -1.0619946056195963e-16, \
  1.9571441241754002, 8.960767791036668e-17,

static const double sgd_exp_tab_host[64] = {SGD_EXP_TABLE_INIT};
#if defined(__CUDACC__)
static __device__ __constant__ double sgd_exp_tab_dev[64] = {SGD_EXP_TABLE_INIT};
#endif

/*
 * exp(x): x = k*(ln2/32) + r, |r| <= ln2/64; e^x = 2^(k>>5) * 2^((k&31)/32) * (1 + expm1(r)),
 * expm1(r) by its degree-6 Taylor polynomial (remainder < 4e-18), table value carried as hi + lo.
 */
SGD_HD double sgd_exp(double x) {
  if (!(x == x)) return x;
  if (x > 709.782712893384) return INFINITY;
  if (x < -745.1332191019412) return 0.0;
  const double kInvStep = 46.16624130844683;          /* 32/ln2 */
  const double kStepHi = 0.021660849392446835;        /* ln2/32, top 36 bits */
  const double kStepLo = 5.145609244655338e-14;
  const double kShift = 6755399441055744.0;           /* 1.5 * 2^52: rounds to nearest integer */
  const double kd = fma(x, kInvStep, kShift) - kShift;
  const int32_t k = (int32_t)kd;
  double r = fma(-kd, kStepHi, x);
  r = fma(-kd, kStepLo, r);
  double p = 1.0 / 720.0;
  p = fma(p, r, 1.0 / 120.0);
  p = fma(p, r, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  const double q = fma(r * r, p, r);                   /* expm1(r) */
  const int32_t j = k & 31;
  const int32_t m = k >> 5;
#if defined(__CUDA_ARCH__)
  const double thi = sgd_exp_tab_dev[2 * j], tlo = sgd_exp_tab_dev[2 * j + 1];
#else
  const double thi = sgd_exp_tab_host[2 * j], tlo = sgd_exp_tab_host[2 * j + 1];
#endif
  const double res = thi + fma(thi, q, tlo);
  if (m >= -1021 && m <= 1023) return res * sgd_from_bits((uint64_t)(m + 1023) << 52);
  if (m > 1023) return (res * sgd_from_bits((uint64_t)(m - 1 + 1023) << 52)) * 2.0;
  return (res * sgd_from_bits((uint64_t)(m + 1000 + 1023) << 52)) * sgd_from_bits((uint64_t)(-1000 + 1023) << 52);
}

/*
 * log(x): x = 2^k * m, m in [sqrt(1/2), sqrt(2)); f = m - 1, s = f/(2+f), log(1+f) = 2s + s*R(s^2) with the
 * classical degree-14 odd minimax polynomial; assembled as k*ln2_hi - ((f^2/2 - (s*(f^2/2 + R) + k*ln2_lo)) - f).
 * Only +,-,*,/ : bit-identical wherever IEEE-754 double arithmetic is.
 */
SGD_HD double sgd_log(double x) {
  if (!(x == x)) return x;
  if (x < 0.0) return NAN;
  if (x == 0.0) return -INFINITY;
  if (x == INFINITY) return x;
  int32_t k = 0;
  uint64_t u = sgd_bits(x);
  if ((u >> 52) == 0) {                               /* subnormal: scale up by 2^54 */
    x *= 18014398509481984.0;
    k = -54;
    u = sgd_bits(x);
  }
  k += (int32_t)((u >> 52) & 0x7ff) - 1023;
  uint64_t mant = u & 0x000fffffffffffffULL;
  if (mant > 0x6a09e667f3bccULL) {                    /* m >= sqrt(2): use m/2 and k+1 */
    u = mant | 0x3fe0000000000000ULL;
    k += 1;
  } else {
    u = mant | 0x3ff0000000000000ULL;
  }
  const double f = sgd_from_bits(u) - 1.0;
  const double dk = (double)k;
  const double kLn2Hi = 6.93147180369123816490e-01, kLn2Lo = 1.90821492927058770002e-10;
  const double s = f / (2.0 + f);
  const double z = s * s;
  const double w = z * z;
  const double t1 = w * (3.999999999940941908e-01 + w * (2.222219843214978396e-01 + w * 1.531383769920937332e-01));
  const double t2 = z * (6.666666666666735130e-01 +
                         w * (2.857142874366239149e-01 + w * (1.818357216161805012e-01 + w * 1.479819860511658591e-01)));
  const double R = t2 + t1;
  const double hfsq = 0.5 * f * f;
  return dk * kLn2Hi - ((hfsq - (s * (hfsq + R) + dk * kLn2Lo)) - f);
}

#endif /* SGDNET_ARITH_H_ */
