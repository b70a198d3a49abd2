/*
 * sgdnet_oracle.cpp — CPU restatement of sgdnet's SAGA path.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing in the product (sgdnet_b200/, libsgdnet_b200.so) includes, links or calls this file.
 * It is used by tests/, by __graft_entry__.smoke() as the checker, and by bench.py's
 * cpu_baseline / --impl reference legs as the timed CPU arm.
 *
 * What it follows (all paths relative to /root/reference):
 *   driver            src/sgdnet.cpp:119-285 (SetupSgdnet), :71-100 (RunSaga penalty choice)
 *   dense solver      src/saga-dense.h:127-223
 *   sparse solver     src/saga-sparse.h:76-155 (LaggedUpdate/AddWeighted/Reset), :222-382 (Saga)
 *   penalties / prox  src/penalties.h:27-79, src/prox.h:32-39
 *   families          src/families.h:64-410
 *   utilities         src/utils.h:31-378, src/math.h:25-199, src/constants.h:22
 *   sampling          R core RNG (MT19937 + set.seed scrambling + unif_rand fixup), call sites
 *                     src/saga-dense.h:152, src/saga-sparse.h:261
 *
 * PARITY STATUS.  PINNED AGAINST THE REFERENCE'S OWN CODE.  R, Rcpp and Eigen are not in the image, so the reference's
 * package cannot be built; but its solver sources compile unmodified, from where they lie, against a minimal stand-in
 * for the Rcpp/Eigen subset they use (oracle/refbuild/: recipe, stand-in header, C entry points -> oracle/_ref/
 * libsgdnet_ref.so).  In "libm" mode this oracle reproduces that build BIT FOR BIT - lambda path, epochs per lambda,
 * return codes, coefficients, intercepts, deviance ratios, null deviance, debug losses - on the bundled datasets and on
 * a grid of families x penalties x {dense, sparse} x intercept x standardize (tests/test_ref_cpu.py, live against the
 * library and against the committed outputs tests/golden/ref_vectors.npz).  Also pinned by (i) the R RNG known-answer
 * values, (ii) the reference tests' own properties re-expressed in tests/test_oracle_cpu.py (closed forms, lambda_max
 * formulas, null deviances, sparse==dense, ...).  What stays unpinned is the part of a real R build that Eigen leaves
 * implementation-defined: the association of dense reductions (GEMV, .sum(), .norm()) and packet exp/log; the stand-in
 * and this oracle's libm mode use plain ascending sequential sums and std::exp/std::log.
 *
 * ARITHMETIC MODES (oracle_set_arith / env SGDNET_ORACLE_ARITH):
 *   1 "portable" (default): inside the solver loop, exp/log are sgd_exp/sgd_log and the dot products / class sums use
 *      the fixed association order of include/sgdnet_arith.h. This is the arithmetic the GPU library is specified to
 *      use, so supports, epoch counts and even coefficients can be compared exactly (see the header of
 *      sgdnet_arith.h for why an exact comparison needs this).
 *   0 "libm": std::exp/std::log and plain ascending sequential sums - the most literal reading of the reference
 *      (Eigen's sparse products are sequential; its dense GEMV order is implementation-defined). Used for the CPU
 *      baseline timings and to bound how much the arithmetic choice matters (tests/test_oracle_cpu.py).
 * Everything outside the solver loop (setup, deviance, rescale, scoring) is identical in both modes.
 *
 * Build: g++ -O2 -ffp-contract=off (R's default -O2, no FMA contraction, no -march=native).
 */
#include "../include/sgdnet_b200.h"
#include "../include/sgdnet_arith.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;
int g_arith = 1;   /* 1 portable, 0 libm */

inline double loop_exp(double x) { return g_arith ? sgd_exp(x) : std::exp(x); }
inline double loop_log(double x) { return g_arith ? sgd_log(x) : std::log(x); }

/* xor-butterfly 16,8,4,2,1 over 32 slots: the value every lane of a warp ends with after an all-reduce */
inline double butterfly32(double* v) {
  double t[32];
  for (int o = 16; o > 0; o >>= 1) {
    for (int i = 0; i < 32; ++i) t[i] = v[i] + v[i ^ o];
    for (int i = 0; i < 32; ++i) v[i] = t[i];
  }
  return v[0];
}

/* dense dot product sum_j a[j*sa] * b[j]: 256 interleaved running sums, butterfly per 32, 8 groups ascending */
inline double dot_dense(const double* a, int64_t sa, const double* b, int64_t p) {
  if (!g_arith) {
    double acc = 0.0;
    for (int64_t j = 0; j < p; ++j) acc += a[j * sa] * b[j];
    return acc;
  }
  double chain[256];
  for (int i = 0; i < 256; ++i) chain[i] = 0.0;
  for (int64_t j = 0; j < p; ++j) chain[j & 255] += a[j * sa] * b[j];
  double total = 0.0;
  for (int g = 0; g < 8; ++g) total += butterfly32(chain + 32 * g);
  return total;
}

/* wide dense dot product (p >= SGD_WIDE_P, include/sgdnet_arith.h item 2): 2048 interleaved running sums (feature j ->
   sum j mod 2048, ascending j), butterfly per 32 consecutive sums, the 8 results of each block of 256 sums added in
   ascending order, the 8 block sums added in ascending order */
inline double dot_dense_wide(const double* a, int64_t sa, const double* b, int64_t p) {
  static thread_local double chain[2048];
  for (int i = 0; i < 2048; ++i) chain[i] = 0.0;
  for (int64_t j = 0; j < p; ++j) chain[j & 2047] += a[j * sa] * b[j];
  double total = 0.0;
  for (int c = 0; c < 8; ++c) {
    double block = 0.0;
    for (int g = 0; g < 8; ++g) block += butterfly32(chain + 256 * c + 32 * g);
    total += block;
  }
  return total;
}

/* sparse row dot product sum_e val[e] * w[idx[e]*sw]: 32 interleaved running sums over positions, butterfly */
inline double dot_sparse(const double* val, const int32_t* idx, int64_t nnz, const double* w, int64_t sw) {
  if (!g_arith) {
    double acc = 0.0;
    for (int64_t e = 0; e < nnz; ++e) acc += val[e] * w[static_cast<size_t>(idx[e]) * sw];
    return acc;
  }
  double chain[32];
  for (int i = 0; i < 32; ++i) chain[i] = 0.0;
  for (int64_t e = 0; e < nnz; ++e) chain[e & 31] += val[e] * w[static_cast<size_t>(idx[e]) * sw];
  return butterfly32(chain);
}

/* ------------------------------------------------------------------ R-compatible RNG */
/* R core src/main/RNG.c: MT_sgenrand / MT_genrand, set.seed -> Randomize -> RNG_Init, fixup(). */
constexpr int MT_N = 624, MT_M = 397;

void mt_set_seed(sgdnet_rng* r, uint32_t seed) {
  for (int j = 0; j < 50; ++j) seed = 69069u * seed + 1u;     /* initial scrambling */
  uint32_t filled[MT_N + 1];
  for (int j = 0; j < MT_N + 1; ++j) { seed = 69069u * seed + 1u; filled[j] = seed; }
  /* i_seed[0] is `mti`; FixupSeeds forces it to N so the first draw regenerates the block */
  r->mti = MT_N;
  for (int j = 0; j < MT_N; ++j) r->mt[j] = filled[j + 1];
}

uint32_t mt_next(sgdnet_rng* r) {
  uint32_t* mt = r->mt;
  if (r->mti >= MT_N) {
    auto twist = [](uint32_t u, uint32_t v) {
      uint32_t y = (u & 0x80000000u) | (v & 0x7fffffffu);
      return (y >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
    };
    int k = 0;
    for (; k < MT_N - MT_M; ++k) mt[k] = mt[k + MT_M] ^ twist(mt[k], mt[k + 1]);
    for (; k < MT_N - 1; ++k)    mt[k] = mt[k + (MT_M - MT_N)] ^ twist(mt[k], mt[k + 1]);
    mt[MT_N - 1] = mt[MT_M - 1] ^ twist(mt[MT_N - 1], mt[0]);
    r->mti = 0;
  }
  uint32_t y = mt[r->mti++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}

double mt_unif(sgdnet_rng* r) {
  double v = mt_next(r) * 2.3283064365386963e-10;   /* [0,1) */
  const double i2_32m1 = 2.328306437080797e-10;     /* 1/(2^32 - 1) */
  if (v <= 0.0) return 0.5 * i2_32m1;
  if (1.0 - v <= 0.0) return 1.0 - 0.5 * i2_32m1;
  return v;
}

/* one sample index: floor(runif(0, n)) = floor(0 + (n - 0) * u) */
bool draw_index(sgdnet_rng* r, uint32_t n, uint32_t* out) {
  switch (r->kind) {
    case SGDNET_RNG_MT:
      *out = static_cast<uint32_t>(std::floor(0.0 + (static_cast<double>(n) - 0.0) * mt_unif(r)));
      return true;
    case SGDNET_RNG_CALLBACK:
      if (!r->unif_rand) return false;
      *out = static_cast<uint32_t>(std::floor(0.0 + (static_cast<double>(n) - 0.0) * r->unif_rand(r->ctx)));
      return true;
    case SGDNET_RNG_SEQUENCE:
      if (!r->seq || r->seq_pos >= r->seq_len) return false;
      *out = r->seq[r->seq_pos++];
      return true;
  }
  return false;
}

/* ------------------------------------------------------------------ containers */
constexpr double SMALL = 100 * std::numeric_limits<double>::epsilon();   /* src/constants.h:22 */

struct Design {            /* the feature matrix with samples as the fast-access unit (after
                              AdaptiveTranspose, src/utils.h:276-288) */
  bool sparse = false;
  int64_t n = 0, p = 0;
  std::vector<double> dense;       /* [n][p] sample-major */
  std::vector<int64_t> rp;         /* CSR row pointers */
  std::vector<int32_t> ci;         /* CSR column ids, ascending in each row */
  std::vector<double> cv;
};

enum Pen { RIDGE = 0, ENET = 1, GROUP = 2 };

struct Model {
  int family = 0;
  int K = 1;          /* n_classes */
  int Ky = 1;         /* columns of y */
  bool fit_intercept = true, standardize = true, is_sparse = false;
};

/* ------------------------------------------------------------------ math.h pieces */
double log_sum_exp(const double* x, int K) {            /* src/math.h:25-33 */
  double mx = x[0];
  for (int k = 1; k < K; ++k) mx = std::max(mx, x[k]);
  double s = 0.0;
  for (int k = 0; k < K; ++k) s += std::exp(x[k] - mx);
  return std::log(s) + mx;
}

double soft_threshold(double x, double s) {             /* src/prox.h:32-39 */
  return std::max(x - s, 0.0) - std::max(-x - s, 0.0);
}

double clampd(double x, double lo, double hi) { return x > hi ? hi : (x < lo ? lo : x); }   /* math.h:167-172 */

/* column mean / population sd of an n x m column-major block (src/math.h:66-79, 114-130) */
void col_mean(const double* a, int64_t n, int m, std::vector<double>& mean) {
  mean.assign(m, 0.0);
  for (int j = 0; j < m; ++j) {
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) s += a[j * n + i];
    mean[j] = s / static_cast<double>(n);
  }
}
void col_sd(const double* a, int64_t n, int m, const std::vector<double>& mean, std::vector<double>& sd) {
  sd.assign(m, 0.0);
  for (int j = 0; j < m; ++j) {
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) { double d = a[j * n + i] - mean[j]; s += d * d; }
    double var = s / static_cast<double>(n);
    sd[j] = (var == 0.0) ? 1.0 : std::sqrt(var);
  }
}

/* ------------------------------------------------------------------ families (src/families.h) */
/* y is K_y x n "samples in columns": yt[s*Ky + k] */
double family_loss(const Model& m, const double* lp, const double* yt, int64_t s) {
  switch (m.family) {
    case SGDNET_GAUSSIAN: {                                   /* :81-87 */
      double y = yt[s];
      return 0.5 * (lp[0] - y) * (lp[0] - y);
    }
    case SGDNET_BINOMIAL:                                     /* :152-159 */
      return std::log(1.0 + std::exp(lp[0])) - yt[s] * lp[0];
    case SGDNET_MULTINOMIAL: {                                /* :235-242 */
      unsigned c = static_cast<unsigned>(yt[s] + 0.5);
      return log_sum_exp(lp, m.K) - lp[c];
    }
    default: {                                                /* :350-356 */
      double acc = 0.0;
      for (int k = 0; k < m.K; ++k) { double d = lp[k] - yt[s * m.Ky + k]; acc += d * d; }
      return 0.5 * acc;
    }
  }
}

/* LogSumExp inside the solver loop (arithmetic mode aware) */
double loop_lse(const double* x, int K) {
  double mx = x[0];
  for (int k = 1; k < K; ++k) mx = std::max(mx, x[k]);
  double total;
  if (g_arith && K <= 32) {
    double slot[32];
    for (int k = 0; k < 32; ++k) slot[k] = (k < K) ? loop_exp(x[k] - mx) : 0.0;
    total = butterfly32(slot);
  } else {
    total = 0.0;
    for (int k = 0; k < K; ++k) total += loop_exp(x[k] - mx);
  }
  return loop_log(total) + mx;
}

void family_gradient(const Model& m, const double* lp, const double* yt, int64_t s, double* g) {
  switch (m.family) {
    case SGDNET_GAUSSIAN: g[0] = lp[0] - yt[s]; break;                                   /* :89-96 */
    case SGDNET_BINOMIAL: g[0] = 1.0 - yt[s] - 1.0 / (1.0 + loop_exp(lp[0])); break;      /* :161-168 */
    case SGDNET_MULTINOMIAL: {                                                           /* :244-260 */
      double lse = loop_lse(lp, m.K);
      unsigned c = static_cast<unsigned>(yt[s] + 0.5);
      for (int k = 0; k < m.K; ++k) {
        g[k] = loop_exp(lp[k] - lse);
        if (static_cast<unsigned>(k) == c) g[k] -= 1.0;
      }
      break;
    }
    default:                                                                             /* :358-365 */
      for (int k = 0; k < m.K; ++k) g[k] = lp[k] - yt[s * m.Ky + k];
  }
}

double binomial_link(double y) {                              /* :139-150 */
  double z = clampd(y, 1e-9, 1.0 - 1e-9);
  return std::log(z / (1.0 - z));
}

void proportions(const double* yt, int64_t n, int K, std::vector<double>& pr) {   /* math.h:184-199 */
  pr.assign(K, 0.0);
  for (int64_t i = 0; i < n; ++i) {
    int64_t c = static_cast<int64_t>(yt[i] + 0.5);
    pr[c] += 1.0 / static_cast<double>(n);
  }
}

/* the "null model" linear predictor shared by NullDeviance and FitNullModel */
void null_predictor(const Model& m, const double* yt, int64_t n, std::vector<double>& lp) {
  lp.assign(m.K, 0.0);
  switch (m.family) {
    case SGDNET_GAUSSIAN:
    case SGDNET_MGAUSSIAN:                                    /* :98-117, :367-385: mean of y, always */
      for (int k = 0; k < m.K; ++k) {
        double s = 0.0;
        for (int64_t i = 0; i < n; ++i) s += yt[i * m.Ky + k];
        lp[k] = s / static_cast<double>(n);
      }
      break;
    case SGDNET_BINOMIAL:                                     /* :170-201 */
      if (m.fit_intercept) {
        double s = 0.0;
        for (int64_t i = 0; i < n; ++i) s += yt[i];
        lp[0] = binomial_link(s / static_cast<double>(n));
      } else {
        lp[0] = 0.0;
      }
      break;
    case SGDNET_MULTINOMIAL: {                                /* :262-298 */
      std::vector<double> pr;
      if (m.fit_intercept) proportions(yt, n, m.K, pr);
      else pr.assign(m.K, 1.0 / m.K);
      double slog = 0.0;
      for (int k = 0; k < m.K; ++k) slog += std::log(pr[k]);
      for (int k = 0; k < m.K; ++k) lp[k] = std::log(pr[k]) - slog / m.K;
      break;
    }
  }
}

double null_deviance(const Model& m, const double* yt, int64_t n) {
  std::vector<double> lp;
  null_predictor(m, yt, n, lp);
  double loss = 0.0;
  if (m.family == SGDNET_MULTINOMIAL) {                       /* :276-284: lse hoisted out of the loop */
    double lse = log_sum_exp(lp.data(), m.K);
    for (int64_t i = 0; i < n; ++i) {
      unsigned c = static_cast<unsigned>(yt[i] + 0.5);
      loss += lse - lp[c];
    }
  } else {
    for (int64_t i = 0; i < n; ++i) loss += family_loss(m, lp.data(), yt, i);
  }
  return 2.0 * loss;
}

/* ------------------------------------------------------------------ X^T * Ymap for LambdaMax */
/* x still "samples in rows" conceptually; we hold it sample-major, so accumulate per feature in
   ascending sample order, which is the order a column-of-X dot product visits them. */
void xt_times(const Design& d, const std::vector<double>& ymap /* [m][n] column-major n x m */, int m,
              std::vector<double>& out /* [m][p] */) {
  out.assign(static_cast<size_t>(m) * d.p, 0.0);
  for (int c = 0; c < m; ++c) {
    const double* yc = &ymap[static_cast<size_t>(c) * d.n];
    double* oc = &out[static_cast<size_t>(c) * d.p];
    if (d.sparse) {
      for (int64_t s = 0; s < d.n; ++s)
        for (int64_t e = d.rp[s]; e < d.rp[s + 1]; ++e) oc[d.ci[e]] += d.cv[e] * yc[s];
    } else {
      for (int64_t s = 0; s < d.n; ++s) {
        const double* xs = &d.dense[s * d.p];
        for (int64_t j = 0; j < d.p; ++j) oc[j] += xs[j] * yc[s];
      }
    }
  }
}

double lambda_max(const Model& m, const Design& d, const std::vector<double>& y /* n x Ky col-major, preprocessed */,
                  const std::vector<double>& y_scale) {
  const int64_t n = d.n;
  std::vector<double> ip;
  switch (m.family) {
    case SGDNET_GAUSSIAN: {                                   /* :119-126 */
      xt_times(d, y, 1, ip);
      double mx = 0.0;
      for (int64_t j = 0; j < d.p; ++j) mx = std::max(mx, std::fabs(ip[j]));
      return y_scale[0] * mx / static_cast<double>(n);
    }
    case SGDNET_BINOMIAL: {                                   /* :203-220 */
      std::vector<double> mean, sd, ymap(n);
      col_mean(y.data(), n, 1, mean);
      col_sd(y.data(), n, 1, mean, sd);
      for (int64_t i = 0; i < n; ++i) ymap[i] = (y[i] - mean[0]) / sd[0];
      xt_times(d, ymap, 1, ip);
      double mx = 0.0;
      for (int64_t j = 0; j < d.p; ++j) mx = std::max(mx, std::fabs(ip[j]));
      return sd[0] * mx / static_cast<double>(n);
    }
    case SGDNET_MULTINOMIAL: {                                /* :300-325 */
      std::vector<double> ymap(static_cast<size_t>(n) * m.K, 0.0), mean, sd;
      for (int64_t i = 0; i < n; ++i) {
        unsigned c = static_cast<unsigned>(y[i] + 0.5);
        ymap[static_cast<size_t>(c) * n + i] = 1.0;
      }
      col_mean(ymap.data(), n, m.K, mean);
      col_sd(ymap.data(), n, m.K, mean, sd);
      for (int k = 0; k < m.K; ++k)
        for (int64_t i = 0; i < n; ++i) {
          double& v = ymap[static_cast<size_t>(k) * n + i];
          v = (v - mean[k]) / sd[k];
        }
      xt_times(d, ymap, m.K, ip);
      double mx = 0.0;
      for (int k = 0; k < m.K; ++k)
        for (int64_t j = 0; j < d.p; ++j) mx = std::max(mx, std::fabs(ip[static_cast<size_t>(k) * d.p + j] * sd[k]));
      return mx / static_cast<double>(n);
    }
    default: {                                                /* :387-406 */
      std::vector<double> ymap(y), mean, sd;
      col_mean(y.data(), n, m.K, mean);
      col_sd(y.data(), n, m.K, mean, sd);
      for (int k = 0; k < m.K; ++k)
        for (int64_t i = 0; i < n; ++i) {
          double& v = ymap[static_cast<size_t>(k) * n + i];
          v = (v - mean[k]) / sd[k];
        }
      xt_times(d, ymap, m.K, ip);
      double mx = 0.0;
      for (int64_t j = 0; j < d.p; ++j) {
        double acc = 0.0;
        for (int k = 0; k < m.K; ++k) {
          double v = ip[static_cast<size_t>(k) * d.p + j] * (y_scale[k] * sd[k]);
          acc += v * v;
        }
        mx = std::max(mx, std::sqrt(acc));
      }
      return mx / static_cast<double>(n);
    }
  }
}

/* ------------------------------------------------------------------ penalties (src/penalties.h) */
inline void penalty_apply(Pen pen, double* w, const double* gs, int K, double gamma, double beta,
                          double w_scale, double scaling) {
  const double step = gamma / w_scale * scaling;
  switch (pen) {
    case RIDGE:                                               /* :27-39 */
      for (int k = 0; k < K; ++k) w[k] -= step * gs[k];
      break;
    case ENET: {                                              /* :41-59 */
      for (int k = 0; k < K; ++k) {
        w[k] -= step * gs[k];
        w[k] = soft_threshold(w[k], beta * gamma * scaling / w_scale);
      }
      break;
    }
    case GROUP: {                                             /* :61-79 */
      for (int k = 0; k < K; ++k) w[k] -= step * gs[k];
      double sq = 0.0;
      for (int k = 0; k < K; ++k) sq += w[k] * w[k];
      double factor = beta * gamma * scaling / std::sqrt(sq);
      if (factor < 1.0) {
        double mult = 1.0 - factor / w_scale;
        for (int k = 0; k < K; ++k) w[k] *= mult;
      } else {
        for (int k = 0; k < K; ++k) w[k] = 0.0;
      }
      break;
    }
  }
}

/* ------------------------------------------------------------------ solver state */
struct State {
  std::vector<double> W;        /* [p][K]  (K x p column-major) */
  std::vector<double> b;        /* [K] */
  std::vector<double> gmem;     /* [n][K] */
  std::vector<double> gsum;     /* [p][K] */
  std::vector<double> gsi;      /* [K] */
};

/* ConvergenceCheck (src/utils.h:240-262) */
struct Convergence {
  std::vector<double> prev;
  double tol;
  bool operator()(const std::vector<double>& w) {
    double max_change = 0.0, max_size = 0.0;
    for (size_t i = 0; i < w.size(); ++i) {
      max_change = std::max(max_change, std::fabs(w[i] - prev[i]));
      max_size = std::max(max_size, std::fabs(w[i]));
    }
    bool all_zero = (max_size == 0.0) && (max_change == 0.0);
    bool no_change = (max_size != 0.0) && (max_change / max_size <= tol);
    prev = w;
    return all_zero || no_change;
  }
};

/* linear predictor of one sample with wscale == 1 (Deviance / EpochLoss, src/utils.h:199-227, 304-329) */
void plain_predictor(const Model& m, const Design& d, const std::vector<double>& c, const State& st, int64_t s,
                     double* lp) {
  const int K = m.K;
  for (int k = 0; k < K; ++k) lp[k] = 0.0;
  if (d.sparse) {
    for (int64_t e = d.rp[s]; e < d.rp[s + 1]; ++e) {
      const double* wj = &st.W[static_cast<size_t>(d.ci[e]) * K];
      for (int k = 0; k < K; ++k) lp[k] += d.cv[e] * wj[k];
    }
  } else {
    const double* xs = &d.dense[s * d.p];
    for (int64_t j = 0; j < d.p; ++j) {
      const double* wj = &st.W[static_cast<size_t>(j) * K];
      for (int k = 0; k < K; ++k) lp[k] += wj[k] * xs[j];
    }
  }
  for (int k = 0; k < K; ++k) lp[k] += st.b[k];
  if (m.standardize && m.is_sparse) {
    for (int k = 0; k < K; ++k) {
      double wc = 0.0;
      for (int64_t j = 0; j < d.p; ++j) wc += st.W[static_cast<size_t>(j) * K + k] * c[j];
      lp[k] -= wc;
    }
  }
}

double deviance(const Model& m, const Design& d, const std::vector<double>& c, const std::vector<double>& yt,
                const State& st) {
  std::vector<double> lp(m.K);
  double loss = 0.0;
  for (int64_t s = 0; s < d.n; ++s) {
    plain_predictor(m, d, c, st, s, lp.data());
    loss += family_loss(m, lp.data(), yt.data(), s);
  }
  return 2.0 * loss;
}

double epoch_loss(const Model& m, const Design& d, const std::vector<double>& c, const std::vector<double>& yt,
                  const State& st) {
  std::vector<double> lp(m.K);
  double loss = 0.0;
  for (int64_t s = 0; s < d.n; ++s) {
    plain_predictor(m, d, c, st, s, lp.data());
    loss += family_loss(m, lp.data(), yt.data(), s) / static_cast<double>(d.n);
  }
  return loss;
}

struct SagaArgs {
  Pen pen;
  double gamma, alpha, beta;
  unsigned max_iter;
  double tol;
  bool debug;
};

/* ------------------------------------------------------------------ dense Saga (src/saga-dense.h:127-223) */
int saga_dense(const Model& m, const Design& d, const std::vector<double>& c, const std::vector<double>& yt,
               State& st, const SagaArgs& a, sgdnet_rng* rng, unsigned* epochs_out, unsigned* code_out,
               std::vector<double>& losses) {
  const int K = m.K;
  const int64_t n = d.n, p = d.p;
  const double nd = static_cast<double>(static_cast<unsigned>(n));
  double wscale = 1.0;
  const double wscale_update = 1.0 - a.alpha * a.gamma;
  std::vector<double> g(K, 0.0), gch(K, 0.0), lp(K);
  Convergence conv{st.W, a.tol};

  unsigned it_outer = 0;
  bool converged = false;
  do {
    for (unsigned it = 0; it < static_cast<unsigned>(n); ++it) {
      uint32_t s;
      if (!draw_index(rng, static_cast<uint32_t>(n), &s)) return SGDNET_ERR_RNG;
      const double* xs = &d.dense[static_cast<size_t>(s) * p];

      for (int k = 0; k < K; ++k) lp[k] = (g_arith && p >= SGD_WIDE_P) ? dot_dense_wide(&st.W[k], K, xs, p) : dot_dense(&st.W[k], K, xs, p);
      for (int k = 0; k < K; ++k) lp[k] = lp[k] * wscale + st.b[k];

      family_gradient(m, lp.data(), yt.data(), s, g.data());
      for (int k = 0; k < K; ++k) {
        gch[k] = g[k] - st.gmem[static_cast<size_t>(s) * K + k];
        st.gmem[static_cast<size_t>(s) * K + k] = g[k];
      }

      if (wscale < SMALL) {
        for (double& w : st.W) w *= wscale;
        wscale = 1.0;
      }
      wscale *= wscale_update;

      if (m.fit_intercept) {
        for (int k = 0; k < K; ++k) {
          st.gsi[k] += gch[k] / nd;
          st.b[k] -= a.gamma * (st.gsi[k] + gch[k] / nd);
        }
      }

      const double gw = a.gamma / wscale;
      for (int64_t j = 0; j < p; ++j) {
        double* wj = &st.W[static_cast<size_t>(j) * K];
        for (int k = 0; k < K; ++k) wj[k] -= gch[k] * xs[j] * gw;
      }
      for (int64_t j = 0; j < p; ++j)
        penalty_apply(a.pen, &st.W[static_cast<size_t>(j) * K], &st.gsum[static_cast<size_t>(j) * K], K, a.gamma,
                      a.beta, wscale, 1.0);
      for (int64_t j = 0; j < p; ++j) {
        double* gj = &st.gsum[static_cast<size_t>(j) * K];
        for (int k = 0; k < K; ++k) gj[k] += gch[k] * xs[j] / nd;
      }
    }

    for (double& w : st.W) w *= wscale;
    wscale = 1.0;

    if (a.debug) losses.push_back(epoch_loss(m, d, c, yt, st));
    converged = conv(st.W);
    ++it_outer;
  } while (!converged && it_outer < a.max_iter);

  *epochs_out = it_outer;
  *code_out = (it_outer == a.max_iter) ? 1u : 0u;
  return SGDNET_OK;
}

/* ------------------------------------------------------------------ sparse Saga (src/saga-sparse.h) */
int saga_sparse(const Model& m, const Design& d, const std::vector<double>& c, const std::vector<double>& yt,
                State& st, const SagaArgs& a, sgdnet_rng* rng, unsigned* epochs_out, unsigned* code_out,
                std::vector<double>& losses) {
  const int K = m.K;
  const int64_t n = d.n, p = d.p;
  const unsigned nu = static_cast<unsigned>(n);
  const double nd = static_cast<double>(nu);
  const bool stdz = m.standardize;

  std::vector<unsigned> lag(p, 0u);
  double wscale = 1.0;
  const double wscale_update = 1.0 - a.alpha * a.gamma;

  /* :229-240 — running geometric sum, rebuilt on every call */
  std::vector<double> lag_scaling;
  lag_scaling.reserve(static_cast<size_t>(n) + 1);
  lag_scaling.push_back(0.0);
  lag_scaling.push_back(1.0);
  double geo = 1.0;
  for (unsigned i = 2; i < nu + 1; ++i) {
    geo *= wscale_update;
    lag_scaling.push_back(lag_scaling.back() + geo);
  }

  std::vector<double> g(K, 0.0), gch(K, 0.0), lp(K);
  Convergence conv{st.W, a.tol};

  auto lagged_update = [&](unsigned k_it, uint32_t s) {      /* :76-100 */
    for (int64_t e = d.rp[s]; e < d.rp[s + 1]; ++e) {
      const int32_t j = d.ci[e];
      const unsigned lagged = k_it - lag[j];
      if (lagged != 0) {
        penalty_apply(a.pen, &st.W[static_cast<size_t>(j) * K], &st.gsum[static_cast<size_t>(j) * K], K, a.gamma,
                      a.beta, wscale, lag_scaling[lagged]);
        lag[j] = k_it;
      }
    }
  };
  auto add_weighted = [&](std::vector<double>& arr, uint32_t s, double scaling) {   /* :114-130 */
    for (int k = 0; k < K; ++k) {
      for (int64_t e = d.rp[s]; e < d.rp[s + 1]; ++e)
        arr[static_cast<size_t>(d.ci[e]) * K + k] += d.cv[e] * gch[k] * scaling;
      if (stdz)
        for (int64_t j = 0; j < p; ++j) arr[static_cast<size_t>(j) * K + k] -= c[j] * gch[k] * scaling;
    }
  };
  auto reset = [&](unsigned k_it) {                          /* :132-155 */
    for (int64_t j = 0; j < p; ++j) {
      const unsigned lagged = k_it - lag[j];
      if (lagged != 0)
        penalty_apply(a.pen, &st.W[static_cast<size_t>(j) * K], &st.gsum[static_cast<size_t>(j) * K], K, a.gamma,
                      a.beta, wscale, lag_scaling[lagged]);
    }
    for (double& w : st.W) w *= wscale;
    return 1.0;
  };

  unsigned it_outer = 0;
  bool converged = false;
  do {
    for (unsigned it = 0; it < nu; ++it) {
      uint32_t s;
      if (!draw_index(rng, nu, &s)) return SGDNET_ERR_RNG;

      lagged_update(it, s);

      for (int k = 0; k < K; ++k)
        lp[k] = dot_sparse(&d.cv[d.rp[s]], &d.ci[d.rp[s]], d.rp[s + 1] - d.rp[s], &st.W[k], K);
      for (int k = 0; k < K; ++k) lp[k] = lp[k] * wscale + st.b[k];
      if (stdz) {
        for (int k = 0; k < K; ++k) {
          double wc = dot_dense(&st.W[k], K, c.data(), p);
          lp[k] -= wc * wscale;
        }
      }

      family_gradient(m, lp.data(), yt.data(), s, g.data());
      for (int k = 0; k < K; ++k) {
        gch[k] = g[k] - st.gmem[static_cast<size_t>(s) * K + k];
        st.gmem[static_cast<size_t>(s) * K + k] = g[k];
      }

      if (wscale < SMALL) {
        wscale = reset(it);
        lag.assign(lag.size(), it);
      }
      wscale *= wscale_update;

      if (m.fit_intercept) {
        for (int k = 0; k < K; ++k) {
          st.gsi[k] += gch[k] / nd;
          st.b[k] -= a.gamma * (st.gsi[k] * 0.01 + gch[k] / nd);
        }
      }

      add_weighted(st.W, s, -a.gamma / wscale);
      lagged_update(it + 1, s);
      add_weighted(st.gsum, s, 1.0 / nd);
    }

    wscale = reset(nu);
    lag.assign(lag.size(), 0u);

    if (a.debug) losses.push_back(epoch_loss(m, d, c, yt, st));
    converged = conv(st.W);
    ++it_outer;
  } while (!converged && it_outer < a.max_iter);

  *epochs_out = it_outer;
  *code_out = (it_outer == a.max_iter) ? 1u : 0u;
  return SGDNET_OK;
}

/* ------------------------------------------------------------------ driver (src/sgdnet.cpp:119-285) */
double* dup(const std::vector<double>& v) {
  double* o = static_cast<double*>(std::malloc(std::max<size_t>(1, v.size()) * sizeof(double)));
  if (!v.empty()) std::memcpy(o, v.data(), v.size() * sizeof(double));
  return o;
}

thread_local bool g_path_only = false;   /* oracle_fit_batch_*: setup up to the lambda path only (sgdnet_fit_spec::path_only) */

int fit_path(Design& d /* raw, samples-major */, std::vector<double> y /* n x Ky column-major */, int Ky,
             const sgdnet_control* ctl, sgdnet_rng* rng, sgdnet_result* out) {
  auto t_begin = std::chrono::steady_clock::now();
  Model m;
  m.family = ctl->family;
  m.K = ctl->n_classes;
  m.Ky = Ky;
  m.fit_intercept = ctl->intercept != 0;
  m.standardize = ctl->standardize != 0;
  m.is_sparse = d.sparse;
  const int K = m.K;
  const int64_t n = d.n, p = d.p;
  const int n_lambda = ctl->n_lambda;
  const double mix = ctl->elasticnet_mix;

  /* -- PreprocessFeatures (src/utils.h:99-121; math.h:66-160) */
  std::vector<double> x_center(p, 0.0), x_scale(p, 1.0);
  if (m.standardize) {
    if (d.sparse) {
      std::vector<double> sum(p, 0.0), var(p, 0.0);
      std::vector<int64_t> cnt(p, 0);
      /* column sums in ascending row order == CSR traversal order per column */
      for (int64_t s = 0; s < n; ++s)
        for (int64_t e = d.rp[s]; e < d.rp[s + 1]; ++e) { sum[d.ci[e]] += d.cv[e]; cnt[d.ci[e]]++; }
      for (int64_t j = 0; j < p; ++j) x_center[j] = sum[j] / static_cast<double>(n);
      for (int64_t s = 0; s < n; ++s)
        for (int64_t e = d.rp[s]; e < d.rp[s + 1]; ++e) {
          int32_t j = d.ci[e];
          var[j] += std::pow(d.cv[e] - x_center[j], 2) / static_cast<double>(n);
        }
      for (int64_t j = 0; j < p; ++j) {
        int64_t n_zeros = n - cnt[j];
        var[j] += static_cast<double>(n_zeros) * x_center[j] * x_center[j] / static_cast<double>(n);
        x_scale[j] = (var[j] == 0.0) ? 1.0 : std::sqrt(var[j]);
      }
      for (size_t e = 0; e < d.cv.size(); ++e) d.cv[e] /= x_scale[d.ci[e]];      /* scale only */
    } else {
      for (int64_t j = 0; j < p; ++j) {
        double s1 = 0.0;
        for (int64_t s = 0; s < n; ++s) s1 += d.dense[s * p + j];
        x_center[j] = s1 / static_cast<double>(n);
        double s2 = 0.0;
        for (int64_t s = 0; s < n; ++s) { double dd = d.dense[s * p + j] - x_center[j]; s2 += dd * dd; }
        double var = s2 / static_cast<double>(n);
        x_scale[j] = (var == 0.0) ? 1.0 : std::sqrt(var);
        for (int64_t s = 0; s < n; ++s) d.dense[s * p + j] = (d.dense[s * p + j] - x_center[j]) / x_scale[j];
      }
    }
  }
  std::vector<double> c(p, 0.0);                               /* x_center_scaled, sgdnet.cpp:150-151 */
  if (d.sparse) for (int64_t j = 0; j < p; ++j) c[j] = x_center[j] / x_scale[j];

  /* -- y in "samples in columns" layout for the family routines */
  auto to_yt = [&](const std::vector<double>& ycm) {
    std::vector<double> yt(static_cast<size_t>(n) * Ky);
    for (int k = 0; k < Ky; ++k)
      for (int64_t i = 0; i < n; ++i) yt[static_cast<size_t>(i) * Ky + k] = ycm[static_cast<size_t>(k) * n + i];
    return yt;
  };

  /* -- null deviance on the original y (sgdnet.cpp:154) */
  const double nulldev = null_deviance(m, to_yt(y).data(), n);

  /* -- family.Preprocess (families.h:69-79, :337-348) */
  std::vector<double> y_center(K, 0.0), y_scale(K, 1.0);
  if (m.family == SGDNET_GAUSSIAN) {
    std::vector<double> mean, sd;
    col_mean(y.data(), n, 1, mean);
    col_sd(y.data(), n, 1, mean, sd);
    y_center[0] = mean[0];
    y_scale[0] = sd[0];
    for (int64_t i = 0; i < n; ++i) y[i] = (y[i] - y_center[0]) / y_scale[0];
  } else if (m.family == SGDNET_MGAUSSIAN && ctl->standardize_response) {
    std::vector<double> mean, sd;
    col_mean(y.data(), n, K, mean);
    col_sd(y.data(), n, K, mean, sd);
    for (int k = 0; k < K; ++k)
      for (int64_t i = 0; i < n; ++i) {
        double& v = y[static_cast<size_t>(k) * n + i];
        v = (v - mean[k]) / sd[k];
      }
  }

  /* -- RegularizationPath (utils.h:142-181) */
  std::vector<double> lambda;
  if (ctl->lambda_len > 0 && ctl->lambda) lambda.assign(ctl->lambda, ctl->lambda + ctl->lambda_len);
  if (lambda.empty()) {
    double lmax = lambda_max(m, d, y, y_scale) / std::max(mix, 0.001);
    if (lmax != 0.0) {                                          /* LogSpace, math.h:42-56 */
      double log_from = std::log(lmax);
      double step = (std::log(lmax * ctl->lambda_min_ratio) - log_from) /
                    static_cast<double>(static_cast<unsigned>(n_lambda) - 1u);
      for (unsigned i = 0; i < static_cast<unsigned>(n_lambda); ++i)
        lambda.push_back(std::exp(log_from + static_cast<double>(i) * step));
    } else {
      lambda.assign(n_lambda, 0.0);
    }
  }
  double max_scale = y_scale[0];
  for (int k = 1; k < K; ++k) max_scale = std::max(max_scale, y_scale[k]);
  std::vector<double> alpha, beta;
  for (double l : lambda) {
    alpha.push_back((1.0 - mix) * l / max_scale);
    beta.push_back(mix * l / max_scale);
  }
  if (static_cast<int>(lambda.size()) < n_lambda) { g_err = "lambda shorter than n_lambda"; return SGDNET_ERR_ARG; }

  const std::vector<double> yt = to_yt(y);

  /* -- ColNormsMax + StepSize (utils.h:31-85) */
  double norm_max = 0.0;
  for (int64_t s = 0; s < n; ++s) {
    double norm = 0.0;
    if (d.sparse) {
      if (m.standardize) {
        int64_t e = d.rp[s];
        for (int64_t j = 0; j < p; ++j) {
          double v = 0.0;
          if (e < d.rp[s + 1] && d.ci[e] == j) v = d.cv[e++];
          double dd = v - c[j];
          norm += dd * dd;
        }
      } else {
        for (int64_t e = d.rp[s]; e < d.rp[s + 1]; ++e) norm += d.cv[e] * d.cv[e];
      }
    } else {
      for (int64_t j = 0; j < p; ++j) norm += d.dense[s * p + j] * d.dense[s * p + j];
    }
    norm_max = std::max(norm_max, norm);
  }
  const double L_scaling = (m.family == SGDNET_BINOMIAL || m.family == SGDNET_MULTINOMIAL) ? 0.25 : 1.0;
  std::vector<double> step_size;
  for (double a_i : alpha) {
    double L = (norm_max + static_cast<double>(m.fit_intercept)) * L_scaling + a_i;
    double mu_n = 2.0 * static_cast<double>(static_cast<unsigned>(n)) * a_i;
    step_size.push_back(1.0 / (2.0 * L + std::min(L, mu_n)));
  }

  /* -- state (sgdnet.cpp:186-211) */
  State st;
  st.W.assign(static_cast<size_t>(p) * K, 0.0);
  st.b.assign(K, 0.0);
  st.gmem.assign(static_cast<size_t>(n) * K, 0.0);
  st.gsum.assign(static_cast<size_t>(p) * K, 0.0);
  st.gsi.assign(K, 0.0);
  {
    std::vector<double> lp0;
    null_predictor(m, yt.data(), n, lp0);
    st.b = lp0;
  }
  const double nulldev_scaled = null_deviance(m, yt.data(), n);

  std::vector<double> a0_arch, beta_arch, dev_ratio, all_losses;
  std::vector<int64_t> losses_ptr{0};
  std::vector<uint32_t> codes, epochs;
  unsigned n_iter = 0;
  double solver_seconds = 0.0, dev_seconds = 0.0;

  const bool group = (m.family == SGDNET_MGAUSSIAN) || (m.family == SGDNET_MULTINOMIAL && ctl->grouped_multinomial);

  for (int li = 0; li < (g_path_only ? 0 : n_lambda); ++li) {
    SagaArgs a;
    a.pen = (mix == 0.0) ? RIDGE : (group ? GROUP : ENET);      /* sgdnet.cpp:80-99 */
    a.gamma = step_size[li];
    a.alpha = alpha[li];
    a.beta = beta[li];
    a.max_iter = ctl->max_iter;
    a.tol = ctl->tol;
    a.debug = ctl->debug != 0;
    std::vector<double> losses;
    unsigned ep = 0, code = 0;
    auto t0 = std::chrono::steady_clock::now();
    int rc = d.sparse ? saga_sparse(m, d, c, yt, st, a, rng, &ep, &code, losses)
                      : saga_dense(m, d, c, yt, st, a, rng, &ep, &code, losses);
    auto t1 = std::chrono::steady_clock::now();
    solver_seconds += std::chrono::duration<double>(t1 - t0).count();
    if (rc != SGDNET_OK) { g_err = "index source exhausted"; return rc; }
    n_iter += ep;
    epochs.push_back(ep);
    codes.push_back(code);

    double dev = deviance(m, d, c, yt, st);
    dev_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
    dev_ratio.push_back(1.0 - dev / nulldev_scaled);

    /* Rescale (utils.h:352-378) on copies */
    std::vector<double> w(st.W), b(st.b), xbs(K, 0.0);
    for (int64_t j = 0; j < p; ++j)
      for (int k = 0; k < K; ++k) {
        double& v = w[static_cast<size_t>(j) * K + k];
        v *= y_scale[k] / x_scale[j];
        xbs[k] += x_center[j] * v;
      }
    if (m.fit_intercept)
      for (int k = 0; k < K; ++k) b[k] = b[k] * y_scale[k] + y_center[k] - xbs[k];
    beta_arch.insert(beta_arch.end(), w.begin(), w.end());
    a0_arch.insert(a0_arch.end(), b.begin(), b.end());

    if (a.debug) all_losses.insert(all_losses.end(), losses.begin(), losses.end());
    losses_ptr.push_back(static_cast<int64_t>(all_losses.size()));
  }

  if (g_path_only) {
    a0_arch.assign(static_cast<size_t>(n_lambda) * K, 0.0);
    beta_arch.assign(static_cast<size_t>(n_lambda) * p * K, 0.0);
    dev_ratio.assign(n_lambda, 0.0);
    codes.assign(n_lambda, 0u);
    epochs.assign(n_lambda, 0u);
    losses_ptr.assign(n_lambda + 1, 0);
  }
  std::memset(out, 0, sizeof(*out));
  out->n_lambda = n_lambda;
  out->n_classes = K;
  out->n_features = p;
  out->a0 = dup(a0_arch);
  out->beta = dup(beta_arch);
  lambda.resize(std::max<size_t>(lambda.size(), n_lambda));
  out->lambda = dup(lambda);
  out->dev_ratio = dup(dev_ratio);
  out->return_codes = static_cast<uint32_t*>(std::malloc(sizeof(uint32_t) * std::max(1, n_lambda)));
  out->epochs = static_cast<uint32_t*>(std::malloc(sizeof(uint32_t) * std::max(1, n_lambda)));
  std::memcpy(out->return_codes, codes.data(), sizeof(uint32_t) * n_lambda);
  std::memcpy(out->epochs, epochs.data(), sizeof(uint32_t) * n_lambda);
  out->losses = dup(all_losses);
  out->losses_ptr = static_cast<int64_t*>(std::malloc(sizeof(int64_t) * (n_lambda + 1)));
  std::memcpy(out->losses_ptr, losses_ptr.data(), sizeof(int64_t) * (n_lambda + 1));
  out->nulldev = nulldev;
  out->npasses = n_iter;
  out->seconds_solver = solver_seconds;
  out->seconds_deviance = dev_seconds;
  out->seconds_total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
  out->seconds_setup = out->seconds_total - solver_seconds - dev_seconds;
  return SGDNET_OK;
}

bool check_args(int64_t n, int64_t p, const void* y, const sgdnet_control* ctl, const sgdnet_rng* rng,
                const sgdnet_result* out) {
  static bool env_read = false;
  if (!env_read) {
    env_read = true;
    if (const char* e = std::getenv("SGDNET_ORACLE_ARITH")) g_arith = (std::string(e) == "libm" || std::string(e) == "0") ? 0 : 1;
  }
  if (n <= 0 || p <= 0 || !y || !ctl || !rng || !out) { g_err = "null or empty argument"; return false; }
  if (ctl->family < 0 || ctl->family > 3) { g_err = "unknown family"; return false; }
  if (ctl->n_lambda <= 0 || ctl->n_classes <= 0) { g_err = "n_lambda and n_classes must be positive"; return false; }
  return true;
}

/* held-out measure: predict (R/predict.sgdnet.R:377, 437, 507-538) + score "deviance" (R/score.R) */
void score_deviance(const Design& d, const double* y /* n x Ky col-major */, int Ky, int family, const double* a0,
                    const double* beta, int n_lambda, int K, double* score) {
  const int64_t n = d.n, p = d.p;
  std::vector<double> lp(K);
  const double pmin = 1e-5, pmax = 1.0 - 1e-5;
  for (int l = 0; l < n_lambda; ++l) {
    const double* B = beta + static_cast<size_t>(l) * p * K;
    const double* A = a0 + static_cast<size_t>(l) * K;
    double acc = 0.0;                       /* gaussian/binomial/multinomial: sum over samples */
    std::vector<double> acc_k(K, 0.0);      /* mgaussian: per-response sums */
    for (int64_t s = 0; s < n; ++s) {
      for (int k = 0; k < K; ++k) lp[k] = A[k];
      if (d.sparse) {
        for (int64_t e = d.rp[s]; e < d.rp[s + 1]; ++e)
          for (int k = 0; k < K; ++k) lp[k] += d.cv[e] * B[static_cast<size_t>(d.ci[e]) * K + k];
      } else {
        for (int64_t j = 0; j < p; ++j)
          for (int k = 0; k < K; ++k) lp[k] += d.dense[s * p + j] * B[static_cast<size_t>(j) * K + k];
      }
      switch (family) {
        case SGDNET_GAUSSIAN: { double r = lp[0] - y[s]; acc += r * r; break; }            /* score.R:66-69 */
        case SGDNET_BINOMIAL: {                                                          /* score.R:103-110 */
          double pr = 1.0 / (1.0 + std::exp(-lp[0]));
          pr = std::min(std::max(pr, pmin), pmax);
          double lpv = (y[s] > 0.5) ? std::log(pr) : std::log(1.0 - pr);
          acc += 2.0 * (0.0 - lpv);
          break;
        }
        case SGDNET_MULTINOMIAL: {                                                       /* score.R:145-151 */
          double tot = 0.0;
          for (int k = 0; k < K; ++k) tot += std::exp(lp[k]);
          unsigned cls = static_cast<unsigned>(y[s] + 0.5);
          double pr = std::exp(lp[cls]) / tot;
          pr = std::min(std::max(pr, pmin), pmax);
          acc += 2.0 * (0.0 - std::log(pr));
          break;
        }
        default:                                                                         /* score.R:175 */
          for (int k = 0; k < K; ++k) { double r = lp[k] - y[static_cast<size_t>(k) * n + s]; acc_k[k] += r * r; }
      }
    }
    if (family == SGDNET_MGAUSSIAN) {
      double tot = 0.0;
      for (int k = 0; k < K; ++k) tot += acc_k[k];
      score[l] = tot / K;                  /* colMeans over responses of colSums over samples */
    } else {
      score[l] = acc / static_cast<double>(n);
    }
  }
}

/* score() for every type.measure (R/score.R:55-178; auc :203-232), on the linear predictors of predict.sgdnet
   (R/predict.sgdnet.R:377, 437, 507-538). y: n x Ky column-major, binomial 0/1, multinomial class ids. */
void score_measure(const Design& d, const double* y, int Ky, int family, int measure, const double* a0, const double* beta,
                   int n_lambda, int K, sgdnet_rng* rng, double* score) {
  if (measure == SGDNET_MEASURE_DEVIANCE) { score_deviance(d, y, Ky, family, a0, beta, n_lambda, K, score); return; }
  const int64_t n = d.n, p = d.p;
  std::vector<double> eta(static_cast<size_t>(n) * K);        /* [s][k] at one lambda */
  std::vector<int> pred_all;                                   /* multinomial class: predicted class per (lambda, sample) */
  if (family == SGDNET_MULTINOMIAL && measure == SGDNET_MEASURE_CLASS) pred_all.resize(static_cast<size_t>(n_lambda) * n);
  for (int l = 0; l < n_lambda; ++l) {
    const double* B = beta + static_cast<size_t>(l) * p * K;
    const double* A = a0 + static_cast<size_t>(l) * K;
    for (int64_t s = 0; s < n; ++s) {
      double* lp = &eta[static_cast<size_t>(s) * K];
      for (int k = 0; k < K; ++k) lp[k] = A[k];
      if (d.sparse) {
        for (int64_t e = d.rp[s]; e < d.rp[s + 1]; ++e)
          for (int k = 0; k < K; ++k) lp[k] += d.cv[e] * B[static_cast<size_t>(d.ci[e]) * K + k];
      } else {
        for (int64_t j = 0; j < p; ++j)
          for (int k = 0; k < K; ++k) lp[k] += d.dense[s * p + j] * B[static_cast<size_t>(j) * K + k];
      }
    }
    double acc = 0.0;
    if (family == SGDNET_GAUSSIAN) {
      for (int64_t s = 0; s < n; ++s) {
        const double r = eta[s] - y[s];
        acc += (measure == SGDNET_MEASURE_MAE) ? std::fabs(r) : r * r;
      }
      score[l] = acc / static_cast<double>(n);
    } else if (family == SGDNET_BINOMIAL) {
      if (measure == SGDNET_MEASURE_AUC) {
        /* auc(): doubled data, ties broken by runif (2n draws), weighted rank sum */
        std::vector<double> r(static_cast<size_t>(2 * n));
        for (double& v : r) v = mt_unif(rng);
        std::vector<int64_t> order(static_cast<size_t>(2 * n));
        for (int64_t i = 0; i < 2 * n; ++i) order[i] = i;
        auto prob = [&](int64_t i) { return 1.0 / (1.0 + std::exp(-eta[i % n])); };
        std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
          const double pa = prob(a), pb = prob(b);
          if (pa != pb) return pa < pb;
          return r[a] < r[b];
        });
        double cw = 0.0, cw1 = 0.0, total = 0.0;
        for (int64_t idx : order) {
          const bool is_one = idx >= n;                                  /* rep(c(0, 1), c(ny, ny)) */
          const double y2 = (y[idx % n] > 0.5) ? 1.0 : 0.0;
          const double w = is_one ? y2 : 1.0 - y2;                        /* as.vector(weights * y) */
          cw += w;
          if (is_one) { cw1 += w; total += w * (cw - cw1); }
        }
        score[l] = std::exp(std::log(total) - std::log(cw1) - std::log(cw - cw1));
        continue;
      }
      for (int64_t s = 0; s < n; ++s) {
        const double pr = 1.0 / (1.0 + std::exp(-eta[s]));
        const double y2 = (y[s] > 0.5) ? 1.0 : 0.0, y1 = 1.0 - y2;
        if (measure == SGDNET_MEASURE_CLASS) acc += y1 * (pr > 0.5 ? 1.0 : 0.0) + y2 * (pr <= 0.5 ? 1.0 : 0.0);
        else {
          const double u = (pr + y1) - 1.0, v = pr - y2;
          acc += (measure == SGDNET_MEASURE_MAE) ? std::fabs(u) + std::fabs(v) : u * u + v * v;
        }
      }
      score[l] = acc / static_cast<double>(n);
    } else if (family == SGDNET_MULTINOMIAL) {
      for (int64_t s = 0; s < n; ++s) {
        const double* lp = &eta[static_cast<size_t>(s) * K];
        double tot = 0.0;
        for (int k = 0; k < K; ++k) tot += std::exp(lp[k]);
        const unsigned cls = static_cast<unsigned>(y[s] + 0.5);
        double best = 0.0;
        int best_k = 0;
        for (int k = 0; k < K; ++k) {
          const double pk = std::exp(lp[k]) / tot;
          const double yk = (static_cast<unsigned>(k) == cls) ? 1.0 : 0.0;
          if (measure == SGDNET_MEASURE_MSE) acc += (yk - pk) * (yk - pk);
          else if (measure == SGDNET_MEASURE_MAE) acc += std::fabs(yk - pk);
          if (k == 0 || pk > best) { best = pk; best_k = k; }
        }
        if (measure == SGDNET_MEASURE_CLASS) pred_all[static_cast<size_t>(l) * n + s] = best_k;
      }
      score[l] = acc / static_cast<double>(n);
    } else {
      for (int64_t s = 0; s < n; ++s)
        for (int k = 0; k < K; ++k) {
          const double r = eta[static_cast<size_t>(s) * K + k] - y[static_cast<size_t>(k) * n + s];
          acc += (measure == SGDNET_MEASURE_MAE) ? std::fabs(r) : r * r;
        }
      score[l] = acc / K;
    }
  }
  if (!pred_all.empty()) {
    /* classid <- as.numeric(as.factor(.)) over ALL samples and lambdas: ranks among the classes predicted anywhere */
    std::vector<int> present(K, 0), code(K, 0);
    for (int c : pred_all) present[c] = 1;
    int rank = 0;
    for (int k = 0; k < K; ++k) { code[k] = rank; rank += present[k]; }
    for (int l = 0; l < n_lambda; ++l) {
      double acc = 0.0;
      for (int64_t s = 0; s < n; ++s) {
        const unsigned cls = static_cast<unsigned>(y[s] + 0.5);
        acc += 1.0 - ((static_cast<unsigned>(code[pred_all[static_cast<size_t>(l) * n + s]]) == cls) ? 1.0 : 0.0);
      }
      score[l] = acc / static_cast<double>(n);
    }
  }
}

void csc_to_design(const int32_t* ci, const int32_t* cp, const double* cx, int64_t n, int64_t p, Design& d) {
  d.sparse = true;
  d.n = n;
  d.p = p;
  const int64_t nnz = cp[p];
  d.rp.assign(n + 1, 0);
  for (int64_t e = 0; e < nnz; ++e) d.rp[ci[e] + 1]++;
  for (int64_t i = 0; i < n; ++i) d.rp[i + 1] += d.rp[i];
  d.ci.resize(nnz);
  d.cv.resize(nnz);
  std::vector<int64_t> fill(d.rp.begin(), d.rp.end() - 1);
  for (int64_t j = 0; j < p; ++j)
    for (int64_t e = cp[j]; e < cp[j + 1]; ++e) {
      int64_t pos = fill[ci[e]]++;
      d.ci[pos] = static_cast<int32_t>(j);
      d.cv[pos] = cx[e];
    }
}

void colmajor_to_design(const double* x, int64_t n, int64_t p, Design& d) {
  d.sparse = false;
  d.n = n;
  d.p = p;
  d.dense.resize(static_cast<size_t>(n) * p);
  for (int64_t j = 0; j < p; ++j)
    for (int64_t i = 0; i < n; ++i) d.dense[i * p + j] = x[j * n + i];
}

}  // namespace

extern "C" {

void oracle_set_arith(int portable) { g_arith = portable ? 1 : 0; }
int oracle_get_arith(void) { return g_arith; }

void oracle_rng_set_seed(sgdnet_rng* rng, uint32_t seed) {
  std::memset(rng, 0, sizeof(*rng));
  rng->kind = SGDNET_RNG_MT;
  mt_set_seed(rng, seed);
}
double oracle_rng_unif(sgdnet_rng* rng) { return mt_unif(rng); }

/* the sampling sequence the solver would see: `count` draws of floor(runif(0, n)) */
int oracle_draw_indices(sgdnet_rng* rng, uint32_t n, int64_t count, uint32_t* out) {
  for (int64_t i = 0; i < count; ++i)
    if (!draw_index(rng, n, &out[i])) return SGDNET_ERR_RNG;
  return SGDNET_OK;
}

const char* oracle_last_error(void) { return g_err.c_str(); }

int oracle_fit_dense(const double* x, int64_t n, int64_t p, const double* y, int32_t y_cols,
                     const sgdnet_control* control, sgdnet_rng* rng, sgdnet_result* out) {
  if (!x || !check_args(n, p, y, control, rng, out)) return SGDNET_ERR_ARG;
  Design d;
  colmajor_to_design(x, n, p, d);
  return fit_path(d, std::vector<double>(y, y + static_cast<size_t>(n) * y_cols), y_cols, control, rng, out);
}

int oracle_fit_sparse(const int32_t* csc_i, const int32_t* csc_p, const double* csc_x, int64_t n, int64_t p,
                      const double* y, int32_t y_cols, const sgdnet_control* control, sgdnet_rng* rng,
                      sgdnet_result* out) {
  if (!csc_i || !csc_p || !csc_x || !check_args(n, p, y, control, rng, out)) return SGDNET_ERR_ARG;
  Design d;
  csc_to_design(csc_i, csc_p, csc_x, n, p, d);
  return fit_path(d, std::vector<double>(y, y + static_cast<size_t>(n) * y_cols), y_cols, control, rng, out);
}

void oracle_result_free(sgdnet_result* r) {
  if (!r) return;
  std::free(r->a0); std::free(r->beta); std::free(r->lambda); std::free(r->dev_ratio);
  std::free(r->return_codes); std::free(r->epochs); std::free(r->losses); std::free(r->losses_ptr);
  std::memset(r, 0, sizeof(*r));
}

int oracle_score_deviance_dense(const double* x, int64_t n, int64_t p, const double* y, int32_t y_cols,
                                int32_t family, const double* a0, const double* beta, int32_t n_lambda,
                                int32_t n_classes, double* score) {
  Design d;
  colmajor_to_design(x, n, p, d);
  score_deviance(d, y, y_cols, family, a0, beta, n_lambda, n_classes, score);
  return SGDNET_OK;
}

int oracle_score_deviance_sparse(const int32_t* csc_i, const int32_t* csc_p, const double* csc_x, int64_t n,
                                 int64_t p, const double* y, int32_t y_cols, int32_t family, const double* a0,
                                 const double* beta, int32_t n_lambda, int32_t n_classes, double* score) {
  Design d;
  csc_to_design(csc_i, csc_p, csc_x, n, p, d);
  score_deviance(d, y, y_cols, family, a0, beta, n_lambda, n_classes, score);
  return SGDNET_OK;
}

int oracle_score_dense(const double* x, int64_t n, int64_t p, const double* y, int32_t y_cols, int32_t family, int32_t measure,
                       const double* a0, const double* beta, int32_t n_lambda, int32_t n_classes, sgdnet_rng* rng, double* score) {
  if (measure == SGDNET_MEASURE_AUC && !rng) { g_err = "auc needs a generator"; return SGDNET_ERR_RNG; }
  Design d;
  colmajor_to_design(x, n, p, d);
  score_measure(d, y, y_cols, family, measure, a0, beta, n_lambda, n_classes, rng, score);
  return SGDNET_OK;
}

int oracle_score_sparse(const int32_t* csc_i, const int32_t* csc_p, const double* csc_x, int64_t n, int64_t p, const double* y,
                        int32_t y_cols, int32_t family, int32_t measure, const double* a0, const double* beta, int32_t n_lambda,
                        int32_t n_classes, sgdnet_rng* rng, double* score) {
  if (measure == SGDNET_MEASURE_AUC && !rng) { g_err = "auc needs a generator"; return SGDNET_ERR_RNG; }
  Design d;
  csc_to_design(csc_i, csc_p, csc_x, n, p, d);
  score_measure(d, y, y_cols, family, measure, a0, beta, n_lambda, n_classes, rng, score);
  return SGDNET_OK;
}

/* The cv_sgdnet double loop (R/cv_sgdnet.R:160-200) one fit after the other: the sequential CPU counterpart of
   sgdnet_fit_batch_* (same specs: row subsets, lambda_from, path_only, held-out deviance). */
static void subset_design(const Design& d, const int32_t* rows, int64_t n_rows, Design& out) {
  out.sparse = d.sparse;
  out.p = d.p;
  out.n = rows ? n_rows : d.n;
  if (!rows) { out = d; return; }
  if (d.sparse) {
    out.rp.assign(1, 0);
    out.ci.clear();
    out.cv.clear();
    for (int64_t i = 0; i < n_rows; ++i) {
      const int64_t r = rows[i];
      out.ci.insert(out.ci.end(), d.ci.begin() + d.rp[r], d.ci.begin() + d.rp[r + 1]);
      out.cv.insert(out.cv.end(), d.cv.begin() + d.rp[r], d.cv.begin() + d.rp[r + 1]);
      out.rp.push_back(static_cast<int64_t>(out.ci.size()));
    }
  } else {
    out.dense.resize(static_cast<size_t>(n_rows) * d.p);
    for (int64_t i = 0; i < n_rows; ++i)
      std::memcpy(&out.dense[static_cast<size_t>(i) * d.p], &d.dense[static_cast<size_t>(rows[i]) * d.p], sizeof(double) * d.p);
  }
}

static std::vector<double> subset_y(const double* y, int64_t n, int Ky, const int32_t* rows, int64_t n_rows) {
  const int64_t m = rows ? n_rows : n;
  std::vector<double> out(static_cast<size_t>(m) * Ky);
  for (int k = 0; k < Ky; ++k)
    for (int64_t i = 0; i < m; ++i) out[static_cast<size_t>(k) * m + i] = y[static_cast<size_t>(k) * n + (rows ? rows[i] : i)];
  return out;
}

static int fit_batch_impl(const Design& raw, const double* y, int32_t y_cols, sgdnet_fit_spec* specs, int32_t n_fits,
                          sgdnet_result* results, double* scores) {
  int max_lambda = 0;
  for (int i = 0; i < n_fits; ++i) max_lambda = std::max(max_lambda, specs[i].control.n_lambda);
  for (int i = 0; i < n_fits; ++i) {
    sgdnet_fit_spec& s = specs[i];
    sgdnet_control ctl = s.control;
    if (s.lambda_from >= 0) {
      if (s.lambda_from >= i) { g_err = "lambda_from must name an earlier fit"; return SGDNET_ERR_ARG; }
      ctl.lambda = results[s.lambda_from].lambda;
      ctl.lambda_len = results[s.lambda_from].n_lambda;
      ctl.n_lambda = ctl.lambda_len;
    }
    Design d;
    subset_design(raw, s.train_rows, s.n_train, d);
    g_path_only = s.path_only != 0;
    const int rc = fit_path(d, subset_y(y, raw.n, y_cols, s.train_rows, s.n_train), y_cols, &ctl, &s.rng, &results[i]);
    g_path_only = false;
    if (rc != SGDNET_OK) return rc;
    if (scores && !s.path_only && s.test_rows && s.n_test > 0) {
      Design dt;
      subset_design(raw, s.test_rows, s.n_test, dt);
      const std::vector<double> yt = subset_y(y, raw.n, y_cols, s.test_rows, s.n_test);
      std::vector<double> sc(results[i].n_lambda);
      score_deviance(dt, yt.data(), y_cols, ctl.family, results[i].a0, results[i].beta, results[i].n_lambda, ctl.n_classes, sc.data());
      std::memcpy(scores + static_cast<size_t>(i) * max_lambda, sc.data(), sizeof(double) * std::min<int>(results[i].n_lambda, max_lambda));
    }
  }
  return SGDNET_OK;
}

int oracle_fit_batch_dense(const double* x, int64_t n, int64_t p, const double* y, int32_t y_cols, sgdnet_fit_spec* specs,
                           int32_t n_fits, sgdnet_result* results, double* scores) {
  if (!x || !y || !specs || !results || n_fits <= 0) { g_err = "null or empty argument"; return SGDNET_ERR_ARG; }
  Design d;
  colmajor_to_design(x, n, p, d);
  return fit_batch_impl(d, y, y_cols, specs, n_fits, results, scores);
}

int oracle_fit_batch_sparse(const int32_t* csc_i, const int32_t* csc_p, const double* csc_x, int64_t n, int64_t p,
                            const double* y, int32_t y_cols, sgdnet_fit_spec* specs, int32_t n_fits,
                            sgdnet_result* results, double* scores) {
  if (!csc_i || !csc_p || !csc_x || !y || !specs || !results || n_fits <= 0) { g_err = "null or empty argument"; return SGDNET_ERR_ARG; }
  Design d;
  csc_to_design(csc_i, csc_p, csc_x, n, p, d);
  return fit_batch_impl(d, y, y_cols, specs, n_fits, results, scores);
}

/* link[l][k][s] = a0[l][k] + x_s . beta[l][:,k]   (R/predict.sgdnet.R:377, 507-510) */
int oracle_predict_dense(const double* x, int64_t n, int64_t p, const double* a0, const double* beta,
                         int32_t n_lambda, int32_t K, double* link) {
  for (int l = 0; l < n_lambda; ++l)
    for (int k = 0; k < K; ++k)
      for (int64_t s = 0; s < n; ++s) {
        double acc = a0[static_cast<size_t>(l) * K + k];
        for (int64_t j = 0; j < p; ++j) acc += x[j * n + s] * beta[(static_cast<size_t>(l) * p + j) * K + k];
        link[(static_cast<size_t>(l) * K + k) * n + s] = acc;
      }
  return SGDNET_OK;
}

int oracle_predict_sparse(const int32_t* csc_i, const int32_t* csc_p, const double* csc_x, int64_t n, int64_t p,
                          const double* a0, const double* beta, int32_t n_lambda, int32_t K, double* link) {
  for (int l = 0; l < n_lambda; ++l)
    for (int k = 0; k < K; ++k) {
      double* o = link + (static_cast<size_t>(l) * K + k) * n;
      for (int64_t s = 0; s < n; ++s) o[s] = a0[static_cast<size_t>(l) * K + k];
      for (int64_t j = 0; j < p; ++j) {
        double bj = beta[(static_cast<size_t>(l) * p + j) * K + k];
        for (int64_t e = csc_p[j]; e < csc_p[j + 1]; ++e) o[csc_i[e]] += csc_x[e] * bj;
      }
    }
  return SGDNET_OK;
}

}  // extern "C"
