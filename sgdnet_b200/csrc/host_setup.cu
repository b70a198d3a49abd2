// host_setup.cu — see host_setup.h. Host code only (compiled by nvcc's host compiler, no FMA contraction on x86-64).
#include "host_setup.h"

#include <algorithm>
#include <cmath>
#include <atomic>
#include <cstring>
#include <thread>

namespace sgd {

// ------------------------------------------------------------------------------------------ FitPlan
namespace {

void column_stats(const std::vector<double>& a, int64_t n, int m, std::vector<double>& mean, std::vector<double>& sd) {
  mean.assign(m, 0.0);
  sd.assign(m, 1.0);
  const double nd = static_cast<double>(n);
  for (int j = 0; j < m; ++j) {
    const double* col = &a[static_cast<size_t>(j) * n];
    double total = 0.0;
    for (int64_t i = 0; i < n; ++i) total += col[i];
    mean[j] = total / nd;
    double ss = 0.0;
    for (int64_t i = 0; i < n; ++i) {
      const double dev = col[i] - mean[j];
      ss += dev * dev;
    }
    const double var = ss / nd;
    sd[j] = (var == 0.0) ? 1.0 : std::sqrt(var);
  }
}

double lse(const double* x, int K) {
  double mx = x[0];
  for (int k = 1; k < K; ++k) mx = std::max(mx, x[k]);
  double total = 0.0;
  for (int k = 0; k < K; ++k) total += std::exp(x[k] - mx);
  return std::log(total) + mx;
}

double host_loss(int family, int K, int Ky, const double* lp, const double* yt, int64_t s) {
  switch (family) {
    case kGaussian: return 0.5 * (lp[0] - yt[s]) * (lp[0] - yt[s]);
    case kBinomial: return std::log(1.0 + std::exp(lp[0])) - yt[s] * lp[0];
    case kMultinomial: return lse(lp, K) - lp[static_cast<unsigned>(yt[s] + 0.5)];
    default: {
      double total = 0.0;
      for (int k = 0; k < K; ++k) {
        const double d = lp[k] - yt[s * Ky + k];
        total += d * d;
      }
      return 0.5 * total;
    }
  }
}

// FitNullModel / the predictor NullDeviance evaluates (families.h:98-117, 170-201, 262-298, 367-385)
std::vector<double> null_model(int family, int K, int Ky, bool fit_intercept, const std::vector<double>& yt, int64_t n) {
  std::vector<double> lp(K, 0.0);
  const double nd = static_cast<double>(n);
  if (family == kGaussian || family == kMGaussian) {
    for (int k = 0; k < K; ++k) {
      double total = 0.0;
      for (int64_t i = 0; i < n; ++i) total += yt[i * Ky + k];
      lp[k] = total / nd;
    }
  } else if (family == kBinomial) {
    if (fit_intercept) {
      double total = 0.0;
      for (int64_t i = 0; i < n; ++i) total += yt[i];
      const double ybar = total / nd;
      const double lo = 1e-9, hi = 1.0 - 1e-9;
      const double z = ybar > hi ? hi : (ybar < lo ? lo : ybar);
      lp[0] = std::log(z / (1.0 - z));
    }
  } else {
    std::vector<double> prop(K, fit_intercept ? 0.0 : 1.0 / K);
    if (fit_intercept)
      for (int64_t i = 0; i < n; ++i) prop[static_cast<int64_t>(yt[i] + 0.5)] += 1.0 / nd;
    double log_total = 0.0;
    for (int k = 0; k < K; ++k) log_total += std::log(prop[k]);
    for (int k = 0; k < K; ++k) lp[k] = std::log(prop[k]) - log_total / K;
  }
  return lp;
}

double null_dev(int family, int K, int Ky, bool fit_intercept, const std::vector<double>& yt, int64_t n) {
  const std::vector<double> lp = null_model(family, K, Ky, fit_intercept, yt, n);
  double total = 0.0;
  if (family == kMultinomial) {
    const double l = lse(lp.data(), K);
    for (int64_t i = 0; i < n; ++i) total += l - lp[static_cast<unsigned>(yt[i] + 0.5)];
  } else {
    for (int64_t i = 0; i < n; ++i) total += host_loss(family, K, Ky, lp.data(), yt.data(), i);
  }
  return 2.0 * total;
}

std::vector<double> sample_major(const std::vector<double>& y_cm, int64_t n, int Ky) {
  std::vector<double> yt(static_cast<size_t>(n) * Ky);
  for (int k = 0; k < Ky; ++k)
    for (int64_t i = 0; i < n; ++i) yt[static_cast<size_t>(i) * Ky + k] = y_cm[static_cast<size_t>(k) * n + i];
  return yt;
}

}  // namespace

std::string FitPlan::build(const HostDesign& d, std::vector<double> y, int Ky_, const sgdnet_control& ctl) {
  family = ctl.family;
  K = ctl.n_classes;
  Ky = Ky_;
  n_lambda = ctl.n_lambda;
  fit_intercept = ctl.intercept != 0;
  standardize = ctl.standardize != 0;
  max_iter = ctl.max_iter;
  tol = ctl.tol;
  debug = ctl.debug != 0;
  const int64_t n = d.n;
  const double nd = static_cast<double>(n);
  const double mix = ctl.elasticnet_mix;
  if (family < 0 || family > 3) return "unknown family";
  if (K < 1 || K > 32) return "n_classes must be in [1, 32] in this build";
  if ((family == kGaussian || family == kBinomial) && K != 1) return "n_classes must be 1 for gaussian/binomial";
  if (family == kMGaussian && Ky != K) return "mgaussian needs one response column per class";
  if (family != kMGaussian && Ky != 1) return "response must have one column";
  if (family == kMultinomial)
    for (int64_t i = 0; i < n; ++i)
      if (!(y[i] + 0.5 >= 0.0) || static_cast<int64_t>(y[i] + 0.5) >= K) return "class id out of range";
  const bool group = (family == kMGaussian) || (family == kMultinomial && ctl.grouped_multinomial);
  penalty = (mix == 0.0) ? kRidge : (group ? kGroupLasso : kElasticNet);      // src/sgdnet.cpp:80-99

  nulldev = null_dev(family, K, Ky, fit_intercept, sample_major(y, n, Ky), n);   // original y, src/sgdnet.cpp:154

  // family.Preprocess (families.h:69-79, 337-348)
  y_center.assign(K, 0.0);
  y_scale.assign(K, 1.0);
  if (family == kGaussian) {
    std::vector<double> mean, sd;
    column_stats(y, n, 1, mean, sd);
    y_center[0] = mean[0];
    y_scale[0] = sd[0];
    for (int64_t i = 0; i < n; ++i) y[i] = (y[i] - mean[0]) / sd[0];
  } else if (family == kMGaussian && ctl.standardize_response) {
    std::vector<double> mean, sd;
    column_stats(y, n, K, mean, sd);
    for (int k = 0; k < K; ++k)
      for (int64_t i = 0; i < n; ++i) {
        double& v = y[static_cast<size_t>(k) * n + i];
        v = (v - mean[k]) / sd[k];
      }
  }

  // RegularizationPath (utils.h:142-181) with LambdaMax (families.h:119-126, 203-220, 300-325, 387-406)
  lambda.clear();
  if (ctl.lambda_len > 0 && ctl.lambda) lambda.assign(ctl.lambda, ctl.lambda + ctl.lambda_len);
  if (lambda.empty()) {
    std::vector<double> ip, mean, sd;
    double lmax = 0.0;
    if (family == kGaussian) {
      d.xt_times(y, 1, ip);
      double mx = 0.0;
      for (int32_t j = 0; j < d.p; ++j) mx = std::max(mx, std::fabs(ip[j]));
      lmax = y_scale[0] * mx / nd;
    } else if (family == kBinomial) {
      column_stats(y, n, 1, mean, sd);
      std::vector<double> ymap(n);
      for (int64_t i = 0; i < n; ++i) ymap[i] = (y[i] - mean[0]) / sd[0];
      d.xt_times(ymap, 1, ip);
      double mx = 0.0;
      for (int32_t j = 0; j < d.p; ++j) mx = std::max(mx, std::fabs(ip[j]));
      lmax = sd[0] * mx / nd;
    } else if (family == kMultinomial) {
      std::vector<double> ymap(static_cast<size_t>(n) * K, 0.0);
      for (int64_t i = 0; i < n; ++i) ymap[static_cast<size_t>(static_cast<unsigned>(y[i] + 0.5)) * n + i] = 1.0;
      column_stats(ymap, n, K, mean, sd);
      for (int k = 0; k < K; ++k)
        for (int64_t i = 0; i < n; ++i) {
          double& v = ymap[static_cast<size_t>(k) * n + i];
          v = (v - mean[k]) / sd[k];
        }
      d.xt_times(ymap, K, ip);
      double mx = 0.0;
      for (int k = 0; k < K; ++k)
        for (int32_t j = 0; j < d.p; ++j) mx = std::max(mx, std::fabs(ip[static_cast<size_t>(k) * d.p + j] * sd[k]));
      lmax = mx / nd;
    } else {
      std::vector<double> ymap(y);
      column_stats(y, n, K, mean, sd);
      for (int k = 0; k < K; ++k)
        for (int64_t i = 0; i < n; ++i) {
          double& v = ymap[static_cast<size_t>(k) * n + i];
          v = (v - mean[k]) / sd[k];
        }
      d.xt_times(ymap, K, ip);
      double mx = 0.0;
      for (int32_t j = 0; j < d.p; ++j) {
        double sq = 0.0;
        for (int k = 0; k < K; ++k) {
          const double v = ip[static_cast<size_t>(k) * d.p + j] * (y_scale[k] * sd[k]);
          sq += v * v;
        }
        mx = std::max(mx, std::sqrt(sq));
      }
      lmax = mx / nd;
    }
    lmax = lmax / std::max(mix, 0.001);
    if (lmax != 0.0) {                       // LogSpace (math.h:42-56)
      const double log_from = std::log(lmax);
      const double step = (std::log(lmax * ctl.lambda_min_ratio) - log_from) /
                          static_cast<double>(static_cast<unsigned>(n_lambda) - 1u);
      for (unsigned i = 0; i < static_cast<unsigned>(n_lambda); ++i)
        lambda.push_back(std::exp(log_from + static_cast<double>(i) * step));
    } else {
      lambda.assign(n_lambda, 0.0);
    }
  }
  if (static_cast<int>(lambda.size()) < n_lambda) return "lambda has fewer than n_lambda values";
  lambda.resize(n_lambda);     // R sends nlambda = length(lambda) (R/sgdnet.R:242-245); extra values are never fitted
  double max_scale = y_scale[0];
  for (int k = 1; k < K; ++k) max_scale = std::max(max_scale, y_scale[k]);
  alpha.clear();
  beta.clear();
  gamma.clear();
  for (double l : lambda) {
    alpha.push_back((1.0 - mix) * l / max_scale);
    beta.push_back(mix * l / max_scale);
  }
  // StepSize (utils.h:31-51)
  const double L_scaling = (family == kBinomial || family == kMultinomial) ? 0.25 : 1.0;
  for (double a : alpha) {
    const double L = (d.norm_max + static_cast<double>(fit_intercept)) * L_scaling + a;
    const double mu_n = 2.0 * static_cast<double>(static_cast<unsigned>(n)) * a;
    gamma.push_back(1.0 / (2.0 * L + std::min(L, mu_n)));
  }

  yt = sample_major(y, n, Ky);
  intercept0 = null_model(family, K, Ky, fit_intercept, yt, n);
  nulldev_scaled = null_dev(family, K, Ky, fit_intercept, yt, n);
  return "";
}

// ------------------------------------------------------------------------------------------ RNG
void mt_seed(sgdnet_rng* r, uint32_t seed) {
  // set.seed(): Randomize() scrambles the seed 50 times, RNG_Init fills dummy[0..624] with the same LCG, FixupSeeds
  // sets mti = N so that the first draw regenerates the whole block.
  for (int i = 0; i < 50; ++i) seed = 69069u * seed + 1u;
  for (int i = 0; i < 625; ++i) {
    seed = 69069u * seed + 1u;
    if (i > 0) r->mt[i - 1] = seed;
  }
  r->mti = 624;
}

static inline uint32_t mt_word(sgdnet_rng* r) {
  constexpr int N = 624, M = 397;
  constexpr uint32_t kMatrixA = 0x9908b0dfu, kUpper = 0x80000000u, kLower = 0x7fffffffu;
  uint32_t* mt = r->mt;
  if (r->mti >= N) {
    int kk = 0;
    for (; kk < N - M; ++kk) {
      const uint32_t y = (mt[kk] & kUpper) | (mt[kk + 1] & kLower);
      mt[kk] = mt[kk + M] ^ (y >> 1) ^ ((y & 1u) ? kMatrixA : 0u);
    }
    for (; kk < N - 1; ++kk) {
      const uint32_t y = (mt[kk] & kUpper) | (mt[kk + 1] & kLower);
      mt[kk] = mt[kk + (M - N)] ^ (y >> 1) ^ ((y & 1u) ? kMatrixA : 0u);
    }
    const uint32_t y = (mt[N - 1] & kUpper) | (mt[0] & kLower);
    mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ ((y & 1u) ? kMatrixA : 0u);
    r->mti = 0;
  }
  uint32_t y = mt[r->mti++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

double mt_unif(sgdnet_rng* r) {
  const double u = mt_word(r) * 2.3283064365386963e-10;
  constexpr double half_step = 0.5 * 2.328306437080797e-10;   // fixup(): keep the value inside (0, 1)
  if (u <= 0.0) return half_step;
  if (1.0 - u <= 0.0) return 1.0 - half_step;
  return u;
}

bool draw_indices(sgdnet_rng* r, uint32_t n, int64_t count, uint32_t* out) {
  const double nd = static_cast<double>(n);
  switch (r->kind) {
    case SGDNET_RNG_MT:
      for (int64_t i = 0; i < count; ++i) out[i] = static_cast<uint32_t>(std::floor(0.0 + (nd - 0.0) * mt_unif(r)));
      return true;
    case SGDNET_RNG_CALLBACK:
      if (!r->unif_rand) return false;
      for (int64_t i = 0; i < count; ++i)
        out[i] = static_cast<uint32_t>(std::floor(0.0 + (nd - 0.0) * r->unif_rand(r->ctx)));
      return true;
    case SGDNET_RNG_SEQUENCE:
      if (!r->seq || r->seq_pos + count > r->seq_len) return false;
      std::memcpy(out, r->seq + r->seq_pos, sizeof(uint32_t) * count);
      r->seq_pos += count;
      return true;
  }
  return false;
}

}  // namespace sgd
