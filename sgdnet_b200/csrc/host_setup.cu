// host_setup.cu — see host_setup.h. Host code only (compiled by nvcc's host compiler, no FMA contraction on x86-64).
#include "host_setup.h"

#include <algorithm>
#include <cmath>
#include <atomic>
#include <cstring>
#include <thread>

namespace sgd {

// ------------------------------------------------------------------------------------------ RawX
void RawX::from_dense(const double* x, int64_t n_, int64_t p_) {
  sparse = false;
  n = n_;
  p = static_cast<int32_t>(p_);
  dense_cm = x;
}

// Runs body(k, K) on K host threads (K = 1 runs inline). Used only where the result does not depend on K.
template <typename F>
static void parallel_blocks(int K, F&& body) {
  if (K <= 1) {
    body(0, 1);
    return;
  }
  std::vector<std::thread> th;
  th.reserve(K);
  for (int k = 0; k < K; ++k) th.emplace_back([&body, k, K] { body(k, K); });
  for (auto& t : th) t.join();
}

static int host_threads(int64_t work) {
  if (work < (int64_t(1) << 22)) return 1;
  const unsigned hc = std::thread::hardware_concurrency();
  return static_cast<int>(std::max(1u, std::min(16u, hc)));
}

// CSC -> padded CSR (the reference's AdaptiveTranspose, src/utils.h:276-281). Row counts come from per-thread
// histograms over slices of the entry array; the fill walks, for one block of rows at a time (small enough that its
// output stays cache-resident), every column's ascending row-index run restricted to the block - column ids therefore
// stay ascending inside every row and the result is the same for any thread count.
void RawX::from_csc(const int32_t* ci_, const int32_t* cp_, const double* cx_, int64_t n_, int64_t p_) {
  sparse = true;
  n = n_;
  p = static_cast<int32_t>(p_);
  csc_i = ci_;
  csc_p = cp_;
  csc_x = cx_;
  const int64_t nnz = cp_[p_];
  const int K = host_threads(nnz);
  // ---- counts
  std::vector<std::vector<int32_t>> hist(K);
  parallel_blocks(K, [&](int k, int Kt) {
    std::vector<int32_t>& h = hist[k];
    h.assign(n, 0);
    for (int64_t e = nnz * k / Kt; e < nnz * (k + 1) / Kt; ++e) ++h[ci_[e]];
  });
  rows.resize(n);
  int64_t cursor = 0;
  for (int64_t i = 0; i < n; ++i) {
    int32_t c = 0;
    for (int k = 0; k < K; ++k) c += hist[k][i];
    rows[i].start = cursor;
    rows[i].nnz = c;
    rows[i].pad_ = 0;
    cursor += (c + 3) & ~3;
  }
  hist.clear();
  ci.allocate(static_cast<size_t>(cursor) + 4);
  cv.allocate(static_cast<size_t>(cursor) + 4);
  for (int q = 0; q < 4; ++q) {
    ci[cursor + q] = 0;
    cv[cursor + q] = 0.0;
  }
  // ---- fill, one block of rows at a time
  const int64_t block = 16384;
  const int64_t n_blocks = (n + block - 1) / block;
  std::atomic<int64_t> next{0};
  parallel_blocks(K, [&](int, int) {
    std::vector<int64_t> fill(block);
    for (;;) {
      const int64_t bl = next.fetch_add(1);
      if (bl >= n_blocks) break;
      const int64_t r0 = bl * block, r1 = std::min(n, r0 + block);
      for (int64_t i = r0; i < r1; ++i) fill[i - r0] = rows[i].start;
      for (int64_t j = 0; j < p_; ++j) {          // ascending j => ascending column ids inside every row
        int64_t lo = cp_[j], hi = cp_[j + 1];
        if (n_blocks > 1) {
          lo = std::lower_bound(ci_ + lo, ci_ + hi, static_cast<int32_t>(r0)) - ci_;
          hi = std::lower_bound(ci_ + lo, ci_ + hi, static_cast<int32_t>(r1)) - ci_;
        }
        for (int64_t e = lo; e < hi; ++e) {
          const int64_t dst = fill[ci_[e] - r0]++;
          ci[dst] = static_cast<int32_t>(j);
          cv[dst] = cx_[e];
        }
      }
      for (int64_t i = r0; i < r1; ++i)           // the (at most 3) pad entries of every row
        for (int64_t q = rows[i].start + rows[i].nnz; q < rows[i].start + ((rows[i].nnz + 3) & ~3); ++q) {
          ci[q] = 0;
          cv[q] = 0.0;
        }
    }
  });
}

// ------------------------------------------------------------------------------------------ HostDesign
void HostDesign::build(const RawX& raw, const int32_t* subset, int64_t n_rows, bool standardize) {
  sparse = raw.sparse;
  standardized = standardize;
  n = subset ? n_rows : raw.n;
  p = raw.p;
  ld = (p + 1) & ~1;
  x_center.assign(p, 0.0);
  x_scale.assign(p, 1.0);
  c.assign(p, 0.0);
  const double nd = static_cast<double>(n);
  auto src_row = [&](int64_t i) -> int64_t { return subset ? subset[i] : i; };

  if (!sparse) {
    xd.assign(static_cast<size_t>(n) * ld, 0.0);
    for (int32_t j = 0; j < p; ++j) {
      const double* col = raw.dense_cm + static_cast<size_t>(j) * raw.n;
      for (int64_t i = 0; i < n; ++i) xd[static_cast<size_t>(i) * ld + j] = col[src_row(i)];
    }
    if (standardize) {
      // Mean, StandardDeviation, Standardize (src/math.h:66-79, 114-130, 139-150): centre and scale in place
      for (int32_t j = 0; j < p; ++j) {
        double total = 0.0;
        for (int64_t i = 0; i < n; ++i) total += xd[static_cast<size_t>(i) * ld + j];
        const double mean = total / nd;
        double ss = 0.0;
        for (int64_t i = 0; i < n; ++i) {
          const double dev = xd[static_cast<size_t>(i) * ld + j] - mean;
          ss += dev * dev;
        }
        const double var = ss / nd;
        const double sd = (var == 0.0) ? 1.0 : std::sqrt(var);
        x_center[j] = mean;
        x_scale[j] = sd;
        for (int64_t i = 0; i < n; ++i) {
          double& v = xd[static_cast<size_t>(i) * ld + j];
          v = (v - mean) / sd;
        }
      }
    }
    norm_max = 0.0;
    for (int64_t i = 0; i < n; ++i) {
      double sq = 0.0;
      const double* row = &xd[static_cast<size_t>(i) * ld];
      for (int32_t j = 0; j < p; ++j) sq += row[j] * row[j];
      norm_max = std::max(norm_max, sq);
    }
    max_nnz = p;
    return;
  }

  // sparse: padded CSR of the selected rows
  raw_src = &raw;
  subset_src = subset;
  max_nnz = 0;
  if (!subset && !standardize) {
    // all rows, values as given: the RawX arrays are the design (no copy)
    rows_v = raw.rows.data();
    ci_v = raw.ci.data();
    cv_v = raw.cv.data();
    n_entries = raw.ci.size();
    for (int64_t i = 0; i < n; ++i) max_nnz = std::max(max_nnz, rows_v[i].nnz);
  } else {
    rows.resize(n);
    int64_t cursor = 0;
    for (int64_t i = 0; i < n; ++i) {
      const int64_t r = src_row(i);
      const int32_t nnz = raw.rows[r].nnz;
      rows[i].start = cursor;
      rows[i].nnz = nnz;
      rows[i].pad_ = 0;
      cursor += (nnz + 3) & ~3;
      max_nnz = std::max(max_nnz, nnz);
    }
    ci.assign(static_cast<size_t>(cursor) + 4, 0);
    cv.assign(static_cast<size_t>(cursor) + 4, 0.0);
    parallel_blocks(host_threads(cursor), [&](int k, int Kt) {
      for (int64_t i = n * k / Kt; i < n * (k + 1) / Kt; ++i) {
        const int64_t b = raw.rows[src_row(i)].start;
        std::memcpy(&ci[rows[i].start], &raw.ci[b], sizeof(int32_t) * rows[i].nnz);
        std::memcpy(&cv[rows[i].start], &raw.cv[b], sizeof(double) * rows[i].nnz);
      }
    });
    rows_v = rows.data();
    ci_v = ci.data();
    cv_v = cv.data();
    n_entries = ci.size();
  }
  if (standardize) {
    // sparse Mean / StandardDeviation (src/math.h:66-79, 89-112): per-column running sums in ascending row order;
    // then scale only (src/utils.h:118-120), centring stays virtual through c = center/scale (src/sgdnet.cpp:150)
    std::vector<double> total(p, 0.0), var(p, 0.0);
    std::vector<int64_t> count(p, 0);
    for (int64_t i = 0; i < n; ++i)
      for (int32_t e = 0; e < rows[i].nnz; ++e) {
        const int32_t j = ci[rows[i].start + e];
        total[j] += cv[rows[i].start + e];
        ++count[j];
      }
    for (int32_t j = 0; j < p; ++j) x_center[j] = total[j] / nd;
    for (int64_t i = 0; i < n; ++i)
      for (int32_t e = 0; e < rows[i].nnz; ++e) {
        const int32_t j = ci[rows[i].start + e];
        var[j] += std::pow(cv[rows[i].start + e] - x_center[j], 2) / nd;
      }
    for (int32_t j = 0; j < p; ++j) {
      const int64_t zeros = n - count[j];
      var[j] += static_cast<double>(zeros) * x_center[j] * x_center[j] / nd;
      x_scale[j] = (var[j] == 0.0) ? 1.0 : std::sqrt(var[j]);
    }
    for (int64_t i = 0; i < n; ++i)
      for (int32_t e = 0; e < rows[i].nnz; ++e) cv[rows[i].start + e] /= x_scale[ci[rows[i].start + e]];
    for (int32_t j = 0; j < p; ++j) c[j] = x_center[j] / x_scale[j];
  }
  // ColNormsMax (src/utils.h:60-85): ||x_s - c||^2 over ALL features when centring is virtual
  const int Kn = host_threads(standardize ? int64_t(n) * p : static_cast<int64_t>(n_entries));
  std::vector<double> part(Kn, 0.0);
  parallel_blocks(Kn, [&](int k, int Kt) {     // a max over rows: the same for any split
    double best = 0.0;
    for (int64_t i = n * k / Kt; i < n * (k + 1) / Kt; ++i) {
      double sq = 0.0;
      if (standardize) {
        int32_t e = 0;
        const int64_t b = rows_v[i].start;
        for (int32_t j = 0; j < p; ++j) {
          double v = 0.0;
          if (e < rows_v[i].nnz && ci_v[b + e] == j) v = cv_v[b + e++];
          const double dev = v - c[j];
          sq += dev * dev;
        }
      } else {
        const double* rv = cv_v + rows_v[i].start;
        for (int32_t e = 0; e < rows_v[i].nnz; ++e) sq += rv[e] * rv[e];
      }
      best = std::max(best, sq);
    }
    part[k] = best;
  });
  norm_max = 0.0;
  for (double v : part) norm_max = std::max(norm_max, v);
}

void HostDesign::xt_times(const std::vector<double>& ymap, int m, std::vector<double>& out) const {
  out.assign(static_cast<size_t>(m) * p, 0.0);
  for (int col = 0; col < m; ++col) {
    const double* yc = &ymap[static_cast<size_t>(col) * n];
    double* oc = &out[static_cast<size_t>(col) * p];
    if (sparse && raw_src && raw_src->csc_p) {
      // column by column over the caller's CSC (rows ascending inside a column): the same products added in the
      // same order as the row-major walk below, but columns are independent, so they are spread over host threads
      const RawX& raw = *raw_src;
      std::vector<int32_t> local;               // source row -> position in this view (or -1)
      if (subset_src) {
        local.assign(raw.n, -1);
        for (int64_t i = 0; i < n; ++i) local[subset_src[i]] = static_cast<int32_t>(i);
      }
      const int32_t* loc = subset_src ? local.data() : nullptr;
      parallel_blocks(host_threads(raw.csc_p[p]), [&](int k, int Kt) {
        for (int64_t j = int64_t(p) * k / Kt; j < int64_t(p) * (k + 1) / Kt; ++j) {
          double acc = 0.0;
          const double sc = x_scale[j];
          for (int64_t e = raw.csc_p[j]; e < raw.csc_p[j + 1]; ++e) {
            const int32_t src = raw.csc_i[e];
            const int32_t i = loc ? loc[src] : src;
            if (i < 0) continue;
            const double v = standardized ? raw.csc_x[e] / sc : raw.csc_x[e];
            acc += v * yc[i];
          }
          oc[j] = acc;
        }
      });
    } else if (sparse) {
      for (int64_t i = 0; i < n; ++i) {
        const int64_t b = rows_v[i].start;
        for (int32_t e = 0; e < rows_v[i].nnz; ++e) oc[ci_v[b + e]] += cv_v[b + e] * yc[i];
      }
    } else {
      for (int64_t i = 0; i < n; ++i) {
        const double* row = &xd[static_cast<size_t>(i) * ld];
        for (int32_t j = 0; j < p; ++j) oc[j] += row[j] * yc[i];
      }
    }
  }
}

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
