// host_setup.h — host side of SetupSgdnet (reference src/sgdnet.cpp:119-215). Everything that touches X - the
// transposition, column statistics, scaling, row norms, X^T y - runs on the device (setup.cu); what stays here is
// O(n) work on the response: Preprocess, NullDeviance, FitNullModel (serial sums and libm exp / log whose bits the
// reference returns to R), then the lambda path and the step sizes from the device's scalars.
#pragma once
#include <cstdint>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "../../include/sgdnet_b200.h"
#include "common.cuh"

namespace sgd {

// What the host needs to know about a (row subset, standardize) view of X once the device has prepared it
// (setup.cu: PreprocessFeatures, src/utils.h:99-121; ColNormsMax, src/utils.h:60-85): its shape, the largest squared
// row norm, and a way to ask the device for X^T y (LambdaMax).
struct HostDesign {
  bool sparse = false;
  bool standardized = false;
  int64_t n = 0;
  int32_t p = 0, ld = 0;
  double norm_max = 0.0;                  // ColNormsMax
  // (X^T * ymap)[c][j] for an n x m column-major ymap, every feature accumulated in ascending sample order
  std::function<void(const std::vector<double>& ymap, int m, std::vector<double>& out)> xt_times_fn;
  void xt_times(const std::vector<double>& ymap, int m, std::vector<double>& out) const { xt_times_fn(ymap, m, out); }
};

// Everything y- and control-dependent (families.h Preprocess / NullDeviance / FitNullModel / LambdaMax;
// utils.h RegularizationPath / StepSize).
struct FitPlan {
  int family = 0, K = 1, Ky = 1, penalty = 0, n_lambda = 0;
  bool fit_intercept = true, standardize = true;
  uint32_t max_iter = 0;
  double tol = 0.0;
  bool debug = false;
  std::vector<double> yt;                 // [n][Ky] preprocessed
  std::vector<double> y_center, y_scale;
  std::vector<double> lambda, alpha, beta, gamma;
  std::vector<double> intercept0;         // FitNullModel
  double nulldev = 0.0, nulldev_scaled = 0.0;

  // y_rows: n x Ky column-major response restricted to the fit's rows
  std::string build(const HostDesign& d, std::vector<double> y_cm, int Ky_, const sgdnet_control& ctl);
};

// R-compatible Mersenne-Twister and the sample-index stream (R core RNG.c; call sites src/saga-dense.h:152,
// src/saga-sparse.h:261).
void mt_seed(sgdnet_rng* r, uint32_t seed);
double mt_unif(sgdnet_rng* r);
// Appends `count` indices floor(runif(0, n)); returns false when the source cannot supply them.
bool draw_indices(sgdnet_rng* r, uint32_t n, int64_t count, uint32_t* out);

}  // namespace sgd
