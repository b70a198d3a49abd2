// host_setup.h — host side of SetupSgdnet (reference src/sgdnet.cpp:119-215): everything that happens once per fit
// before the lambda loop. Phase 1 of the build keeps these O(nnz) passes on the host (SURVEY.md section 8a row 16,
// section 8f rank 1 moves them to the device); the per-sample loop, the per-lambda deviance and the rescale/archive
// all run on the GPU.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "../../include/sgdnet_b200.h"
#include "common.cuh"

namespace sgd {

// A plain buffer that is NOT value-initialised on allocation (zero-filling 1.2 GB on one thread costs more than the
// transpose that fills it); whoever fills it touches the pages first, in parallel.
template <typename T>
struct UninitBuf {
  std::unique_ptr<T[]> p;
  size_t n = 0;
  void allocate(size_t count) {
    p.reset(new T[count]);
    n = count;
  }
  void release() {
    p.reset();
    n = 0;
  }
  T* data() { return p.get(); }
  const T* data() const { return p.get(); }
  size_t size() const { return n; }
  T& operator[](size_t i) { return p[i]; }
  const T& operator[](size_t i) const { return p[i]; }
};

// The caller's matrix, viewed by rows. Dense input is only referenced; sparse input is converted CSC -> CSR once
// (the reference's AdaptiveTranspose, src/utils.h:276-281).
struct RawX {
  bool sparse = false;
  int64_t n = 0;
  int32_t p = 0;
  const double* dense_cm = nullptr;       // n x p column-major (not owned)
  const int32_t *csc_i = nullptr, *csc_p = nullptr;   // the caller's CSC (not owned; valid for the duration of the call)
  const double* csc_x = nullptr;
  std::vector<RowInfo> rows;              // padded CSR: every row starts on a multiple of 4 entries
  UninitBuf<int32_t> ci;
  UninitBuf<double> cv;
  void release_rows() {                   // once every design has been cut from it
    std::vector<RowInfo>().swap(rows);
    ci.release();
    cv.release();
  }

  void from_dense(const double* x, int64_t n_, int64_t p_);
  void from_csc(const int32_t* ci_, const int32_t* cp_, const double* cx_, int64_t n_, int64_t p_);
};

// A (row subset, standardize) view of X prepared for the solver: PreprocessFeatures (src/utils.h:99-121) applied,
// samples contiguous, padded for 16-byte bulk copies.
struct HostDesign {
  bool sparse = false;
  bool standardized = false;
  int64_t n = 0;
  int32_t p = 0, ld = 0;
  std::vector<double> xd;                 // dense [n][ld]
  std::vector<RowInfo> rows;              // sparse, padded CSR (own copy: row subset and/or scaled values) ...
  std::vector<int32_t> ci;
  std::vector<double> cv;
  const RowInfo* rows_v = nullptr;        // ... or a view of the RawX arrays (all rows, unscaled): what is uploaded
  const int32_t* ci_v = nullptr;
  const double* cv_v = nullptr;
  size_t n_entries = 0;                   // length of ci_v / cv_v (padding included)
  std::vector<double> x_center, x_scale, c;   // c = x_center_scaled (zeros for dense)
  double norm_max = 0.0;                  // ColNormsMax (src/utils.h:60-85)
  int32_t max_nnz = 0;
  const RawX* raw_src = nullptr;          // sparse: the matrix this view was cut from, and its row subset (or null);
  const int32_t* subset_src = nullptr;    //   used by xt_times to walk columns in parallel. Valid during setup only.

  void build(const RawX& raw, const int32_t* rows_subset, int64_t n_rows, bool standardize);
  // (X^T * ymap)[c][j] for an n x m column-major ymap, features accumulated in ascending sample order
  void xt_times(const std::vector<double>& ymap, int m, std::vector<double>& out) const;
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
