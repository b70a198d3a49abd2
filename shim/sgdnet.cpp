// shim/sgdnet.cpp — the drop-in replacement for the reference's src/sgdnet.cpp.
//
// A maintainer copies this file over src/sgdnet.cpp (and may delete saga-dense.h, saga-sparse.h, penalties.h, prox.h,
// families.h, utils.h, math.h, constants.h). Nothing else in the R package changes: the two exported functions keep
// their names and their exact C++ signatures (reference src/sgdnet.cpp:359-375), so the generated glue
// src/RcppExports.cpp / R/RcppExports.R and every R caller (R/sgdnet.R:362-366, cv_sgdnet, predict(exact = TRUE))
// stay as they are. Instead of running SetupFamily -> SetupSgdnet -> RunSaga -> Saga on the CPU (src/sgdnet.cpp:71-335)
// the bodies hand plain pointers to the extern "C" entry points of include/sgdnet_b200.h (libsgdnet_b200.so: the
// sm_100a CUDA backend) and wrap the result in the list of src/sgdnet.cpp:275-284 - same names, same element types
// (a0: list of K-vectors, beta: list of K x p arrays, losses: list of numeric vectors, npasses, nulldev, dev.ratio,
// lambda, return_codes).
//
// Build (src/Makevars):  PKG_CPPFLAGS = -I$(SGDNET_B200_HOME)/include
//                        PKG_LIBS = -L$(SGDNET_B200_HOME)/lib -lsgdnet_b200 -Wl,-rpath,$(SGDNET_B200_HOME)/lib
//
// Sampling. The reference draws floor(R::runif(0, n)) from R's global generator once per update, inside the
// RNGScope opened by the generated wrapper (src/RcppExports.cpp:14,27). The shim passes unif_rand() as the library's
// SGDNET_RNG_CALLBACK source: called on this thread, once per draw, in order, exactly n * npasses times, so
// .Random.seed ends where the reference leaves it.
//
// This file compiles against real Rcpp/RcppEigen and, for the tests in this repository (no R in the image), against
// the stand-in oracle/refbuild/standin/RcppEigen.h: tests/test_shim_gpu.py calls these two functions with Rcpp::List
// controls and R's Mersenne-Twister and compares the lists they return with the reference's own SgdnetDense /
// SgdnetSparse (oracle/_ref).
#include <RcppEigen.h>

#include <string>
#include <vector>

#include "sgdnet_b200.h"

namespace sgdnet_shim {

// the control list of R/sgdnet.R:346-359 -> sgdnet_control (keys read at src/sgdnet.cpp:76-78, 129-138)
inline sgdnet_control unpack_control(const Rcpp::List& control, std::vector<double>& lambda_keep) {
  sgdnet_control c{};
  const std::string family = Rcpp::as<std::string>(control["family"]);
  if (family == "gaussian") c.family = SGDNET_GAUSSIAN;
  else if (family == "binomial") c.family = SGDNET_BINOMIAL;
  else if (family == "multinomial") c.family = SGDNET_MULTINOMIAL;
  else if (family == "mgaussian") c.family = SGDNET_MGAUSSIAN;
  else c.family = -1;                        // the reference returns an empty list (src/sgdnet.cpp:333-334)
  c.intercept = Rcpp::as<bool>(control["intercept"]);
  c.standardize = Rcpp::as<bool>(control["standardize"]);
  c.standardize_response = Rcpp::as<bool>(control["standardize_response"]);
  c.n_lambda = static_cast<int32_t>(Rcpp::as<unsigned>(control["n_lambda"]));
  c.n_classes = static_cast<int32_t>(Rcpp::as<unsigned>(control["n_classes"]));
  c.debug = Rcpp::as<bool>(control["debug"]);
  c.grouped_multinomial = Rcpp::as<std::string>(control["type_multinomial"]) == "grouped";
  c.max_iter = Rcpp::as<unsigned>(control["max_iter"]);
  c.elasticnet_mix = Rcpp::as<double>(control["elasticnet_mix"]);
  c.lambda_min_ratio = Rcpp::as<double>(control["lambda_min_ratio"]);
  c.tol = Rcpp::as<double>(control["tol"]);
  lambda_keep = Rcpp::as<std::vector<double>>(control["lambda"]);
  c.lambda_len = static_cast<int32_t>(lambda_keep.size());
  c.lambda = lambda_keep.empty() ? nullptr : lambda_keep.data();
  return c;
}

inline double r_unif_rand(void*) { return unif_rand(); }   // R API, R_ext/Random.h

inline sgdnet_rng r_global_generator() {
  sgdnet_rng rng{};
  rng.kind = SGDNET_RNG_CALLBACK;
  rng.unif_rand = r_unif_rand;
  return rng;
}

// sgdnet_result -> the list of src/sgdnet.cpp:275-284
inline Rcpp::List wrap_result(sgdnet_result& r) {
  const int L = r.n_lambda, K = r.n_classes;
  const long p = static_cast<long>(r.n_features);
  std::vector<Eigen::ArrayXd> a0;
  std::vector<Eigen::ArrayXXd> beta;
  std::vector<std::vector<double>> losses;
  for (int l = 0; l < L; ++l) {
    Eigen::ArrayXd a(K);
    for (int k = 0; k < K; ++k) a(k) = r.a0[static_cast<size_t>(l) * K + k];
    a0.push_back(a);
    Eigen::ArrayXXd b(K, p);                                   // K x p column-major, as the reference's weights
    const double* src = r.beta + static_cast<size_t>(l) * p * K;
    for (long j = 0; j < p; ++j)
      for (int k = 0; k < K; ++k) b(k, j) = src[static_cast<size_t>(j) * K + k];
    beta.push_back(b);
  }
  if (r.losses_ptr != nullptr && r.losses_ptr[L] > 0)          // the reference archives losses only in debug mode
    for (int l = 0; l < L; ++l) losses.emplace_back(r.losses + r.losses_ptr[l], r.losses + r.losses_ptr[l + 1]);
  const std::vector<double> dev(r.dev_ratio, r.dev_ratio + L), lambda(r.lambda, r.lambda + L);
  const std::vector<unsigned> codes(r.return_codes, r.return_codes + L);
  const unsigned npasses = r.npasses;
  const double nulldev = r.nulldev;
  sgdnet_result_free(&r);
  return Rcpp::List::create(
      Rcpp::Named("a0") = Rcpp::wrap(a0),
      Rcpp::Named("beta") = Rcpp::wrap(beta),
      Rcpp::Named("losses") = Rcpp::wrap(losses),
      Rcpp::Named("npasses") = npasses,
      Rcpp::Named("nulldev") = nulldev,
      Rcpp::Named("dev.ratio") = Rcpp::wrap(dev),
      Rcpp::Named("lambda") = Rcpp::wrap(lambda),
      Rcpp::Named("return_codes") = Rcpp::wrap(codes));
}

inline void check(int status) {
  if (status != SGDNET_OK) Rcpp::stop(std::string("sgdnet_b200: ") + sgdnet_last_error());   // no CPU fallback
}

}  // namespace sgdnet_shim

// [[Rcpp::export]]
Rcpp::List
SgdnetDense(const Eigen::MatrixXd& x,
            const Eigen::MatrixXd& y,
            const Rcpp::List&      control)
{
  std::vector<double> lambda;
  sgdnet_control c = sgdnet_shim::unpack_control(control, lambda);
  if (c.family < 0) return Rcpp::List::create();
  sgdnet_rng rng = sgdnet_shim::r_global_generator();
  sgdnet_result res{};
  sgdnet_shim::check(sgdnet_fit_dense(x.data(), x.rows(), x.cols(), y.data(), static_cast<int32_t>(y.cols()), &c, &rng, &res));
  return sgdnet_shim::wrap_result(res);
}

// [[Rcpp::export]]
Rcpp::List
SgdnetSparse(const Eigen::SparseMatrix<double>& x,
             const Eigen::MatrixXd&             y,
             const Rcpp::List&                  control)
{
  std::vector<double> lambda;
  sgdnet_control c = sgdnet_shim::unpack_control(control, lambda);
  if (c.family < 0) return Rcpp::List::create();
  sgdnet_rng rng = sgdnet_shim::r_global_generator();
  sgdnet_result res{};
  if (x.isCompressed()) {      // a dgCMatrix mapped by RcppEigen: column pointers, row ids, values - CSC, 0-based, int
    sgdnet_shim::check(sgdnet_fit_sparse(x.innerIndexPtr(), x.outerIndexPtr(), x.valuePtr(), x.rows(), x.cols(), y.data(),
                                         static_cast<int32_t>(y.cols()), &c, &rng, &res));
  } else {
    Eigen::SparseMatrix<double> xc(x);
    xc.makeCompressed();
    sgdnet_shim::check(sgdnet_fit_sparse(xc.innerIndexPtr(), xc.outerIndexPtr(), xc.valuePtr(), xc.rows(), xc.cols(), y.data(),
                                         static_cast<int32_t>(y.cols()), &c, &rng, &res));
  }
  return sgdnet_shim::wrap_result(res);
}
