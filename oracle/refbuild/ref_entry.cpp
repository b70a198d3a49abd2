/*
 * ref_entry.cpp — C entry points around the REFERENCE'S OWN translation unit.  TEST INFRASTRUCTURE ONLY.
 *
 * `#include "sgdnet.cpp"` below pulls in /root/reference/src/sgdnet.cpp (and through it saga-dense.h, saga-sparse.h,
 * penalties.h, prox.h, families.h, utils.h, math.h, constants.h) unmodified, from where it lies (the Makefile passes
 * -I$(REF_SRC)); nothing of the reference is copied into this repository. The two external libraries it includes are
 * replaced by oracle/refbuild/standin/RcppEigen.h (see that header for what the stand-in does and does not pin).
 * The result, oracle/_ref/libsgdnet_ref.so, exports the oracle's C ABI under the prefix `ref_`:
 *   ref_fit_dense / ref_fit_sparse  ->  SgdnetDense / SgdnetSparse (reference src/sgdnet.cpp:359-375), called with a
 *                                       `control` list holding exactly the keys R/sgdnet.R:346-359 builds
 * Used by tests/test_ref_cpu.py to check oracle/sgdnet_oracle.cpp against the reference's code, and by nothing else.
 *
 * Sampling: the reference calls R::runif(0, n) (src/saga-dense.h:152, src/saga-sparse.h:261). R is not here, so
 * R's unif_rand() for the default Mersenne-Twister is restated below (R core src/main/RNG.c: MT_genrand + fixup;
 * set.seed scrambling: 50 LCG steps, then 625 LCG words, mti = 624) over the caller's sgdnet_rng state.
 */
#include <RcppEigen.h>

#include <chrono>
#include <cstdlib>
#include <cstring>

#include "../../include/sgdnet_b200.h"

namespace refentry {

thread_local sgdnet_rng* g_rng = nullptr;
thread_local int64_t g_n_for_sequence = 0;
thread_local std::string g_err;

inline uint32_t mt_next(sgdnet_rng* r) {
  const int N = 624, M = 397;
  const uint32_t kMatrixA = 0x9908b0dfu, kUpper = 0x80000000u, kLower = 0x7fffffffu;
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

inline double fixup(double x) {
  const double i2_32m1 = 2.328306437080797e-10;
  if (x <= 0.0) return 0.5 * i2_32m1;
  if ((1.0 - x) <= 0.0) return 1.0 - 0.5 * i2_32m1;
  return x;
}

inline double unif(sgdnet_rng* r) {
  if (r->kind == SGDNET_RNG_MT) return fixup(static_cast<double>(mt_next(r)) * 2.3283064365386963e-10);
  if (r->kind == SGDNET_RNG_CALLBACK) {
    if (!r->unif_rand) throw std::runtime_error("rng callback missing");
    return r->unif_rand(r->ctx);
  }
  if (r->seq_pos >= r->seq_len) throw std::runtime_error("index sequence exhausted");
  /* explicit indices (testing): hand back a draw whose floor(n*u) is the given index */
  return (static_cast<double>(r->seq[r->seq_pos++]) + 0.5) / static_cast<double>(g_n_for_sequence);
}

}  // namespace refentry

double R::unif_rand_hook() {
  if (!refentry::g_rng) throw std::runtime_error("no generator installed");
  return refentry::unif(refentry::g_rng);
}

/* ------------------------------------------------------------------ the reference, unmodified
   (-I$(REF_SRC) resolves this to /root/reference/src/sgdnet.cpp; the shim test build puts -I../../shim first, so the
   same entry points then wrap shim/sgdnet.cpp, the drop-in replacement that calls libsgdnet_b200.so) */
#include "sgdnet.cpp"

#ifndef REF_SYM
#define REF_SYM(name) ref_##name
#endif

/* ------------------------------------------------------------------ C ABI */
namespace refentry {

int g_force_debug = 1;   /* the reference returns per-lambda epoch counts only through `losses` (debug) */

Rcpp::List make_control(const sgdnet_control* c, bool is_sparse) {
  static const char* fam[] = {"gaussian", "binomial", "multinomial", "mgaussian"};
  if (c->family < 0 || c->family > 3) throw std::runtime_error("unknown family");
  Rcpp::List l;
  l.set("debug", static_cast<bool>(c->debug != 0 || g_force_debug != 0));
  l.set("elasticnet_mix", static_cast<double>(c->elasticnet_mix));
  l.set("family", std::string(fam[c->family]));
  l.set("intercept", static_cast<bool>(c->intercept != 0));
  l.set("is_sparse", is_sparse);
  l.set("lambda", c->lambda_len > 0 ? std::vector<double>(c->lambda, c->lambda + c->lambda_len) : std::vector<double>());
  l.set("lambda_min_ratio", static_cast<double>(c->lambda_min_ratio));
  l.set("max_iter", static_cast<unsigned>(c->max_iter));
  l.set("n_lambda", static_cast<unsigned>(c->lambda_len > 0 ? c->lambda_len : c->n_lambda));
  l.set("n_classes", static_cast<unsigned>(c->n_classes));
  l.set("standardize", static_cast<bool>(c->standardize != 0));
  l.set("standardize_response", static_cast<bool>(c->standardize_response != 0));
  l.set("tol", static_cast<double>(c->tol));
  l.set("type_multinomial", std::string(c->grouped_multinomial ? "grouped" : "ungrouped"));
  return l;
}

template <typename T>
T* dup(const std::vector<T>& v) {
  T* p = static_cast<T*>(std::malloc(sizeof(T) * (v.empty() ? 1 : v.size())));
  if (!v.empty()) std::memcpy(p, v.data(), sizeof(T) * v.size());
  return p;
}

void unpack(const Rcpp::List& res, const sgdnet_control* c, int64_t p, sgdnet_result* out, double seconds) {
  const auto a0 = Rcpp::as<std::vector<Eigen::ArrayXd>>(res["a0"]);
  const auto beta = Rcpp::as<std::vector<Eigen::ArrayXXd>>(res["beta"]);
  const auto losses = Rcpp::as<std::vector<std::vector<double>>>(res["losses"]);
  const auto dev = Rcpp::as<std::vector<double>>(res["dev.ratio"]);
  const auto lambda = Rcpp::as<std::vector<double>>(res["lambda"]);
  const auto codes = Rcpp::as<std::vector<unsigned>>(res["return_codes"]);
  const size_t L = lambda.size();
  const size_t K = static_cast<size_t>(c->n_classes);
  std::memset(out, 0, sizeof(*out));
  out->n_lambda = static_cast<int32_t>(L);
  out->n_classes = c->n_classes;
  out->n_features = p;
  std::vector<double> a0f, bf, lf;
  std::vector<int64_t> lptr(L + 1, 0);
  std::vector<uint32_t> epochs(L, 0), rc(codes.begin(), codes.end());
  for (size_t l = 0; l < L; ++l) {
    if (static_cast<size_t>(a0[l].size()) != K || static_cast<size_t>(beta[l].size()) != K * static_cast<size_t>(p))
      throw std::runtime_error("unexpected archive shape");
    a0f.insert(a0f.end(), a0[l].data(), a0[l].data() + K);
    bf.insert(bf.end(), beta[l].data(), beta[l].data() + K * p);    /* K x p column-major = [p][K] */
    if (l < losses.size()) {
      epochs[l] = static_cast<uint32_t>(losses[l].size());
      if (c->debug) lf.insert(lf.end(), losses[l].begin(), losses[l].end());
    }
    lptr[l + 1] = static_cast<int64_t>(lf.size());
  }
  out->a0 = dup(a0f);
  out->beta = dup(bf);
  out->lambda = dup(lambda);
  out->dev_ratio = dup(dev);
  out->return_codes = dup(rc);
  out->epochs = dup(epochs);
  out->losses = dup(lf);
  out->losses_ptr = dup(lptr);
  out->nulldev = Rcpp::as<double>(res["nulldev"]);
  out->npasses = Rcpp::as<unsigned>(res["npasses"]);
  out->seconds_total = out->seconds_solver = seconds;
}

template <typename F>
int guarded(sgdnet_rng* rng, int64_t n, F f) {
  try {
    if (!rng) throw std::runtime_error("rng is null");
    g_rng = rng;
    g_n_for_sequence = n;
    f();
    g_rng = nullptr;
    return SGDNET_OK;
  } catch (const std::exception& e) {
    g_rng = nullptr;
    g_err = e.what();
    return SGDNET_ERR_INTERNAL;
  }
}

}  // namespace refentry

extern "C" {

void REF_SYM(set_force_debug)(int on) { refentry::g_force_debug = on; }

void REF_SYM(rng_set_seed)(sgdnet_rng* rng, uint32_t seed) {
  std::memset(rng, 0, sizeof(*rng));
  rng->kind = SGDNET_RNG_MT;
  for (int j = 0; j < 50; ++j) seed = 69069u * seed + 1u;
  uint32_t first = 0;
  for (int j = 0; j < 625; ++j) {
    seed = 69069u * seed + 1u;
    if (j == 0) first = seed; else rng->mt[j - 1] = seed;
  }
  (void)first;          /* i_seed[0] is overwritten by FixupSeeds: mti = 624 */
  rng->mti = 624;
}

double REF_SYM(rng_unif)(sgdnet_rng* rng) { return refentry::unif(rng); }

const char* REF_SYM(last_error)(void) { return refentry::g_err.c_str(); }

int REF_SYM(fit_dense)(const double* x, int64_t n, int64_t p, const double* y, int32_t y_cols, const sgdnet_control* control,
                  sgdnet_rng* rng, sgdnet_result* out) {
  return refentry::guarded(rng, n, [&] {
    Eigen::MatrixXd xm(n, p), ym(n, y_cols);
    std::memcpy(xm.data(), x, sizeof(double) * static_cast<size_t>(n * p));
    std::memcpy(ym.data(), y, sizeof(double) * static_cast<size_t>(n * y_cols));
    const auto t0 = std::chrono::steady_clock::now();
    const Rcpp::List res = SgdnetDense(xm, ym, refentry::make_control(control, false));
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    refentry::unpack(res, control, p, out, s);
  });
}

int REF_SYM(fit_sparse)(const int32_t* csc_i, const int32_t* csc_p, const double* csc_x, int64_t n, int64_t p, const double* y,
                   int32_t y_cols, const sgdnet_control* control, sgdnet_rng* rng, sgdnet_result* out) {
  return refentry::guarded(rng, n, [&] {
    Eigen::SparseMatrix<double> xm(n, p, csc_p, csc_i, csc_x);
    Eigen::MatrixXd ym(n, y_cols);
    std::memcpy(ym.data(), y, sizeof(double) * static_cast<size_t>(n * y_cols));
    const auto t0 = std::chrono::steady_clock::now();
    const Rcpp::List res = SgdnetSparse(xm, ym, refentry::make_control(control, true));
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    refentry::unpack(res, control, p, out, s);
  });
}

void REF_SYM(result_free)(sgdnet_result* r) {
  if (!r) return;
  std::free(r->a0); std::free(r->beta); std::free(r->lambda); std::free(r->dev_ratio);
  std::free(r->return_codes); std::free(r->epochs); std::free(r->losses); std::free(r->losses_ptr);
  std::memset(r, 0, sizeof(*r));
}

}  // extern "C"
