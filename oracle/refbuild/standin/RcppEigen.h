/*
 * RcppEigen.h — a minimal STAND-IN for the two external libraries the reference's C++ sources include
 * (`#include <RcppEigen.h>`: Rcpp >= 0.12.16 and Eigen 3.3.x through RcppEigen; reference DESCRIPTION:33-35).
 * TEST INFRASTRUCTURE ONLY. It exists so that the reference's own solver sources (src/sgdnet.cpp, saga-dense.h,
 * saga-sparse.h, penalties.h, prox.h, families.h, utils.h, math.h, constants.h) can be compiled UNMODIFIED, from where
 * they lie under /root/reference/src, into oracle/_ref/libsgdnet_ref.so (recipe: oracle/refbuild/Makefile), and the
 * restated oracle (oracle/sgdnet_oracle.cpp) can be checked against them bit for bit.
 *
 * This file is our own code, written against the list of expressions the reference uses (nothing here is taken from
 * Eigen or Rcpp). It implements exactly that subset, eagerly: every expression returns a concrete temporary.
 * Semantics follow Eigen's documented behaviour for each expression:
 *   - column-major dense storage; Array types are coefficient-wise, Matrix types are linear-algebra;
 *   - `dense_row_block (+|-)= column_vector_expression` is transposed automatically (Eigen's compile-time-vector rule),
 *     which is what src/saga-sparse.h:124-128 relies on;
 *   - `colvec.rowwise() * rowvec` is the outer product array (src/saga-dense.h:176, 183);
 *   - dense x sparse-vector and sparse^T x dense products accumulate over the stored nonzeros in ascending index order
 *     starting from 0 (this is also what Eigen does);
 *   - sparse transpose-eval yields sorted inner indices.
 * WHAT IT DOES NOT PIN: Eigen evaluates dense reductions (GEMV, GEMM, .sum(), .squaredNorm(), .norm()) and array
 * exp()/log() with SIMD packets in an implementation-defined association; here every reduction is a plain ascending
 * sequential sum and exp/log are std::exp/std::log. That is the same reading oracle/sgdnet_oracle.cpp takes in its
 * "libm" mode, so the two can be compared exactly; it is not a claim about the bits of a real R/Eigen build.
 */
#ifndef SGDNET_STANDIN_RCPPEIGEN_H_
#define SGDNET_STANDIN_RCPPEIGEN_H_

#include <algorithm>
#include <any>
#include <cmath>
#include <cstddef>
#include <limits>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace Eigen {

typedef std::ptrdiff_t Index;

class Arr;
class Mat;

/* ---------------------------------------------------------------- small read-only views */
struct RowArr {               /* a 1 x n array expression (x.col(s).transpose().array()) */
  std::vector<double> v;
};
struct RowVecT {              /* transpose of a dense column */
  const double* p;
  Index n;
  RowArr array() const { return RowArr{std::vector<double>(p, p + n)}; }
};
struct SparseVec {            /* an evaluated sparse column expression */
  std::vector<Index> idx;
  std::vector<double> val;
  SparseVec operator*(double s) const {
    SparseVec o{idx, val};
    for (double& x : o.val) x = x * s;
    return o;
  }
};

/* ---------------------------------------------------------------- Array (ArrayXd, ArrayXXd) */
class ArrCol;
class ArrRow;
class ArrRowwise;

class Arr {
 public:
  Arr() : r_(0), c_(0) {}
  explicit Arr(Index n) : r_(n), c_(1), d_(static_cast<size_t>(n)) {}
  Arr(Index r, Index c) : r_(r), c_(c), d_(static_cast<size_t>(r * c)) {}
  Arr(const Mat& m);                 /* ArrayXXd a = matrix_expression; */
  Arr(const ArrCol& c);
  static Arr Zero(Index n) { Arr a(n); return a; }
  static Arr Zero(Index r, Index c) { Arr a(r, c); return a; }
  static Arr Ones(Index n) { Arr a(n); std::fill(a.d_.begin(), a.d_.end(), 1.0); return a; }
  Arr& operator=(double v) { std::fill(d_.begin(), d_.end(), v); return *this; }
  void setConstant(double v) { *this = v; }
  Arr eval() const { return *this; }

  Index rows() const { return r_; }
  Index cols() const { return c_; }
  Index size() const { return r_ * c_; }
  double* data() { return d_.data(); }
  const double* data() const { return d_.data(); }
  double& operator()(Index i) { return d_[static_cast<size_t>(i)]; }
  double operator()(Index i) const { return d_[static_cast<size_t>(i)]; }
  double& operator[](Index i) { return d_[static_cast<size_t>(i)]; }
  double operator[](Index i) const { return d_[static_cast<size_t>(i)]; }
  double& operator()(Index i, Index j) { return d_[static_cast<size_t>(i + j * r_)]; }
  double operator()(Index i, Index j) const { return d_[static_cast<size_t>(i + j * r_)]; }

  ArrCol col(Index j);
  Arr col(Index j) const {
    Arr o(r_);
    for (Index i = 0; i < r_; ++i) o(i) = (*this)(i, j);
    return o;
  }
  ArrRow row(Index k);
  ArrRowwise rowwise() const;
  struct MatView matrix() const;

  /* reductions: ascending sequential */
  double sum() const { double s = 0.0; for (double v : d_) s += v; return s; }
  double maxCoeff() const { double m = d_.at(0); for (double v : d_) m = v > m ? v : m; return m; }
  /* coefficient-wise maps */
  template <typename F> Arr map(F f) const { Arr o(r_, c_); for (size_t i = 0; i < d_.size(); ++i) o.d_[i] = f(d_[i]); return o; }
  Arr abs() const { return map([](double v) { return std::fabs(v); }); }
  Arr exp() const { return map([](double v) { return std::exp(v); }); }
  Arr log() const { return map([](double v) { return std::log(v); }); }
  Arr sqrt() const { return map([](double v) { return std::sqrt(v); }); }
  Arr square() const { return map([](double v) { return v * v; }); }

  Arr& operator*=(double s) { for (double& v : d_) v *= s; return *this; }
  Arr& operator+=(const Arr& o) { chk(o); for (size_t i = 0; i < d_.size(); ++i) d_[i] += o.d_[i]; return *this; }
  Arr& operator-=(const Arr& o) { chk(o); for (size_t i = 0; i < d_.size(); ++i) d_[i] -= o.d_[i]; return *this; }

  void chk(const Arr& o) const { if (o.r_ != r_ || o.c_ != c_) throw std::logic_error("standin: array shape mismatch"); }

 private:
  Index r_, c_;
  std::vector<double> d_;
};

template <typename F>
inline Arr zip(const Arr& a, const Arr& b, F f) {
  a.chk(b);
  Arr o(a.rows(), a.cols());
  for (Index i = 0; i < a.size(); ++i) o(i) = f(a(i), b(i));
  return o;
}
inline Arr operator+(const Arr& a, const Arr& b) { return zip(a, b, [](double x, double y) { return x + y; }); }
inline Arr operator-(const Arr& a, const Arr& b) { return zip(a, b, [](double x, double y) { return x - y; }); }
inline Arr operator*(const Arr& a, const Arr& b) { return zip(a, b, [](double x, double y) { return x * y; }); }
inline Arr operator/(const Arr& a, const Arr& b) { return zip(a, b, [](double x, double y) { return x / y; }); }
inline Arr operator*(const Arr& a, double s) { return a.map([s](double v) { return v * s; }); }
inline Arr operator*(double s, const Arr& a) { return a.map([s](double v) { return s * v; }); }
inline Arr operator/(const Arr& a, double s) { return a.map([s](double v) { return v / s; }); }
inline Arr operator-(const Arr& a, double s) { return a.map([s](double v) { return v - s; }); }
inline Arr operator+(const Arr& a, double s) { return a.map([s](double v) { return v + s; }); }

/* mutable column of an Array (w.col(j), g_memory.col(s), ...) */
class ArrCol {
 public:
  ArrCol(double* p, Index n) : p_(p), n_(n) {}
  Index size() const { return n_; }
  double operator()(Index i) const { return p_[i]; }
  ArrCol& operator=(const Arr& a) { chk(a); for (Index i = 0; i < n_; ++i) p_[i] = a(i); return *this; }
  ArrCol& operator=(const ArrCol& a) { for (Index i = 0; i < n_; ++i) p_[i] = a.p_[i]; return *this; }
  ArrCol& operator=(double v) { for (Index i = 0; i < n_; ++i) p_[i] = v; return *this; }
  ArrCol& operator-=(const Arr& a) { chk(a); for (Index i = 0; i < n_; ++i) p_[i] -= a(i); return *this; }
  ArrCol& operator*=(const Arr& a) { chk(a); for (Index i = 0; i < n_; ++i) p_[i] *= a(i); return *this; }
  ArrCol& operator*=(double s) { for (Index i = 0; i < n_; ++i) p_[i] *= s; return *this; }
 private:
  void chk(const Arr& a) const { if (a.size() != n_) throw std::logic_error("standin: column length mismatch"); }
  double* p_;
  Index n_;
};
inline Arr::Arr(const ArrCol& c) : r_(c.size()), c_(1), d_(static_cast<size_t>(c.size())) {
  for (Index i = 0; i < r_; ++i) d_[static_cast<size_t>(i)] = c(i);
}
inline ArrCol Arr::col(Index j) { return ArrCol(d_.data() + j * r_, r_); }

/* mutable row of an Array; a column-vector right-hand side is transposed automatically (Eigen's vector rule) */
class ArrRow {
 public:
  ArrRow(Arr* a, Index k) : a_(a), k_(k) {}
  ArrRow& operator+=(const SparseVec& v) {
    for (size_t e = 0; e < v.idx.size(); ++e) (*a_)(k_, v.idx[e]) += v.val[e];
    return *this;
  }
  ArrRow& operator-=(const Arr& v) {
    if (v.size() != a_->cols()) throw std::logic_error("standin: row length mismatch");
    for (Index j = 0; j < a_->cols(); ++j) (*a_)(k_, j) -= v(j);
    return *this;
  }
 private:
  Arr* a_;
  Index k_;
};
inline ArrRow Arr::row(Index k) { return ArrRow(this, k); }

/* colvec.rowwise() * rowvec -> outer product; matrix.rowwise().sum() -> per-row sums over ascending columns */
class ArrRowwise {
 public:
  explicit ArrRowwise(const Arr& a) : a_(a) {}
  Arr operator*(const RowArr& x) const {
    if (a_.cols() != 1) throw std::logic_error("standin: rowwise()*row expects a column vector");
    const Index K = a_.rows(), p = static_cast<Index>(x.v.size());
    Arr o(K, p);
    for (Index j = 0; j < p; ++j)
      for (Index k = 0; k < K; ++k) o(k, j) = a_(k) * x.v[static_cast<size_t>(j)];
    return o;
  }
  Arr sum() const {
    Arr o(a_.rows());
    for (Index i = 0; i < a_.rows(); ++i) {
      double s = 0.0;
      for (Index j = 0; j < a_.cols(); ++j) s += a_(i, j);
      o(i) = s;
    }
    return o;
  }
 private:
  const Arr& a_;
};
inline ArrRowwise Arr::rowwise() const { return ArrRowwise(*this); }

/* ---------------------------------------------------------------- Matrix (MatrixXd, VectorXd) */
struct MatColC {              /* read-only dense column */
  const double* p;
  Index n;
  double sum() const { double s = 0.0; for (Index i = 0; i < n; ++i) s += p[i]; return s; }
  double squaredNorm() const { double s = 0.0; for (Index i = 0; i < n; ++i) s += p[i] * p[i]; return s; }
  double norm() const { return std::sqrt(squaredNorm()); }
  Arr array() const { Arr a(n); for (Index i = 0; i < n; ++i) a(i) = p[i]; return a; }
  RowVecT transpose() const { return RowVecT{p, n}; }
};
struct MatCol : MatColC {     /* mutable dense column */
  double* q;
  MatCol(double* ptr, Index len) : MatColC{ptr, len}, q(ptr) {}
  MatCol& operator=(const Arr& a) {
    if (a.size() != n) throw std::logic_error("standin: column length mismatch");
    for (Index i = 0; i < n; ++i) q[i] = a(i);
    return *this;
  }
};

struct MatView {              /* read-only column-major matrix view (w.matrix(), x_center_scaled.matrix()) */
  const double* d;
  Index r, c;
  Index rows() const { return r; }
  Index cols() const { return c; }
  double operator()(Index i, Index j) const { return d[i + j * r]; }
  MatColC col(Index j) const { return MatColC{d + j * r, r}; }
};
inline MatView Arr::matrix() const { return MatView{d_.data(), r_, c_}; }

class MatColwise;
struct MatArrayView;

class Mat {
 public:
  Mat() : r_(0), c_(0) {}
  explicit Mat(Index n) : r_(n), c_(1), d_(static_cast<size_t>(n)) {}
  Mat(Index r, Index c) : r_(r), c_(c), d_(static_cast<size_t>(r * c)) {}
  static Mat Zero(Index r, Index c) { return Mat(r, c); }
  Index rows() const { return r_; }
  Index cols() const { return c_; }
  Index size() const { return r_ * c_; }
  double* data() { return d_.data(); }
  const double* data() const { return d_.data(); }
  double& operator()(Index i) { return d_[static_cast<size_t>(i)]; }
  double operator()(Index i) const { return d_[static_cast<size_t>(i)]; }
  double& operator()(Index i, Index j) { return d_[static_cast<size_t>(i + j * r_)]; }
  double operator()(Index i, Index j) const { return d_[static_cast<size_t>(i + j * r_)]; }
  MatCol col(Index j) { return MatCol(d_.data() + j * r_, r_); }
  MatColC col(Index j) const { return MatColC{d_.data() + j * r_, r_}; }
  operator MatView() const { return MatView{d_.data(), r_, c_}; }
  Mat transpose() const {
    Mat t(c_, r_);
    for (Index j = 0; j < c_; ++j)
      for (Index i = 0; i < r_; ++i) t(j, i) = (*this)(i, j);
    return t;
  }
  void transposeInPlace() { *this = transpose(); }
  MatArrayView array() const;   /* y.array().col(i); (W x).array() * wscale + b */
  MatColwise colwise() const;
  Mat cwiseAbs() const { Mat o(r_, c_); for (size_t i = 0; i < d_.size(); ++i) o.d_[i] = std::fabs(d_[i]); return o; }
  double maxCoeff() const { double m = d_.at(0); for (double v : d_) m = v > m ? v : m; return m; }
  double squaredNorm() const { double s = 0.0; for (double v : d_) s += v * v; return s; }
 private:
  Index r_, c_;
  std::vector<double> d_;
};
inline Arr::Arr(const Mat& m) : r_(m.rows()), c_(m.cols()), d_(m.data(), m.data() + m.size()) {}

/* matrix.array(): a view (y.array().col(i) must not copy y on every update); converts to the Array value */
struct MatArrayView {
  const Mat& m;
  Arr col(Index i) const { return m.col(i).array(); }
  operator Arr() const { return Arr(m); }
};
inline MatArrayView Mat::array() const { return MatArrayView{*this}; }

class MatColwise {
 public:
  explicit MatColwise(const Mat& m) : m_(m) {}
  Mat squaredNorm() const {
    Mat o(1, m_.cols());
    for (Index j = 0; j < m_.cols(); ++j) o(0, j) = m_.col(j).squaredNorm();
    return o;
  }
 private:
  const Mat& m_;
};
inline MatColwise Mat::colwise() const { return MatColwise(*this); }

/* dense products: res(i,k) = sum over the inner index in ascending order, starting from 0 */
inline Mat matmul(const MatView& a, const MatView& b) {
  if (a.c != b.r) throw std::logic_error("standin: product shape mismatch");
  Mat o(a.r, b.c);
  for (Index k = 0; k < b.c; ++k)
    for (Index i = 0; i < a.r; ++i) {
      double s = 0.0;
      for (Index j = 0; j < a.c; ++j) s += a(i, j) * b(j, k);
      o(i, k) = s;
    }
  return o;
}
inline Mat operator*(const MatView& a, const MatView& b) { return matmul(a, b); }
inline Mat operator*(const MatView& a, const Mat& b) { return matmul(a, MatView(b)); }
inline Mat operator*(const Mat& a, const Mat& b) { return matmul(MatView(a), MatView(b)); }
inline Mat operator*(const MatView& a, const MatColC& x) { return matmul(a, MatView{x.p, x.n, 1}); }

/* ---------------------------------------------------------------- SparseMatrix<double>, column-major */
template <typename T>
class SparseMatrix;

struct SparseCol {
  const int* idx;
  const double* val;
  Index nnz;
  Index len;
  Index nonZeros() const { return nnz; }
  double sum() const { double s = 0.0; for (Index e = 0; e < nnz; ++e) s += val[e]; return s; }
  double squaredNorm() const { double s = 0.0; for (Index e = 0; e < nnz; ++e) s += val[e] * val[e]; return s; }
  SparseVec operator*(double s) const {
    SparseVec o;
    o.idx.assign(idx, idx + nnz);
    o.val.resize(static_cast<size_t>(nnz));
    for (Index e = 0; e < nnz; ++e) o.val[static_cast<size_t>(e)] = val[e] * s;
    return o;
  }
};
/* sparse column - dense vector -> dense vector */
inline Mat operator-(const SparseCol& x, const MatView& c) {
  if (c.r * c.c != x.len) throw std::logic_error("standin: sparse - dense length mismatch");
  Mat o(x.len);
  for (Index j = 0; j < x.len; ++j) o(j) = 0.0 - c.d[j];
  for (Index e = 0; e < x.nnz; ++e) o(x.idx[e]) = x.val[e] - c.d[x.idx[e]];
  return o;
}
/* dense x sparse vector: per output row, tmp = 0; tmp += value * lhs(row, index) over ascending nonzeros */
inline Mat operator*(const MatView& a, const SparseCol& x) {
  if (a.c != x.len) throw std::logic_error("standin: dense x sparse shape mismatch");
  Mat o(a.r);
  for (Index k = 0; k < a.r; ++k) {
    double s = 0.0;
    for (Index e = 0; e < x.nnz; ++e) s += x.val[e] * a(k, x.idx[e]);
    o(k) = s;
  }
  return o;
}

template <typename T>
class SparseTransposed;

template <>
class SparseMatrix<double> {
 public:
  SparseMatrix() : r_(0), c_(0), ptr_(1, 0) {}
  /* from compressed-column arrays (a dgCMatrix) */
  SparseMatrix(Index rows, Index cols, const int* cp, const int* ri, const double* xv) : r_(rows), c_(cols) {
    ptr_.assign(cp, cp + cols + 1);
    idx_.assign(ri, ri + cp[cols]);
    val_.assign(xv, xv + cp[cols]);
  }
  Index rows() const { return r_; }
  Index cols() const { return c_; }
  /* compressed-column storage access (StorageIndex = int, as in Eigen) */
  bool isCompressed() const { return true; }
  void makeCompressed() {}
  const int* outerIndexPtr() const { return ptr_.data(); }
  const int* innerIndexPtr() const { return idx_.data(); }
  const double* valuePtr() const { return val_.data(); }
  SparseCol col(Index j) const {
    const Index s = ptr_[static_cast<size_t>(j)];
    return SparseCol{idx_.data() + s, val_.data() + s, ptr_[static_cast<size_t>(j) + 1] - s, r_};
  }
  SparseTransposed<double> transpose() const;

  class InnerIterator {
   public:
    InnerIterator(const SparseMatrix& m, Index outer)
        : m_(const_cast<SparseMatrix*>(&m)), e_(m.ptr_[static_cast<size_t>(outer)]), end_(m.ptr_[static_cast<size_t>(outer) + 1]) {}
    operator bool() const { return e_ < end_; }
    InnerIterator& operator++() { ++e_; return *this; }
    Index index() const { return m_->idx_[static_cast<size_t>(e_)]; }
    double value() const { return m_->val_[static_cast<size_t>(e_)]; }
    double& valueRef() { return m_->val_[static_cast<size_t>(e_)]; }
   private:
    SparseMatrix* m_;
    Index e_, end_;
  };

  std::vector<int> ptr_, idx_;
  std::vector<double> val_;
 private:
  Index r_, c_;
  friend class SparseTransposed<double>;
};

template <>
class SparseTransposed<double> {
 public:
  explicit SparseTransposed(const SparseMatrix<double>& m) : m_(m) {}
  /* the transposed matrix in compressed-column form, inner indices ascending */
  SparseMatrix<double> eval() const {
    SparseMatrix<double> t;
    t.r_ = m_.c_;
    t.c_ = m_.r_;
    t.ptr_.assign(static_cast<size_t>(m_.r_) + 1, 0);
    for (int i : m_.idx_) ++t.ptr_[static_cast<size_t>(i) + 1];
    for (size_t i = 1; i < t.ptr_.size(); ++i) t.ptr_[i] += t.ptr_[i - 1];
    t.idx_.resize(m_.idx_.size());
    t.val_.resize(m_.val_.size());
    std::vector<int> fill(t.ptr_.begin(), t.ptr_.end() - 1);
    for (Index j = 0; j < m_.c_; ++j)
      for (Index e = m_.ptr_[static_cast<size_t>(j)]; e < m_.ptr_[static_cast<size_t>(j) + 1]; ++e) {
        const Index i = m_.idx_[static_cast<size_t>(e)];
        const Index dst = fill[static_cast<size_t>(i)]++;
        t.idx_[static_cast<size_t>(dst)] = static_cast<int>(j);
        t.val_[static_cast<size_t>(dst)] = m_.val_[static_cast<size_t>(e)];
      }
    return t;
  }
  /* x^T * dense: row j of x^T is column j of x; tmp += value * rhs(index, k) over ascending nonzeros */
  Mat operator*(const Mat& y) const {
    if (y.rows() != m_.rows()) throw std::logic_error("standin: sparse^T x dense shape mismatch");
    Mat o(m_.cols(), y.cols());
    for (Index k = 0; k < y.cols(); ++k)
      for (Index j = 0; j < m_.cols(); ++j) {
        const SparseCol cj = m_.col(j);
        double s = 0.0;
        for (Index e = 0; e < cj.nnz; ++e) s += cj.val[e] * y(cj.idx[e], k);
        o(j, k) = s;
      }
    return o;
  }
 private:
  const SparseMatrix<double>& m_;
};
inline SparseTransposed<double> SparseMatrix<double>::transpose() const { return SparseTransposed<double>(*this); }

typedef Arr ArrayXd;
typedef Arr ArrayXXd;
typedef Mat MatrixXd;
typedef Mat VectorXd;

}  // namespace Eigen

/* ---------------------------------------------------------------- Rcpp subset */
namespace Rcpp {

struct NamedValue {
  std::string name;
  std::any value;
};
struct Named {
  explicit Named(const std::string& n) : name(n) {}
  template <typename T>
  NamedValue operator=(const T& v) const { return NamedValue{name, std::any(v)}; }
  std::string name;
};
template <typename T>
inline T wrap(const T& v) { return v; }

class List {
 public:
  static List create() { return List(); }
  template <typename... A>
  static List create(const A&... a) {
    List l;
    (l.items_.insert_or_assign(a.name, a.value), ...);
    return l;
  }
  const std::any& operator[](const std::string& key) const {
    auto it = items_.find(key);
    if (it == items_.end()) throw std::out_of_range("standin Rcpp::List: no element named " + key);
    return it->second;
  }
  template <typename T>
  void set(const std::string& key, const T& v) { items_.insert_or_assign(key, std::any(v)); }
  bool has(const std::string& key) const { return items_.count(key) != 0; }
 private:
  std::map<std::string, std::any> items_;
};

template <typename T>
inline T as(const std::any& v) { return std::any_cast<T>(v); }

[[noreturn]] inline void stop(const std::string& msg) { throw std::runtime_error(msg); }

}  // namespace Rcpp

/* ---------------------------------------------------------------- R's runif (R core nmath/runif.c) */
namespace R {
double unif_rand_hook();   /* defined by the entry file: R's unif_rand() on the caller's generator */
inline double runif(double a, double b) {
  if (a == b) return a;
  double u;
  do { u = unif_rand_hook(); } while (u <= 0 || u >= 1);
  return a + (b - a) * u;
}
}  // namespace R

/* R API (R_ext/Random.h): one draw from R's global generator */
inline double unif_rand() { return R::unif_rand_hook(); }

#endif /* SGDNET_STANDIN_RCPPEIGEN_H_ */
