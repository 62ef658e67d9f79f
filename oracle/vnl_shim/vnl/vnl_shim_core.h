// TEST INFRASTRUCTURE ONLY -- not part of the shipped engine.
//
// Stand-in for the subset of VNL (VXL / ITK `vnl`, `vnl_algo`) that the reference
// (zivy/LSQRRecipes) uses on the RANSAC hot path.  VNL is a third-party dependency of the
// reference that is NOT vendored in /root/reference and NOT version-pinned anywhere
// (CMakeLists.txt:26-70 just does find_package(VXL|ITK REQUIRED)).  It is absent from this
// image and there is no network, so the reference's own sources are compiled UNMODIFIED
// against this shim to obtain `oracle/_ref/`.
//
// What is restated here (from VNL's published behaviour, not from its source):
//   * vnl_vector / vnl_matrix / vnl_vector_ref / vnl_vector_fixed containers, with the
//     accumulation orders VNL documents: dot products and matrix products accumulate
//     left-to-right starting from T(0); vnl_vector::normalize() multiplies by 1/sqrt(sum x^2).
//   * vnl_symmetric_eigensystem: eigenvalues ascending, eigenvectors are the COLUMNS of V,
//     sign arbitrary.  (VNL: EISPACK rs = tred2+tql2.  Here: cyclic Jacobi.)
//   * vnl_svd: singular values descending; zero_out_absolute(t) zeroes |s|<=t and sets rank;
//     nullvector() = last column of V; pinverse() = V W^-1 U^T.  (VNL: LINPACK dsvdc.
//     Here: one-sided Jacobi (Hestenes).)
//   * vnl_matrix_inverse<T> = vnl_svd<T> whose product with a vector applies pinverse().
//   * vnl_levenberg_marquardt: VNL calls MINPACK lmder (mode 1, factor 100) and returns true iff
//     info is 1..4.  Here: the MINPACK algorithm restated step for step (oracle/minpack_lm.h)
//     with VNL's defaults (xtol 1e-8, ftol = xtol * 0.01, gtol 1e-5, maxfev = 400 n).
//
// Consequence: everything in the reference that is plain `double` arithmetic (all estimate()
// and agree() bodies for plane-3D, line, line-2D, circle, sphere, ray intersection) is
// bit-for-bit the reference.  Results that flow through an eigen/SVD/LM routine agree with
// real VNL to rounding level (~1e-12 relative), not bit-for-bit; parity tests on those
// paths use tolerances and say so.
#ifndef LSQR_ORACLE_VNL_SHIM_CORE_H
#define LSQR_ORACLE_VNL_SHIM_CORE_H

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <vector>

#include "../../minpack_lm.h"

template <class T> class vnl_matrix;

template <class T>
class vnl_vector {
 public:
  vnl_vector() {}
  explicit vnl_vector(unsigned n) : d_(n) {}
  vnl_vector(unsigned n, T const& v) : d_(n, v) {}
  vnl_vector(T const* p, unsigned n) : d_(p, p + n) {}
  unsigned size() const { return static_cast<unsigned>(d_.size()); }
  void set_size(unsigned n) { d_.resize(n); }
  T& operator[](unsigned i) { return d_[i]; }
  T const& operator[](unsigned i) const { return d_[i]; }
  T& operator()(unsigned i) { return d_[i]; }
  T const& operator()(unsigned i) const { return d_[i]; }
  T* data_block() { return d_.data(); }
  T const* data_block() const { return d_.data(); }
  void fill(T const& v) { std::fill(d_.begin(), d_.end(), v); }

  vnl_vector operator+(vnl_vector const& o) const { vnl_vector r(*this); for (unsigned i = 0; i < size(); i++) r[i] = d_[i] + o[i]; return r; }
  vnl_vector operator-(vnl_vector const& o) const { vnl_vector r(*this); for (unsigned i = 0; i < size(); i++) r[i] = d_[i] - o[i]; return r; }
  vnl_vector operator-() const { vnl_vector r(*this); for (unsigned i = 0; i < size(); i++) r[i] = -d_[i]; return r; }
  vnl_vector operator*(T s) const { vnl_vector r(*this); for (unsigned i = 0; i < size(); i++) r[i] = d_[i] * s; return r; }
  vnl_vector operator/(T s) const { vnl_vector r(*this); for (unsigned i = 0; i < size(); i++) r[i] = d_[i] / s; return r; }
  vnl_vector& operator+=(vnl_vector const& o) { for (unsigned i = 0; i < size(); i++) d_[i] += o[i]; return *this; }
  vnl_vector& operator-=(vnl_vector const& o) { for (unsigned i = 0; i < size(); i++) d_[i] -= o[i]; return *this; }
  vnl_vector& operator*=(T s) { for (unsigned i = 0; i < size(); i++) d_[i] *= s; return *this; }
  vnl_vector& operator/=(T s) { for (unsigned i = 0; i < size(); i++) d_[i] /= s; return *this; }

  T squared_magnitude() const { T s(0); for (unsigned i = 0; i < size(); i++) s += d_[i] * d_[i]; return s; }
  T magnitude() const { return std::sqrt(squared_magnitude()); }
  T two_norm() const { return magnitude(); }
  // VNL: scale by the reciprocal of the norm (one division, n multiplications).
  vnl_vector& normalize() {
    T s = squared_magnitude();
    if (s != T(0)) { T inv = T(1) / std::sqrt(s); for (unsigned i = 0; i < size(); i++) d_[i] = inv * d_[i]; }
    return *this;
  }
  vnl_vector& update(vnl_vector const& v, unsigned start = 0) { for (unsigned i = 0; i < v.size(); i++) d_[start + i] = v[i]; return *this; }

 protected:
  std::vector<T> d_;
};

template <class T> inline vnl_vector<T> operator*(T const& s, vnl_vector<T> const& v) { return v * s; }
template <class T> inline T dot_product(vnl_vector<T> const& a, vnl_vector<T> const& b) { T s(0); for (unsigned i = 0; i < a.size(); i++) s += a[i] * b[i]; return s; }
template <class T> inline vnl_vector<T> vnl_cross_3d(vnl_vector<T> const& a, vnl_vector<T> const& b) {
  vnl_vector<T> r(3);
  r[0] = a[1] * b[2] - a[2] * b[1];
  r[1] = a[2] * b[0] - a[0] * b[2];
  r[2] = a[0] * b[1] - a[1] * b[0];
  return r;
}

// The reference only ever reads through these views (Point.h:73,125; Vector.h:155), so a
// by-value snapshot is behaviourally identical.
template <class T>
class vnl_vector_ref : public vnl_vector<T> {
 public:
  vnl_vector_ref(unsigned n, T* p) : vnl_vector<T>(static_cast<T const*>(p), n) {}
};

template <class T, unsigned n>
class vnl_vector_fixed {
 public:
  vnl_vector_fixed() { for (unsigned i = 0; i < n; i++) d_[i] = T(0); }
  T& operator[](unsigned i) { return d_[i]; }
  T const& operator[](unsigned i) const { return d_[i]; }
  T& operator()(unsigned i) { return d_[i]; }
  unsigned size() const { return n; }
 private:
  T d_[n];
};

template <class T>
class vnl_matrix {
 public:
  vnl_matrix() : r_(0), c_(0) {}
  vnl_matrix(unsigned r, unsigned c) : r_(r), c_(c), d_(size_t(r) * c) {}
  vnl_matrix(unsigned r, unsigned c, T const& v) : r_(r), c_(c), d_(size_t(r) * c, v) {}
  vnl_matrix(T const* p, unsigned r, unsigned c) : r_(r), c_(c), d_(p, p + size_t(r) * c) {}
  unsigned rows() const { return r_; }
  unsigned cols() const { return c_; }
  unsigned columns() const { return c_; }
  void set_size(unsigned r, unsigned c) { r_ = r; c_ = c; d_.assign(size_t(r) * c, T(0)); }
  T& operator()(unsigned i, unsigned j) { return d_[size_t(i) * c_ + j]; }
  T const& operator()(unsigned i, unsigned j) const { return d_[size_t(i) * c_ + j]; }
  T* operator[](unsigned i) { return &d_[size_t(i) * c_]; }
  T const* operator[](unsigned i) const { return &d_[size_t(i) * c_]; }
  T* data_block() { return d_.data(); }
  T const* data_block() const { return d_.data(); }
  vnl_matrix& fill(T const& v) { std::fill(d_.begin(), d_.end(), v); return *this; }
  vnl_matrix& fill_diagonal(T const& v) { for (unsigned i = 0; i < r_ && i < c_; i++) (*this)(i, i) = v; return *this; }
  vnl_matrix transpose() const { vnl_matrix t(c_, r_); for (unsigned i = 0; i < r_; i++) for (unsigned j = 0; j < c_; j++) t(j, i) = (*this)(i, j); return t; }

  vnl_matrix operator*(vnl_matrix const& b) const {
    vnl_matrix r(r_, b.c_);
    for (unsigned i = 0; i < r_; i++)
      for (unsigned k = 0; k < b.c_; k++) { T s(0); for (unsigned j = 0; j < c_; j++) s += (*this)(i, j) * b(j, k); r(i, k) = s; }
    return r;
  }
  vnl_vector<T> operator*(vnl_vector<T> const& v) const {
    vnl_vector<T> r(r_);
    for (unsigned i = 0; i < r_; i++) { T s(0); for (unsigned j = 0; j < c_; j++) s += (*this)(i, j) * v[j]; r[i] = s; }
    return r;
  }
  vnl_matrix operator*(T s) const { vnl_matrix r(*this); for (size_t i = 0; i < d_.size(); i++) r.d_[i] = d_[i] * s; return r; }
  vnl_matrix operator+(vnl_matrix const& o) const { vnl_matrix r(*this); for (size_t i = 0; i < d_.size(); i++) r.d_[i] = d_[i] + o.d_[i]; return r; }
  vnl_matrix operator-(vnl_matrix const& o) const { vnl_matrix r(*this); for (size_t i = 0; i < d_.size(); i++) r.d_[i] = d_[i] - o.d_[i]; return r; }
  vnl_matrix& operator+=(vnl_matrix const& o) { for (size_t i = 0; i < d_.size(); i++) d_[i] += o.d_[i]; return *this; }
  vnl_matrix& operator-=(vnl_matrix const& o) { for (size_t i = 0; i < d_.size(); i++) d_[i] -= o.d_[i]; return *this; }
  vnl_matrix& operator*=(T s) { for (size_t i = 0; i < d_.size(); i++) d_[i] *= s; return *this; }

  vnl_matrix& update(vnl_matrix const& m, unsigned top = 0, unsigned left = 0) {
    for (unsigned i = 0; i < m.r_; i++) for (unsigned j = 0; j < m.c_; j++) (*this)(top + i, left + j) = m(i, j);
    return *this;
  }
  vnl_matrix& set_column(unsigned j, vnl_vector<T> const& v) { for (unsigned i = 0; i < r_; i++) (*this)(i, j) = v[i]; return *this; }
  vnl_matrix& set_row(unsigned i, vnl_vector<T> const& v) { for (unsigned j = 0; j < c_; j++) (*this)(i, j) = v[j]; return *this; }
  vnl_matrix& set_row(unsigned i, T const* v) { for (unsigned j = 0; j < c_; j++) (*this)(i, j) = v[j]; return *this; }   // DenseLinearEquationSystemParametersEstimator.hxx:34
  vnl_matrix& set_column(unsigned j, T const* v) { for (unsigned i = 0; i < r_; i++) (*this)(i, j) = v[i]; return *this; }
  vnl_vector<T> get_row(unsigned i) const { vnl_vector<T> v(c_); for (unsigned j = 0; j < c_; j++) v[j] = (*this)(i, j); return v; }
  vnl_vector<T> get_column(unsigned j) const { vnl_vector<T> v(r_); for (unsigned i = 0; i < r_; i++) v[i] = (*this)(i, j); return v; }

 private:
  unsigned r_, c_;
  std::vector<T> d_;
};

// ---------------------------------------------------------------------------------------
// Symmetric eigensystem (cyclic Jacobi).  D ascending, eigenvectors = columns of V.
// ---------------------------------------------------------------------------------------
template <class T>
class vnl_symmetric_eigensystem {
 public:
  vnl_matrix<T> V;
  vnl_vector<T> D;  // VNL exposes a diag matrix; the reference never touches it.
  explicit vnl_symmetric_eigensystem(vnl_matrix<T> const& M) : V(M.rows(), M.rows(), T(0)), D(M.rows()) {
    const unsigned n = M.rows();
    vnl_matrix<T> a(M);
    for (unsigned i = 0; i < n; i++) V(i, i) = T(1);
    for (int sweep = 0; sweep < 100; sweep++) {
      T off(0);
      for (unsigned p = 0; p < n; p++) for (unsigned q = p + 1; q < n; q++) off += a(p, q) * a(p, q);
      if (off == T(0)) break;
      for (unsigned p = 0; p < n; p++) {
        for (unsigned q = p + 1; q < n; q++) {
          if (a(p, q) == T(0)) continue;
          T theta = (a(q, q) - a(p, p)) / (T(2) * a(p, q));
          T t = (theta >= T(0) ? T(1) : T(-1)) / (std::fabs(theta) + std::sqrt(theta * theta + T(1)));
          T c = T(1) / std::sqrt(t * t + T(1)), s = t * c;
          for (unsigned k = 0; k < n; k++) { T akp = a(k, p), akq = a(k, q); a(k, p) = c * akp - s * akq; a(k, q) = s * akp + c * akq; }
          for (unsigned k = 0; k < n; k++) { T apk = a(p, k), aqk = a(q, k); a(p, k) = c * apk - s * aqk; a(q, k) = s * apk + c * aqk; }
          for (unsigned k = 0; k < n; k++) { T vkp = V(k, p), vkq = V(k, q); V(k, p) = c * vkp - s * vkq; V(k, q) = s * vkp + c * vkq; }
        }
      }
    }
    // sort ascending (selection sort on columns)
    for (unsigned i = 0; i < n; i++) D[i] = a(i, i);
    for (unsigned i = 0; i + 1 < n; i++) {
      unsigned m = i;
      for (unsigned j = i + 1; j < n; j++) if (D[j] < D[m]) m = j;
      if (m != i) { std::swap(D[i], D[m]); for (unsigned k = 0; k < n; k++) std::swap(V(k, i), V(k, m)); }
    }
  }
  T get_eigenvalue(int i) const { return D[i]; }
  vnl_vector<T> get_eigenvector(int i) const { return V.get_column(i); }
};

// ---------------------------------------------------------------------------------------
// SVD by one-sided Jacobi on the columns of A (m x n).  As in VNL, W has n entries and V is n x n
// also when m < n (the trailing singular values are then zero and nullvector(), the last column of V,
// spans the null space -- PlaneParametersEstimator.hxx:70-90 relies on that).  Singular values descending.
// ---------------------------------------------------------------------------------------
template <class T>
class vnl_svd {
 public:
  explicit vnl_svd(vnl_matrix<T> const& A) { compute(A); }
  void zero_out_absolute(double tol = 1e-8) {
    rank_ = static_cast<unsigned>(W_.size());
    for (unsigned k = 0; k < W_.size(); k++) {
      if (std::fabs(W_[k]) <= tol) { W_[k] = 0; Winv_[k] = 0; --rank_; }
      else Winv_[k] = T(1) / W_[k];
    }
  }
  unsigned rank() const { return rank_; }
  vnl_matrix<T> const& U() const { return U_; }
  vnl_matrix<T> const& V() const { return V_; }
  T W(unsigned i) const { return W_[i]; }
  vnl_vector<T> nullvector() const { return V_.get_column(V_.cols() - 1); }
  vnl_matrix<T> pinverse() const {
    const unsigned n = V_.rows(), m = U_.rows(), r = static_cast<unsigned>(W_.size());
    vnl_matrix<T> P(n, m, T(0));
    for (unsigned i = 0; i < n; i++) for (unsigned j = 0; j < m; j++) { T s(0); for (unsigned k = 0; k < r; k++) s += V_(i, k) * Winv_[k] * U_(j, k); P(i, j) = s; }
    return P;
  }
  // x = V W^-1 U^T b without forming the (n x m) pseudo-inverse (m can be millions).
  vnl_vector<T> solve(vnl_vector<T> const& b) const {
    const unsigned n = V_.rows(), m = U_.rows(), r = static_cast<unsigned>(W_.size());
    std::vector<T> y(r, T(0));
    for (unsigned k = 0; k < r; k++) { T s(0); for (unsigned j = 0; j < m; j++) s += U_(j, k) * b[j]; y[k] = s * Winv_[k]; }
    vnl_vector<T> x(n, T(0));
    for (unsigned i = 0; i < n; i++) { T s(0); for (unsigned k = 0; k < r; k++) s += V_(i, k) * y[k]; x[i] = s; }
    return x;
  }

 private:
  void compute(vnl_matrix<T> const& Ain) {
    const bool tr = false;
    vnl_matrix<T> A = Ain;
    const unsigned m = A.rows(), n = A.cols();
    vnl_matrix<T> V(n, n, T(0));
    for (unsigned i = 0; i < n; i++) V(i, i) = T(1);
    for (int sweep = 0; sweep < 60; sweep++) {
      bool rotated = false;
      for (unsigned p = 0; p < n; p++) {
        for (unsigned q = p + 1; q < n; q++) {
          T alpha(0), beta(0), gamma(0);
          for (unsigned k = 0; k < m; k++) { alpha += A(k, p) * A(k, p); beta += A(k, q) * A(k, q); gamma += A(k, p) * A(k, q); }
          if (gamma == T(0) || std::fabs(gamma) <= T(1e-16) * std::sqrt(alpha * beta)) continue;
          rotated = true;
          T zeta = (beta - alpha) / (T(2) * gamma);
          T t = (zeta >= T(0) ? T(1) : T(-1)) / (std::fabs(zeta) + std::sqrt(T(1) + zeta * zeta));
          T c = T(1) / std::sqrt(T(1) + t * t), s = c * t;
          for (unsigned k = 0; k < m; k++) { T x = A(k, p), y = A(k, q); A(k, p) = c * x - s * y; A(k, q) = s * x + c * y; }
          for (unsigned k = 0; k < n; k++) { T x = V(k, p), y = V(k, q); V(k, p) = c * x - s * y; V(k, q) = s * x + c * y; }
        }
      }
      if (!rotated) break;
    }
    std::vector<T> w(n);
    for (unsigned j = 0; j < n; j++) { T s(0); for (unsigned k = 0; k < m; k++) s += A(k, j) * A(k, j); w[j] = std::sqrt(s); }
    std::vector<unsigned> order(n);
    for (unsigned j = 0; j < n; j++) order[j] = j;
    std::stable_sort(order.begin(), order.end(), [&](unsigned a, unsigned b) { return w[a] > w[b]; });
    vnl_matrix<T> U(m, n, T(0)), Vs(n, n);
    W_.assign(n, T(0)); Winv_.assign(n, T(0));
    for (unsigned jj = 0; jj < n; jj++) {
      unsigned j = order[jj];
      W_[jj] = w[j];
      for (unsigned k = 0; k < m; k++) U(k, jj) = (w[j] > T(0)) ? A(k, j) / w[j] : T(0);
      for (unsigned k = 0; k < n; k++) Vs(k, jj) = V(k, j);
    }
    if (tr) { U_ = Vs; V_ = U; } else { U_ = U; V_ = Vs; }
    rank_ = n;
    for (unsigned k = 0; k < n; k++) Winv_[k] = (W_[k] != T(0)) ? T(1) / W_[k] : T(0);
  }
  vnl_matrix<T> U_, V_;
  std::vector<T> W_, Winv_;
  unsigned rank_;
};

template <class T>
class vnl_matrix_inverse : public vnl_svd<T> {
 public:
  explicit vnl_matrix_inverse(vnl_matrix<T> const& M) : vnl_svd<T>(M) {}
  operator vnl_matrix<T>() const { return this->pinverse(); }
};
template <class T> inline vnl_vector<T> operator*(vnl_matrix_inverse<T> const& i, vnl_vector<T> const& b) { return i.solve(b); }
template <class T> inline vnl_matrix<T> operator*(vnl_matrix_inverse<T> const& i, vnl_matrix<T> const& b) { return i.pinverse() * b; }

// ---------------------------------------------------------------------------------------
// Least-squares function + Levenberg-Marquardt
// ---------------------------------------------------------------------------------------
class vnl_least_squares_function {
 public:
  enum UseGradient { no_gradient, use_gradient };
  vnl_least_squares_function(unsigned nu, unsigned nr, UseGradient g = use_gradient) : p_(nu), n_(nr), use_gradient_(g == use_gradient) {}
  virtual ~vnl_least_squares_function() {}
  virtual void f(vnl_vector<double> const& x, vnl_vector<double>& fx) = 0;
  virtual void gradf(vnl_vector<double> const& /*x*/, vnl_matrix<double>& /*J*/) {}
  unsigned get_number_of_unknowns() const { return p_; }
  unsigned get_number_of_residuals() const { return n_; }
  bool has_gradient() const { return use_gradient_; }
 protected:
  unsigned p_, n_;
  bool use_gradient_;
};

// vnl_levenberg_marquardt: VNL hands the problem to netlib's MINPACK.  With an analytic gradient (every functor of the
// reference) that is lmder with mode 1 (internal scaling), factor 100; without one, lmdif (forward differences with
// epsfcn = xtol * 0.001).  Defaults of vnl_nonlinear_minimizer: xtol 1e-8, ftol = xtol * 0.01, gtol 1e-5, and
// vnl_levenberg_marquardt sets maxfev = 400 * unknowns.  minimize() is true iff MINPACK's info is 1..4.
// The MINPACK algorithm itself is restated in oracle/minpack_lm.h (VNL / netlib are absent from this image).
class vnl_levenberg_marquardt {
 public:
  explicit vnl_levenberg_marquardt(vnl_least_squares_function& f)
      : f_(&f), xtol_(1e-8), ftol_(1e-8 * 0.01), gtol_(1e-5), epsfcn_(1e-8 * 0.001), maxfev_(400 * f.get_number_of_unknowns()), num_evals_(0), info_(0) {}
  void set_x_tolerance(double v) { xtol_ = v; }
  void set_f_tolerance(double v) { ftol_ = v; }
  void set_g_tolerance(double v) { gtol_ = v; }
  void set_epsilon_function(double v) { epsfcn_ = v; }
  void set_max_function_evals(int v) { maxfev_ = v; }
  int get_num_evaluations() const { return num_evals_; }
  int get_failure_code() const { return info_; }
  bool minimize(vnl_vector<double>& x) {
    const int n = (int)f_->get_number_of_unknowns(), m = (int)f_->get_number_of_residuals();
    if (m < n) { info_ = 0; return false; }   // VNL: "Number of unknowns(n) greater than number of data (m)"
    std::vector<double> fvec(m);
    int nfev = 0, njev = 0;
    info_ = mpk_lmder(&vnl_levenberg_marquardt::callback, this, m, n, x.data_block(), fvec.data(), ftol_, xtol_, gtol_, maxfev_, 100.0, &nfev, &njev);
    num_evals_ = nfev;
    return info_ >= 1 && info_ <= 4;
  }

 private:
  static int callback(void* user, int m, int n, const double* x, double* fvec, double* fjac, int ldfjac, int iflag) {
    vnl_levenberg_marquardt* self = static_cast<vnl_levenberg_marquardt*>(user);
    vnl_vector<double> vx(x, (unsigned)n);
    if (iflag == 1) {
      vnl_vector<double> fx((unsigned)m);
      self->f_->f(vx, fx);
      for (int i = 0; i < m; i++) fvec[i] = fx[i];
    } else if (self->f_->has_gradient()) {
      vnl_matrix<double> J((unsigned)m, (unsigned)n);
      self->f_->gradf(vx, J);
      for (int j = 0; j < n; j++) for (int i = 0; i < m; i++) fjac[i + (size_t)ldfjac * j] = J(i, j);
    } else {   // forward differences (MINPACK fdjac2)
      vnl_vector<double> f0((unsigned)m), f1((unsigned)m);
      self->f_->f(vx, f0);
      const double eps = std::sqrt(std::max(self->epsfcn_, 2.220446049250313e-16));
      for (int j = 0; j < n; j++) {
        const double t = vx[j];
        double h = eps * std::fabs(t);
        if (h == 0.0) h = eps;
        vx[j] = t + h;
        self->f_->f(vx, f1);
        vx[j] = t;
        for (int i = 0; i < m; i++) fjac[i + (size_t)ldfjac * j] = (f1[i] - f0[i]) / h;
      }
    }
    return 0;
  }
  vnl_least_squares_function* f_;
  double xtol_, ftol_, gtol_, epsfcn_;
  int maxfev_, num_evals_, info_;
};

// Deterministic stand-in for vnl_random (only tests/examples of the reference use it).
class vnl_random {
 public:
  vnl_random() : s_(0x9E3779B97F4A7C15ull) {}
  explicit vnl_random(unsigned long seed) : s_(seed * 0x9E3779B97F4A7C15ull + 1) {}
  void reseed(unsigned long seed) { s_ = seed * 0x9E3779B97F4A7C15ull + 1; }
  double drand32(double a, double b) { return a + (b - a) * u01(); }
  double drand32() { return u01(); }
  double drand64(double a, double b) { return drand32(a, b); }
  double normal() { double u1 = u01(), u2 = u01(); if (u1 < 1e-300) u1 = 1e-300; return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2); }
 private:
  double u01() { s_ ^= s_ << 13; s_ ^= s_ >> 7; s_ ^= s_ << 17; return double(s_ >> 11) * (1.0 / 9007199254740992.0); }
  unsigned long long s_;
};

#endif  // LSQR_ORACLE_VNL_SHIM_CORE_H
