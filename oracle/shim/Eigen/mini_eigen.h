// ORACLE SHIM — TEST INFRASTRUCTURE ONLY.
//
// A minimal stand-in for the parts of Eigen 3.3.7 (Ubuntu 20.04 libeigen3-dev, the version ros:noetic-perception ships:
// /root/reference/Dockerfile:1,6) that the reference's ndt_representation / radar_preprocessing sources use, so that
// those sources compile UNMODIFIED from /root/reference into oracle/_ref/libref.so (oracle/Makefile, target `ref`).
// Eigen itself is not in this image and there is no network.
//
// What is the reference's and what is restated here:
//   * every statement, type, conversion, loop and operation ORDER of the compiled code is the reference's own source;
//   * the arithmetic behind Eigen's operators is restated below with Eigen 3.3.7's evaluation semantics for fixed-size
//     float/double matrices on baseline x86-64 (SSE2, no FMA: the reference builds RelWithDebInfo without -march, R/CMakeLists.txt:4-6):
//       - coefficient-wise expressions evaluate every coefficient with the scalar operations in source order;
//       - small fixed-size products are coefficient-based, each coefficient = (lhs.row(i).cwiseProduct(rhs.col(j))).sum(), and
//         sum() over n <= 4 unvectorised terms is Eigen's unrolled half-split reduction  (redux_novec_unroller):
//         n = 2: x0 + x1, n = 3: x0 + (x1 + x2), n = 4: (x0 + x1) + (x2 + x3);  nested products evaluate the inner product into a
//         temporary first (EvalBeforeNestingBit);
//       - scalars of another arithmetic type are converted to the matrix scalar first (promote_scalar_arg);
//       - inverse()/determinant() of 2x2 and 3x3 follow Inverse_impl.h / Determinant.h (cofactors, bruteforce_det3_helper);
//       - SelfAdjointEigenSolver<2x2> follows SelfAdjointEigenSolver::compute (scale, tridiagonal QR step with Wilkinson shift,
//         ascending sort), JacobiSVD follows the two-sided Jacobi sweep of JacobiSVD::compute with real_2x2_jacobi_svd, and
//         Transform::rotation() is computeRotationScaling (SVD polar factor) as in Geometry/Transform.h.
//     These were restated from the published 3.3.7 sources from memory; they cannot be diffed against Eigen here.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <iostream>
#include <complex>
#include <cstdlib>
#include <limits>
#include <map>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW

namespace Eigen {

typedef std::ptrdiff_t Index;
enum TransformTraits { Isometry = 0x1, Affine = 0x2, AffineCompact = 0x10 | Affine, Projective = 0x20 };
enum DecompositionOptions { ComputeFullU = 0x04, ComputeThinU = 0x08, ComputeFullV = 0x10, ComputeThinV = 0x20 };

namespace internal {
// Eigen's completely unrolled, unvectorised reduction (Redux.h, redux_novec_unroller): func(first half, second half)
template <typename T, typename F>
inline T redux_sum(int start, int length, const F& f) {
  if (length == 1) return f(start);
  const int half = length / 2;
  const T a = redux_sum<T>(start, half, f);
  const T b = redux_sum<T>(start + half, length - half, f);
  return a + b;
}
}  // namespace internal

template <typename T, int R, int C> class Matrix;
template <typename M, int BR, int BC> class BlockRef;
template <typename M> class CommaInit;

// a 1x1 matrix converts to its scalar (as Eigen's inner products do)
template <typename Derived, typename T, int R, int C> struct scalar_conversion {};
template <typename Derived, typename T> struct scalar_conversion<Derived, T, 1, 1> {
  operator T() const { return static_cast<const Derived*>(this)->d[0]; }
};

template <typename T, int R, int C>
class Matrix : public scalar_conversion<Matrix<T, R, C>, T, R, C> {
 public:
  typedef T Scalar;
  enum { RowsAtCompileTime = R, ColsAtCompileTime = C, SizeAtCompileTime = R * C };
  T d[R * C];   // column-major

  Matrix() {}
  Matrix(const T& x, const T& y) { static_assert(R * C == 2, "2-vector"); d[0] = x; d[1] = y; }
  Matrix(const T& x, const T& y, const T& z) { static_assert(R * C == 3, "3-vector"); d[0] = x; d[1] = y; d[2] = z; }
  Matrix(const T& x, const T& y, const T& z, const T& w) { static_assert(R * C == 4 && (R == 1 || C == 1), "4-vector"); d[0] = x; d[1] = y; d[2] = z; d[3] = w; }
  template <typename M> Matrix(const BlockRef<M, R, C>& b) { for (int j = 0; j < C; ++j) for (int i = 0; i < R; ++i) (*this)(i, j) = b(i, j); }
  template <typename M> Matrix& operator=(const BlockRef<M, R, C>& b) { Matrix t(b); *this = t; return *this; }

  Index rows() const { return R; }
  Index cols() const { return C; }
  Index size() const { return R * C; }
  T* data() { return d; }
  const T* data() const { return d; }
  T& operator()(Index i, Index j) { return d[j * R + i]; }
  const T& operator()(Index i, Index j) const { return d[j * R + i]; }
  T& operator()(Index i) { static_assert(R == 1 || C == 1, "vector"); return d[i]; }
  const T& operator()(Index i) const { static_assert(R == 1 || C == 1, "vector"); return d[i]; }
  T& operator[](Index i) { return d[i]; }
  const T& operator[](Index i) const { return d[i]; }
  T& coeffRef(Index i, Index j) { return (*this)(i, j); }
  const T& coeff(Index i, Index j) const { return (*this)(i, j); }
  T& coeffRef(Index i) { return d[i]; }
  const T& coeff(Index i) const { return d[i]; }
  T& x() { return d[0]; } T& y() { return d[1]; } T& z() { return d[2]; }
  const T& x() const { return d[0]; } const T& y() const { return d[1]; } const T& z() const { return d[2]; }

  static Matrix Zero() { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = T(0); return m; }
  static Matrix Identity() { Matrix m = Zero(); for (int i = 0; i < (R < C ? R : C); ++i) m(i, i) = T(1); return m; }
  static Matrix Constant(const T& v) { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = v; return m; }
  Matrix& setZero() { for (int i = 0; i < R * C; ++i) d[i] = T(0); return *this; }
  Matrix& setIdentity() { *this = Identity(); return *this; }

  template <int BR, int BC> BlockRef<Matrix, BR, BC> block(Index r0, Index c0) { return BlockRef<Matrix, BR, BC>(*this, r0, c0); }
  template <int BR, int BC> BlockRef<const Matrix, BR, BC> block(Index r0, Index c0) const { return BlockRef<const Matrix, BR, BC>(*this, r0, c0); }
  BlockRef<Matrix, R, 1> col(Index j) { return BlockRef<Matrix, R, 1>(*this, 0, j); }
  BlockRef<const Matrix, R, 1> col(Index j) const { return BlockRef<const Matrix, R, 1>(*this, 0, j); }
  BlockRef<Matrix, 1, C> row(Index i) { return BlockRef<Matrix, 1, C>(*this, i, 0); }
  BlockRef<const Matrix, 1, C> row(Index i) const { return BlockRef<const Matrix, 1, C>(*this, i, 0); }

  Matrix<T, C, R> transpose() const { Matrix<T, C, R> t; for (int j = 0; j < C; ++j) for (int i = 0; i < R; ++i) t(j, i) = (*this)(i, j); return t; }
  Matrix<T, C, R> adjoint() const { return transpose(); }
  const Matrix& real() const { return *this; }
  const Matrix& eval() const { return *this; }
  template <typename U> Matrix<U, R, C> cast() const { Matrix<U, R, C> m; for (int i = 0; i < R * C; ++i) m.d[i] = static_cast<U>(d[i]); return m; }

  T sum() const { return internal::redux_sum<T>(0, R * C, [&](int k) { return d[k]; }); }
  T squaredNorm() const { return internal::redux_sum<T>(0, R * C, [&](int k) { return d[k] * d[k]; }); }   // cwiseAbs2().sum()
  T norm() const { return std::sqrt(squaredNorm()); }
  T determinant() const;
  Matrix inverse() const;

  Matrix& operator+=(const Matrix& o) { for (int i = 0; i < R * C; ++i) d[i] = d[i] + o.d[i]; return *this; }
  Matrix& operator-=(const Matrix& o) { for (int i = 0; i < R * C; ++i) d[i] = d[i] - o.d[i]; return *this; }
  template <typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
  Matrix& operator*=(const S& s) { const T t = static_cast<T>(s); for (int i = 0; i < R * C; ++i) d[i] = d[i] * t; return *this; }
  template <typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
  Matrix& operator/=(const S& s) { const T t = static_cast<T>(s); for (int i = 0; i < R * C; ++i) d[i] = d[i] / t; return *this; }
  Matrix operator-() const { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = -d[i]; return m; }

  CommaInit<Matrix> operator<<(const T& v);
};

// ---- comma initialiser (row-major fill, like Eigen) ----
template <typename M>
class CommaInit {
 public:
  CommaInit(M& m, const typename M::Scalar& v) : m_(m), k_(0) { put(v); }
  CommaInit& operator,(const typename M::Scalar& v) { put(v); return *this; }
 private:
  void put(const typename M::Scalar& v) {
    const int r = k_ / M::ColsAtCompileTime, c = k_ % M::ColsAtCompileTime;
    m_(r, c) = v; ++k_;
  }
  M& m_; int k_;
};
template <typename T, int R, int C>
CommaInit<Matrix<T, R, C>> Matrix<T, R, C>::operator<<(const T& v) { return CommaInit<Matrix<T, R, C>>(*this, v); }

// ---- lvalue block proxy ----
template <typename M, int BR, int BC>
class BlockRef {
 public:
  typedef typename std::remove_const<M>::type Plain;
  typedef typename Plain::Scalar T;
  typedef Matrix<T, BR, BC> Value;
  BlockRef(M& m, Index r0, Index c0) : m_(m), r0_(r0), c0_(c0) {}
  const T& operator()(Index i, Index j) const { return m_(r0_ + i, c0_ + j); }
  T& ref(Index i, Index j) const { return const_cast<Plain&>(m_)(r0_ + i, c0_ + j); }
  Value value() const { Value v; for (int j = 0; j < BC; ++j) for (int i = 0; i < BR; ++i) v(i, j) = (*this)(i, j); return v; }
  BlockRef& operator=(const Value& v) { for (int j = 0; j < BC; ++j) for (int i = 0; i < BR; ++i) ref(i, j) = v(i, j); return *this; }
  BlockRef& operator=(const BlockRef& o) { return *this = o.value(); }
  template <typename M2> BlockRef& operator=(const BlockRef<M2, BR, BC>& o) { return *this = o.value(); }
  template <typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
  BlockRef& operator/=(const S& s) { const T t = static_cast<T>(s); for (int j = 0; j < BC; ++j) for (int i = 0; i < BR; ++i) ref(i, j) = (*this)(i, j) / t; return *this; }
  template <typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
  BlockRef& operator*=(const S& s) { const T t = static_cast<T>(s); for (int j = 0; j < BC; ++j) for (int i = 0; i < BR; ++i) ref(i, j) = (*this)(i, j) * t; return *this; }
  T determinant() const { return value().determinant(); }
  Value inverse() const { return value().inverse(); }
  Matrix<T, BC, BR> transpose() const { return value().transpose(); }
  T norm() const { return value().norm(); }
  template <typename U> Matrix<U, BR, BC> cast() const { return value().template cast<U>(); }
 private:
  M& m_; Index r0_, c0_;
};

// ---- coefficient-wise binary operators ----
template <typename T, int R, int C>
Matrix<T, R, C> operator+(const Matrix<T, R, C>& a, const Matrix<T, R, C>& b) { Matrix<T, R, C> m; for (int i = 0; i < R * C; ++i) m.d[i] = a.d[i] + b.d[i]; return m; }
template <typename T, int R, int C>
Matrix<T, R, C> operator-(const Matrix<T, R, C>& a, const Matrix<T, R, C>& b) { Matrix<T, R, C> m; for (int i = 0; i < R * C; ++i) m.d[i] = a.d[i] - b.d[i]; return m; }
template <typename T, int R, int C, typename M2> Matrix<T, R, C> operator+(const Matrix<T, R, C>& a, const BlockRef<M2, R, C>& b) { return a + b.value(); }
template <typename T, int R, int C, typename M2> Matrix<T, R, C> operator+(const BlockRef<M2, R, C>& a, const Matrix<T, R, C>& b) { return a.value() + b; }
template <typename T, int R, int C, typename M2> Matrix<T, R, C> operator-(const Matrix<T, R, C>& a, const BlockRef<M2, R, C>& b) { return a - b.value(); }
template <typename T, int R, int C, typename M2> Matrix<T, R, C> operator-(const BlockRef<M2, R, C>& a, const Matrix<T, R, C>& b) { return a.value() - b; }
template <typename M1, typename M2, int R, int C>
typename BlockRef<M1, R, C>::Value operator-(const BlockRef<M1, R, C>& a, const BlockRef<M2, R, C>& b) { return a.value() - b.value(); }
template <typename M1, typename M2, int R, int C>
typename BlockRef<M1, R, C>::Value operator+(const BlockRef<M1, R, C>& a, const BlockRef<M2, R, C>& b) { return a.value() + b.value(); }

// scalar (any arithmetic type, converted to the matrix scalar first) x matrix
template <typename S, typename T, int R, int C, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
Matrix<T, R, C> operator*(const S& s, const Matrix<T, R, C>& a) { const T t = static_cast<T>(s); Matrix<T, R, C> m; for (int i = 0; i < R * C; ++i) m.d[i] = t * a.d[i]; return m; }
template <typename S, typename T, int R, int C, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
Matrix<T, R, C> operator*(const Matrix<T, R, C>& a, const S& s) { const T t = static_cast<T>(s); Matrix<T, R, C> m; for (int i = 0; i < R * C; ++i) m.d[i] = a.d[i] * t; return m; }
template <typename S, typename T, int R, int C, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
Matrix<T, R, C> operator/(const Matrix<T, R, C>& a, const S& s) { const T t = static_cast<T>(s); Matrix<T, R, C> m; for (int i = 0; i < R * C; ++i) m.d[i] = a.d[i] / t; return m; }

// matrix product: coefficient-based, Eigen's half-split reduction over the inner dimension
template <typename T, int R, int K, int C>
Matrix<T, R, C> operator*(const Matrix<T, R, K>& a, const Matrix<T, K, C>& b) {
  Matrix<T, R, C> m;
  for (int j = 0; j < C; ++j)
    for (int i = 0; i < R; ++i) m(i, j) = internal::redux_sum<T>(0, K, [&](int k) { return a(i, k) * b(k, j); });
  return m;
}
template <typename T, int R, int K, int C, typename M2> Matrix<T, R, C> operator*(const Matrix<T, R, K>& a, const BlockRef<M2, K, C>& b) { return a * b.value(); }
template <typename T, int R, int K, int C, typename M2> Matrix<T, R, C> operator*(const BlockRef<M2, R, K>& a, const Matrix<T, K, C>& b) { return a.value() * b; }

// ---- determinant / inverse (Determinant.h, Inverse_impl.h) ----
namespace internal {
template <typename M> inline typename M::Scalar det3_helper(const M& m, int a, int b, int c) {
  return m(0, a) * (m(1, b) * m(2, c) - m(1, c) * m(2, b));
}
template <typename M, int i, int j> inline typename M::Scalar cofactor_3x3(const M& m) {
  enum { i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3 };
  return m(i1, j1) * m(i2, j2) - m(i1, j2) * m(i2, j1);
}
}  // namespace internal
template <typename T, int R, int C>
T Matrix<T, R, C>::determinant() const {
  // Eigen dispatches on RowsAtCompileTime only; a non-square fixed-size matrix asserts at run time (debug builds only)
  const Matrix& m = *this;
  if (R == 1) return m.d[0];
  if (R == 2) { assert(C >= 2 && "determinant of a non-square matrix"); if (C < 2) return T(0); return m.d[0] * m.d[1 * R + 1] - m.d[1] * m.d[1 * R + 0]; }
  if (R == 3) {
    assert(C == 3);
    if (C < 3) return T(0);
    return internal::det3_helper(m, 0, 1, 2) - internal::det3_helper(m, 1, 0, 2) + internal::det3_helper(m, 2, 0, 1);
  }
  assert(false && "mini_eigen: determinant() only up to 3x3");
  return T(0);
}
namespace internal {
template <typename T> inline Matrix<T, 2, 2> inverse2(const Matrix<T, 2, 2>& m) {
  const T invdet = T(1) / m.determinant();
  Matrix<T, 2, 2> r;
  r(0, 0) = m(1, 1) * invdet;
  r(1, 0) = -m(1, 0) * invdet;
  r(0, 1) = -m(0, 1) * invdet;
  r(1, 1) = m(0, 0) * invdet;
  return r;
}
template <typename T> inline Matrix<T, 3, 3> inverse3(const Matrix<T, 3, 3>& m) {
  typedef Matrix<T, 3, 3> M;
  Matrix<T, 3, 1> c0;
  c0(0) = cofactor_3x3<M, 0, 0>(m); c0(1) = cofactor_3x3<M, 1, 0>(m); c0(2) = cofactor_3x3<M, 2, 0>(m);
  const T det = redux_sum<T>(0, 3, [&](int k) { return c0(k) * m(k, 0); });   // (cofactors_col0.cwiseProduct(matrix.col(0))).sum()
  const T invdet = T(1) / det;
  M r;
  r(0, 0) = c0(0) * invdet; r(0, 1) = c0(1) * invdet; r(0, 2) = c0(2) * invdet;   // result.row(0) = cofactors_col0 * invdet
  r(1, 0) = cofactor_3x3<M, 0, 1>(m) * invdet;
  r(1, 1) = cofactor_3x3<M, 1, 1>(m) * invdet;
  r(1, 2) = cofactor_3x3<M, 2, 1>(m) * invdet;
  r(2, 0) = cofactor_3x3<M, 0, 2>(m) * invdet;
  r(2, 1) = cofactor_3x3<M, 1, 2>(m) * invdet;
  r(2, 2) = cofactor_3x3<M, 2, 2>(m) * invdet;
  return r;
}
template <typename T, int R, int C> struct inverse_impl { static Matrix<T, R, C> run(const Matrix<T, R, C>&) { static_assert(R == C && R <= 3, "mini_eigen: inverse() only up to 3x3"); return Matrix<T, R, C>(); } };
template <typename T> struct inverse_impl<T, 1, 1> { static Matrix<T, 1, 1> run(const Matrix<T, 1, 1>& m) { Matrix<T, 1, 1> r; r.d[0] = T(1) / m.d[0]; return r; } };
template <typename T> struct inverse_impl<T, 2, 2> { static Matrix<T, 2, 2> run(const Matrix<T, 2, 2>& m) { return inverse2(m); } };
template <typename T> struct inverse_impl<T, 3, 3> { static Matrix<T, 3, 3> run(const Matrix<T, 3, 3>& m) { return inverse3(m); } };
}  // namespace internal
template <typename T, int R, int C>
Matrix<T, R, C> Matrix<T, R, C>::inverse() const { return internal::inverse_impl<T, R, C>::run(*this); }

typedef Matrix<float, 2, 1> Vector2f;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<float, 4, 1> Vector4f;
typedef Matrix<float, 2, 2> Matrix2f;
typedef Matrix<float, 3, 3> Matrix3f;
typedef Matrix<float, 4, 4> Matrix4f;
typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 2, 2> Matrix2d;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 4, 4> Matrix4d;

// ---- Jacobi rotations (Jacobi.h) ----
template <typename T>
struct JacobiRotation {
  T c_, s_;
  JacobiRotation() {}
  JacobiRotation(const T& c, const T& s) : c_(c), s_(s) {}
  T& c() { return c_; } T& s() { return s_; }
  const T& c() const { return c_; } const T& s() const { return s_; }
  JacobiRotation operator*(const JacobiRotation& o) const { return JacobiRotation(c_ * o.c_ - s_ * o.s_, c_ * o.s_ + s_ * o.c_); }
  JacobiRotation transpose() const { return JacobiRotation(c_, -s_); }
  JacobiRotation adjoint() const { return JacobiRotation(c_, -s_); }
  // makeJacobi(x, y, z): rotation diagonalising [[x, y], [y, z]]
  bool makeJacobi(const T& x, const T& y, const T& z) {
    const T deno = T(2) * std::abs(y);
    if (deno < (std::numeric_limits<T>::min)()) { c_ = T(1); s_ = T(0); return false; }
    const T tau = (x - z) / deno;
    const T w = std::sqrt(tau * tau + T(1));
    T t;
    if (tau > T(0)) t = T(1) / (tau + w); else t = T(1) / (tau - w);
    const T sign_t = t > T(0) ? T(1) : T(-1);
    const T n = T(1) / std::sqrt(t * t + T(1));
    s_ = -sign_t * (y / std::abs(y)) * std::abs(t) * n;
    c_ = n;
    return true;
  }
  // makeGivens(p, q): G^T (p, q)^T = (r, 0)^T   (real case of Jacobi.h)
  void makeGivens(const T& p, const T& q) {
    if (q == T(0)) { c_ = p < T(0) ? T(-1) : T(1); s_ = T(0); }
    else if (p == T(0)) { c_ = T(0); s_ = q < T(0) ? T(1) : T(-1); }
    else if (std::abs(p) > std::abs(q)) {
      const T t = q / p; T u = std::sqrt(T(1) + t * t); if (p < T(0)) u = -u;
      c_ = T(1) / u; s_ = -t * c_;
    } else {
      const T t = p / q; T u = std::sqrt(T(1) + t * t); if (q < T(0)) u = -u;
      s_ = -T(1) / u; c_ = -t * s_;
    }
  }
};
namespace internal {
// apply_rotation_in_the_plane, scalar path: x_i <- c x_i + s y_i ; y_i <- -s x_i + c y_i
template <typename T, int R, int C> inline void rot_rows(Matrix<T, R, C>& m, int p, int q, const JacobiRotation<T>& j) {   // applyOnTheLeft(p, q, j): rows p, q with j
  const T c = j.c(), s = j.s();
  if (c == T(1) && s == T(0)) return;
  for (int k = 0; k < C; ++k) { const T xi = m(p, k), yi = m(q, k); m(p, k) = c * xi + s * yi; m(q, k) = -s * xi + c * yi; }
}
template <typename T, int R, int C> inline void rot_cols(Matrix<T, R, C>& m, int p, int q, const JacobiRotation<T>& j) {   // applyOnTheRight(p, q, j): columns p, q with j.transpose()
  const T c = j.c(), s = j.s();
  if (c == T(1) && s == T(0)) return;
  for (int k = 0; k < R; ++k) { const T xi = m(k, p), yi = m(k, q); m(k, p) = c * xi + (-s) * yi; m(k, q) = s * xi + c * yi; }
}
}  // namespace internal

// ---- SelfAdjointEigenSolver (Eigenvalues/SelfAdjointEigenSolver.h: compute(), tridiagonal_qr_step) — 2x2 only ----
template <typename MatrixType> class SelfAdjointEigenSolver;
template <typename T>
class SelfAdjointEigenSolver<Matrix<T, 2, 2>> {
 public:
  typedef Matrix<T, 2, 2> M;
  typedef Matrix<T, 2, 1> V;
  explicit SelfAdjointEigenSolver(const M& a) { compute(a); }
  const V& eigenvalues() const { return ev_; }
  const M& eigenvectors() const { return vec_; }
 private:
  void compute(const M& a) {
    // lower triangle only; scale by the largest |coefficient|
    T scale = std::max(std::max(std::abs(a(0, 0)), std::abs(a(1, 0))), std::abs(a(1, 1)));
    if (scale == T(0)) scale = T(1);
    T d0 = a(0, 0) / scale, d1 = a(1, 1) / scale, e = a(1, 0) / scale;   // tridiagonalisation of a 2x2 is the identity
    vec_ = M::Identity();
    const T considerAsZero = (std::numeric_limits<T>::min)();
    const T precision = T(2) * std::numeric_limits<T>::epsilon();
    int iter = 0;
    while (true) {
      if (std::abs(e) <= (std::abs(d0) + std::abs(d1)) * precision || std::abs(e) <= considerAsZero) e = T(0);
      if (e == T(0)) break;
      if (++iter > 30 * 2) break;
      // tridiagonal_qr_step on the 2x2 block
      const T td = (d0 - d1) * T(0.5);
      T mu = d1;
      if (td == T(0)) mu -= std::abs(e);
      else {
        const T e2 = e * e;
        const T ax = std::abs(td), ay = std::abs(e);
        T p, qp; if (ax > ay) { p = ax; qp = ay / p; } else { p = ay; qp = ax / p; }
        const T h = p == T(0) ? T(0) : p * std::sqrt(T(1) + qp * qp);    // numext::hypot
        if (e2 == T(0)) mu -= (e / (td + (td > T(0) ? T(1) : T(-1)))) * (e / h);
        else mu -= e2 / (td + (td > T(0) ? h : -h));
      }
      JacobiRotation<T> rot;
      rot.makeGivens(d0 - mu, e);
      const T c = rot.c(), s = rot.s();
      const T sdk = s * d0 + c * e;
      const T dkp1 = s * e + c * d1;
      const T nd0 = c * (c * d0 - s * e) - s * (c * e - s * d1);
      const T nd1 = s * sdk + c * dkp1;
      const T ne = c * sdk - s * dkp1;
      d0 = nd0; d1 = nd1; e = ne;
      internal::rot_cols(vec_, 0, 1, rot);
    }
    if (d1 < d0) { std::swap(d0, d1); std::swap(vec_(0, 0), vec_(0, 1)); std::swap(vec_(1, 0), vec_(1, 1)); }
    ev_(0) = d0 * scale; ev_(1) = d1 * scale;
  }
  V ev_; M vec_;
};

// ---- JacobiSVD (SVD/JacobiSVD.h: two-sided Jacobi, real square case, full U and V) ----
template <typename MatrixType> class JacobiSVD;
template <typename T, int N>
class JacobiSVD<Matrix<T, N, N>> {
 public:
  typedef Matrix<T, N, N> M;
  typedef Matrix<T, N, 1> V;
  JacobiSVD(const M& a, unsigned int = ComputeFullU | ComputeFullV) { compute(a); }
  const M& matrixU() const { return U_; }
  const M& matrixV() const { return Vm_; }
  const V& singularValues() const { return sv_; }
 private:
  static void real_2x2_jacobi_svd(const M& w, int p, int q, JacobiRotation<T>* j_left, JacobiRotation<T>* j_right) {
    Matrix<T, 2, 2> m;
    m << w(p, p), w(p, q), w(q, p), w(q, q);
    JacobiRotation<T> rot1;
    const T t = m(0, 0) + m(1, 1);
    const T d = m(1, 0) - m(0, 1);
    if (std::abs(d) < (std::numeric_limits<T>::min)()) { rot1.s() = T(0); rot1.c() = T(1); }
    else {
      const T u = t / d;
      const T tmp = std::sqrt(T(1) + u * u);
      rot1.s() = T(1) / tmp;
      rot1.c() = u / tmp;
    }
    internal::rot_rows(m, 0, 1, rot1);
    j_right->makeJacobi(m(0, 0), m(0, 1), m(1, 1));
    *j_left = rot1 * j_right->transpose();
  }
  void compute(const M& a) {
    const T precision = T(2) * std::numeric_limits<T>::epsilon();
    const T considerAsZero = (std::numeric_limits<T>::min)();
    T scale = T(0);
    for (int i = 0; i < N * N; ++i) scale = std::max(scale, std::abs(a.d[i]));
    if (scale == T(0)) scale = T(1);
    M w = a / scale;
    U_ = M::Identity(); Vm_ = M::Identity();
    T maxDiagEntry = T(0);
    for (int i = 0; i < N; ++i) maxDiagEntry = std::max(maxDiagEntry, std::abs(w(i, i)));
    bool finished = false;
    while (!finished) {
      finished = true;
      for (int p = 1; p < N; ++p) {
        for (int q = 0; q < p; ++q) {
          const T threshold = std::max(considerAsZero, precision * maxDiagEntry);
          if (std::abs(w(p, q)) > threshold || std::abs(w(q, p)) > threshold) {
            finished = false;
            JacobiRotation<T> j_left, j_right;
            real_2x2_jacobi_svd(w, p, q, &j_left, &j_right);
            internal::rot_rows(w, p, q, j_left);
            internal::rot_cols(U_, p, q, j_left.transpose());
            internal::rot_cols(w, p, q, j_right);
            internal::rot_cols(Vm_, p, q, j_right);
            maxDiagEntry = std::max(maxDiagEntry, std::max(std::abs(w(p, p)), std::abs(w(q, q))));
          }
        }
      }
    }
    for (int i = 0; i < N; ++i) {
      const T aii = std::abs(w(i, i));
      sv_(i) = aii;
      if (w(i, i) < T(0)) for (int k = 0; k < N; ++k) U_(k, i) = -U_(k, i);     // m_matrixU.col(i) *= m_workMatrix.coeff(i,i) / a   (a = |.|)
    }
    for (int i = 0; i < N; ++i) sv_(i) = sv_(i) * scale;
    for (int i = 0; i < N; ++i) {       // sort singular values in descending order, swap the columns of U and V
      int pos = 0; T best = sv_(i);
      for (int k = 1; k < N - i; ++k) if (sv_(i + k) > best) { best = sv_(i + k); pos = k; }
      if (best == T(0)) break;
      if (pos) {
        pos += i;
        std::swap(sv_(i), sv_(pos));
        for (int k = 0; k < N; ++k) { std::swap(U_(k, pos), U_(k, i)); std::swap(Vm_(k, pos), Vm_(k, i)); }
      }
    }
  }
  M U_, Vm_; V sv_;
};

// ---- Transform (Geometry/Transform.h), Affine mode: (Dim+1)x(Dim+1) homogeneous matrix ----
template <typename T, int Dim, int Mode>
class Transform {
 public:
  typedef Matrix<T, Dim + 1, Dim + 1> MatrixType;
  typedef Matrix<T, Dim, Dim> LinearMatrixType;
  typedef Matrix<T, Dim, 1> VectorType;
  Transform() {}
  explicit Transform(const MatrixType& m) : m_(m) {}
  template <typename M2> explicit Transform(const BlockRef<M2, Dim + 1, Dim + 1>& b) : m_(b.value()) {}
  static Transform Identity() { Transform t; t.m_ = MatrixType::Identity(); return t; }
  MatrixType& matrix() { return m_; }
  const MatrixType& matrix() const { return m_; }
  BlockRef<MatrixType, Dim, Dim> linear() { return BlockRef<MatrixType, Dim, Dim>(m_, 0, 0); }
  BlockRef<const MatrixType, Dim, Dim> linear() const { return BlockRef<const MatrixType, Dim, Dim>(m_, 0, 0); }
  BlockRef<MatrixType, Dim, 1> translation() { return BlockRef<MatrixType, Dim, 1>(m_, 0, Dim); }
  BlockRef<const MatrixType, Dim, 1> translation() const { return BlockRef<const MatrixType, Dim, 1>(m_, 0, Dim); }
  // computeRotationScaling(&rotation, 0): polar factor through the SVD of the linear part
  LinearMatrixType rotation() const {
    const LinearMatrixType lin = linear().value();
    JacobiSVD<LinearMatrixType> svd(lin, ComputeFullU | ComputeFullV);
    const T x = (svd.matrixU() * svd.matrixV().adjoint()).determinant();   // +-1
    LinearMatrixType m(svd.matrixU());
    for (int k = 0; k < Dim; ++k) m(k, 0) = m(k, 0) / x;                   // m.col(0) /= x
    return m * svd.matrixV().adjoint();
  }
  // transform_right_product_impl (Affine, Dim-vector): res = translation; res += linear * v
  VectorType operator*(const VectorType& v) const {
    VectorType res = translation().value();
    const VectorType lv = linear().value() * v;
    for (int i = 0; i < Dim; ++i) res(i) = res(i) + lv(i);
    return res;
  }
  // affine * affine: homogeneous product restricted to the affine part (res.linear = a.linear * b.linear; res.translation = a.linear * b.translation + a.translation)
  Transform operator*(const Transform& o) const {
    Transform r = Identity();
    const LinearMatrixType l = linear().value() * o.linear().value();
    const VectorType lt = linear().value() * o.translation().value();
    for (int i = 0; i < Dim; ++i) { for (int j = 0; j < Dim; ++j) r.m_(i, j) = l(i, j); r.m_(i, Dim) = lt(i) + m_(i, Dim); }
    return r;
  }
  template <typename U> Transform<U, Dim, Mode> cast() const { return Transform<U, Dim, Mode>(m_.template cast<U>()); }
 private:
  MatrixType m_;
};
typedef Transform<float, 2, Affine> Affine2f;
typedef Transform<float, 3, Affine> Affine3f;
typedef Transform<double, 2, Affine> Affine2d;
typedef Transform<double, 3, Affine> Affine3d;

}  // namespace Eigen
