// ref_shim_eigen.hpp — CPU ORACLE support (test infrastructure, NOT product code).
//
// A stand-in for the tiny slice of Eigen 3 that the reference's arithmetic on the hot path uses, so that the
// reference's own source lines (include/pointcloud.h:126-158,194-288,378-463; include/g2o_tools.h:58-69,105-140,
// 149-183) can be compiled VERBATIM from /root/reference into oracle/_ref/liboracle_refmath.so and compared bit for bit
// with the oracle's restatement (tests/test_oracle_refmath.py).  Eigen itself is not installable here.
// Fixed sizes only, eager evaluation.  The one thing assumed about Eigen is the order in which it adds the three
// products of a fixed-size dot / matrix product and the terms of `A + B + C`: left to right, ((p0 + p1) + p2) —
// Eigen's unrolled redux for size-3 vectors and its left-associated expression trees.
#pragma once
#include <cmath>
#include <cstddef>
#include <tuple>

namespace Eigen {
typedef std::ptrdiff_t Index;
enum { ColMajor = 0, Dynamic = -1 };

template <class T, int R, int C> struct Matrix;

template <class T, int R, int C> struct CommaInit {
    Matrix<T, R, C> &m;
    int k;
    CommaInit &operator,(const T &v) { m.d[k++] = v; return *this; }
};

template <class T, int R, int C = 1>
struct Matrix {
    T d[R * C];  // row-major storage; only element access is observable
    Matrix() { for (int i = 0; i < R * C; ++i) d[i] = T(0.0); }
    Matrix(const T &a, const T &b, const T &c) { static_assert(R * C == 3, "3-vector"); d[0] = a; d[1] = b; d[2] = c; }
    T &operator()(int i, int j) { return d[i * C + j]; }
    const T &operator()(int i, int j) const { return d[i * C + j]; }
    T &operator()(int i) { return d[i]; }
    const T &operator()(int i) const { return d[i]; }
    T &operator[](int i) { return d[i]; }
    const T &operator[](int i) const { return d[i]; }
    void setZero() { for (int i = 0; i < R * C; ++i) d[i] = T(0.0); }
    static Matrix Zero() { return Matrix(); }
    static Matrix Identity() { Matrix m; for (int i = 0; i < (R < C ? R : C); ++i) m(i, i) = T(1.0); return m; }
    T maxCoeff() const { T m = d[0]; for (int i = 1; i < R * C; ++i) if (d[i] > m) m = d[i]; return m; }
    Matrix &operator/=(const T &s) { for (int i = 0; i < R * C; ++i) d[i] = d[i] / s; return *this; }
    Matrix &operator*=(const T &s) { for (int i = 0; i < R * C; ++i) d[i] = d[i] * s; return *this; }
    CommaInit<T, R, C> operator<<(const T &v) { d[0] = v; return CommaInit<T, R, C>{*this, 1}; }
    T dot(const Matrix &o) const { static_assert(R * C == 3, "3-vector"); return (d[0] * o.d[0] + d[1] * o.d[1]) + d[2] * o.d[2]; }
    Matrix cross(const Matrix &o) const {
        static_assert(R * C == 3, "3-vector");
        return Matrix(d[1] * o.d[2] - d[2] * o.d[1], d[2] * o.d[0] - d[0] * o.d[2], d[0] * o.d[1] - d[1] * o.d[0]);
    }
    T squaredNorm() const { static_assert(R * C == 3, "3-vector"); return (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]; }
    T norm() const { using std::sqrt; return sqrt(squaredNorm()); }
    Matrix<T, C, R> transpose() const { Matrix<T, C, R> t; for (int i = 0; i < R; ++i) for (int j = 0; j < C; ++j) t(j, i) = (*this)(i, j); return t; }
};

template <class T, int R, int C> Matrix<T, R, C> operator+(const Matrix<T, R, C> &a, const Matrix<T, R, C> &b) { Matrix<T, R, C> o; for (int i = 0; i < R * C; ++i) o.d[i] = a.d[i] + b.d[i]; return o; }
template <class T, int R, int C> Matrix<T, R, C> operator-(const Matrix<T, R, C> &a, const Matrix<T, R, C> &b) { Matrix<T, R, C> o; for (int i = 0; i < R * C; ++i) o.d[i] = a.d[i] - b.d[i]; return o; }
template <class T, int R, int C> Matrix<T, R, C> operator-(const Matrix<T, R, C> &a) { Matrix<T, R, C> o; for (int i = 0; i < R * C; ++i) o.d[i] = -a.d[i]; return o; }
template <class T, int R, int C> Matrix<T, R, C> operator*(const T &s, const Matrix<T, R, C> &a) { Matrix<T, R, C> o; for (int i = 0; i < R * C; ++i) o.d[i] = s * a.d[i]; return o; }
template <class T, int R, int C> Matrix<T, R, C> operator*(const Matrix<T, R, C> &a, const T &s) { Matrix<T, R, C> o; for (int i = 0; i < R * C; ++i) o.d[i] = a.d[i] * s; return o; }
template <class T, int R, int C> Matrix<T, R, C> operator/(const Matrix<T, R, C> &a, const T &s) { Matrix<T, R, C> o; for (int i = 0; i < R * C; ++i) o.d[i] = a.d[i] / s; return o; }
// 3x3 * 3x3 and 3x3 * 3x1: coefficient-based product, sum of three products left to right
template <class T, int K> Matrix<T, 3, K> operator*(const Matrix<T, 3, 3> &a, const Matrix<T, 3, K> &b) {
    Matrix<T, 3, K> o;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < K; ++j) o(i, j) = (a(i, 0) * b(0, j) + a(i, 1) * b(1, j)) + a(i, 2) * b(2, j);
    return o;
}

typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 3, 3> Matrix3d;
}  // namespace Eigen

namespace g2o {  // the aliases of g2o/core/eigen_types.h that g2o_tools.h uses
template <int N, typename T = double> using VectorN = Eigen::Matrix<T, N, 1>;
template <int N, typename T = double> using MatrixN = Eigen::Matrix<T, N, N>;
}  // namespace g2o
