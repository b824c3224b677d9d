// oracle_math.hpp — CPU ORACLE (test infrastructure, NOT product code).
//
// Plain-C++ restatement of the reference's arithmetic for the calibration
// cost-evaluation path.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may build, load or call anything under
// oracle/.  The product (libstlcalib.so) never links or calls it.
//
// Eigen / g2o / Ceres / OpenCV are not installable here, so the
// operation ORDER of every fp64 expression is fixed by this file: strictly
// left-to-right ((a0*b0 + a1*b1) + a2*b2), no FMA contraction (built without
// -march, like the reference: CMakeLists.txt:2,7).  Where the reference goes
// through Eigen's redux/product kernels the true order is unverifiable here
// (SURVEY.md H2); differences are <= 1 ulp per operation.
//
// Every function cites the reference lines it follows (paths relative to
// /root/reference).
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>
#include <cstdint>
#include <limits>

namespace orc {

// ---------------------------------------------------------------- dual numbers
// Forward-mode dual with N partials; mirrors ceres::Jet (value + N-vector) as
// used by AutoDiffCostFunction<...,7> / DynamicAutoDiffCostFunction<...,6>
// (IBACalib2.hpp:207,594,637).  Each partial is propagated independently, so
// the two stride-6 passes of the dynamic variant give the same numbers.
template <int N>
struct Dual {
    double a;
    double v[N];
    Dual() : a(0) { for (int i = 0; i < N; ++i) v[i] = 0; }
    Dual(double x) : a(x) { for (int i = 0; i < N; ++i) v[i] = 0; }  // NOLINT (implicit like Jet)
    static Dual var(double x, int k) { Dual d(x); d.v[k] = 1.0; return d; }
};
template <int N> inline Dual<N> operator+(const Dual<N> &f, const Dual<N> &g) { Dual<N> h; h.a = f.a + g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
template <int N> inline Dual<N> operator-(const Dual<N> &f, const Dual<N> &g) { Dual<N> h; h.a = f.a - g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
template <int N> inline Dual<N> operator-(const Dual<N> &f) { Dual<N> h; h.a = -f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
// Jet: f*g = (f.a*g.a, f.a*g.v + f.v*g.a)
template <int N> inline Dual<N> operator*(const Dual<N> &f, const Dual<N> &g) { Dual<N> h; h.a = f.a * g.a; for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
// Jet: f/g = (f.a/g.a, (f.v - (f.a/g.a)*g.v) * (1/g.a))
template <int N> inline Dual<N> operator/(const Dual<N> &f, const Dual<N> &g) { Dual<N> h; const double gi = 1.0 / g.a; const double q = f.a * gi; h.a = q; for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - q * g.v[i]) * gi; return h; }
template <int N> inline bool operator<(const Dual<N> &f, const Dual<N> &g) { return f.a < g.a; }
template <int N> inline Dual<N> sqrt(const Dual<N> &f) { Dual<N> h; h.a = std::sqrt(f.a); const double t = 1.0 / (2.0 * h.a); for (int i = 0; i < N; ++i) h.v[i] = t * f.v[i]; return h; }
template <int N> inline Dual<N> cos(const Dual<N> &f) { Dual<N> h; h.a = std::cos(f.a); const double t = -std::sin(f.a); for (int i = 0; i < N; ++i) h.v[i] = t * f.v[i]; return h; }
template <int N> inline Dual<N> sin(const Dual<N> &f) { Dual<N> h; h.a = std::sin(f.a); const double t = std::cos(f.a); for (int i = 0; i < N; ++i) h.v[i] = t * f.v[i]; return h; }
// Jet pow(f, double p): (f.a^p, p*f.a^(p-1) * f.v)
template <int N> inline Dual<N> pow(const Dual<N> &f, double p) { Dual<N> h; h.a = std::pow(f.a, p); const double t = p * std::pow(f.a, p - 1.0); for (int i = 0; i < N; ++i) h.v[i] = t * f.v[i]; return h; }
// Jet exp(f): (e^a, e^a * f.v)
template <int N> inline Dual<N> exp(const Dual<N> &f) { Dual<N> h; h.a = std::exp(f.a); for (int i = 0; i < N; ++i) h.v[i] = h.a * f.v[i]; return h; }
inline double exp(double x) { return std::exp(x); }
inline double sqrt(double x) { return std::sqrt(x); }
inline double cos(double x) { return std::cos(x); }
inline double sin(double x) { return std::sin(x); }
inline double pow(double x, double p) { return std::pow(x, p); }
inline double value_of(double x) { return x; }
template <int N> inline double value_of(const Dual<N> &x) { return x.a; }

// ---------------------------------------------------------------- small algebra
template <class T> inline T dot3(const T a[3], const T b[3]) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
template <class T> inline void matvec3(const T M[9], const T p[3], T out[3]) {
    for (int i = 0; i < 3; ++i) out[i] = (M[i * 3] * p[0] + M[i * 3 + 1] * p[1]) + M[i * 3 + 2] * p[2];
}
template <class T> inline void matmul3(const T A[9], const T B[9], T Cm[9]) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Cm[i * 3 + j] = (A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j]) + A[i * 3 + 2] * B[6 + j];
}

// skew (g2o_tools.h:58-69)
template <class T> inline void skew(const T w[3], T O[9]) {
    O[0] = T(0.0); O[1] = -w[2]; O[2] = w[1];
    O[3] = w[2];  O[4] = T(0.0); O[5] = -w[0];
    O[6] = -w[1]; O[7] = w[0];  O[8] = T(0.0);
}

// SE3Exp / Sim3Exp (g2o_tools.h:106-140,150-183): [omega, upsilon(, s)] -> R, t = V*upsilon.
// The scale is NOT exponentiated (g2o_tools.h:138).
template <class T> inline void SE3Exp(const T *x, T R[9], T t[3]) {
    const T w[3] = {x[0], x[1], x[2]}, u[3] = {x[3], x[4], x[5]};
    const T theta = sqrt(dot3(w, w));
    T O[9], O2[9], V[9];
    skew(w, O);
    matmul3(O, O, O2);
    if (value_of(theta) < 1e-4) {
        const T half(0.5), sixth = T(1.0) / T(6.0);
        for (int i = 0; i < 9; ++i) {
            const T I((i % 4 == 0) ? 1.0 : 0.0);
            R[i] = (I + O[i]) + half * O2[i];
            V[i] = (I + half * O[i]) + sixth * O2[i];
        }
    } else {
        const T costh = cos(theta), sinth = sin(theta);
        const T invth2 = pow(theta, -2.0), invth3 = pow(theta, -3.0);
        const T a = sinth / theta, b = (T(1.0) - costh) * invth2, c = (theta - sinth) * invth3;
        for (int i = 0; i < 9; ++i) {
            const T I((i % 4 == 0) ? 1.0 : 0.0);
            R[i] = (I + a * O[i]) + b * O2[i];
            V[i] = (I + b * O[i]) + c * O2[i];
        }
    }
    matvec3(V, u, t);
}
template <class T> inline void Sim3Exp(const T *x, T R[9], T t[3], T &s) { SE3Exp(x, R, t); s = x[6]; }

// Rigid transform [R|t] helpers on plain doubles.
struct Rt { double R[9]; double t[3]; };
// Eigen::Isometry3d * Vector3d (pointcloud.h:85): t_i + ((R_i0 x + R_i1 y) + R_i2 z)
inline void apply(const Rt &T, const double p[3], double q[3]) {
    for (int i = 0; i < 3; ++i) q[i] = T.t[i] + ((T.R[i * 3] * p[0] + T.R[i * 3 + 1] * p[1]) + T.R[i * 3 + 2] * p[2]);
}
// Isometry inverse (iba_global.cpp:234): R^T, -(R^T t)
inline Rt inverse(const Rt &T) {
    Rt I;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) I.R[i * 3 + j] = T.R[j * 3 + i];
    for (int i = 0; i < 3; ++i) I.t[i] = -((I.R[i * 3] * T.t[0] + I.R[i * 3 + 1] * T.t[1]) + I.R[i * 3 + 2] * T.t[2]);
    return I;
}
// 4x4 product restricted to the affine part: (A*B).R = A.R*B.R, (A*B).t = A.R*B.t + A.t
inline Rt compose(const Rt &A, const Rt &B) {
    Rt Cm;
    matmul3(A.R, B.R, Cm.R);
    for (int i = 0; i < 3; ++i) Cm.t[i] = ((A.R[i * 3] * B.t[0] + A.R[i * 3 + 1] * B.t[1]) + A.R[i * 3 + 2] * B.t[2]) + A.t[i];
    return Cm;
}

// SE3Log (g2o_tools.h:78-82) = g2o::SE3Quat(R, t).log().  g2o is a third-party
// dependency absent from /root/reference (README.md:51 pins release
// 20230223_git); restated from its published se3quat.h + Eigen's
// Quaternion(Matrix3) / toRotationMatrix, UNVERIFIED here.
inline void SE3Log(const double Rin[9], const double t[3], double out[6]) {
    // Eigen::Quaterniond(R)
    double q[4];  // x y z w
    double tr = (Rin[0] + Rin[4]) + Rin[8];
    if (tr > 0.0) {
        double s = std::sqrt(tr + 1.0);
        q[3] = 0.5 * s;
        s = 0.5 / s;
        q[0] = (Rin[7] - Rin[5]) * s;
        q[1] = (Rin[2] - Rin[6]) * s;
        q[2] = (Rin[3] - Rin[1]) * s;
    } else {
        int i = 0;
        if (Rin[4] > Rin[0]) i = 1;
        if (Rin[8] > Rin[i * 4]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        double s = std::sqrt(Rin[i * 4] - Rin[j * 4] - Rin[k * 4] + 1.0);
        q[i] = 0.5 * s;
        s = 0.5 / s;
        q[3] = (Rin[k * 3 + j] - Rin[j * 3 + k]) * s;
        q[j] = (Rin[j * 3 + i] + Rin[i * 3 + j]) * s;
        q[k] = (Rin[k * 3 + i] + Rin[i * 3 + k]) * s;
    }
    // SE3Quat::normalizeRotation
    if (q[3] < 0) for (int i = 0; i < 4; ++i) q[i] = -q[i];
    const double n = std::sqrt(((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3]);
    for (int i = 0; i < 4; ++i) q[i] /= n;
    // toRotationMatrix
    const double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
    const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
    const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
    const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
    double R[9];
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
    // SE3Quat::log
    const double d = 0.5 * (((R[0] + R[4]) + R[8]) - 1.0);
    const double dR[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]};
    double w[3], O[9], O2[9], Vinv[9];
    if (std::fabs(d) > 0.99999) {
        for (int i = 0; i < 3; ++i) w[i] = 0.5 * dR[i];
        skew(w, O);
        matmul3(O, O, O2);
        for (int i = 0; i < 9; ++i) Vinv[i] = (((i % 4 == 0) ? 1.0 : 0.0) - 0.5 * O[i]) + (1.0 / 12.0) * O2[i];
    } else {
        const double theta = std::acos(d);
        const double f = theta / (2 * std::sqrt(1 - d * d));
        for (int i = 0; i < 3; ++i) w[i] = f * dR[i];
        skew(w, O);
        matmul3(O, O, O2);
        const double c = (1 - theta / (2 * std::tan(theta / 2))) / (theta * theta);
        for (int i = 0; i < 9; ++i) Vinv[i] = (((i % 4 == 0) ? 1.0 : 0.0) - 0.5 * O[i]) + c * O2[i];
    }
    double u[3];
    matvec3(Vinv, t, u);
    for (int i = 0; i < 3; ++i) { out[i] = w[i]; out[i + 3] = u[i]; }
}

// ---------------------------------------------------------------- plane fit
// ComputeCovariance (pointcloud.h:126-158): nine running sums in list order,
// each / m, then E[ab] - E[a]E[b].  cov = {xx, xy, xz, yy, yz, zz}.
struct Cumulants {
    double c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    inline void add(double x, double y, double z) {
        c[0] += x; c[1] += y; c[2] += z;
        c[3] += x * x; c[4] += x * y; c[5] += x * z;
        c[6] += y * y; c[7] += y * z; c[8] += z * z;
    }
    inline void finish(int m, double cov[6]) {
        for (int i = 0; i < 9; ++i) c[i] /= (double)m;
        cov[0] = c[3] - c[0] * c[0];
        cov[3] = c[6] - c[1] * c[1];
        cov[5] = c[8] - c[2] * c[2];
        cov[1] = c[4] - c[0] * c[1];
        cov[2] = c[5] - c[0] * c[2];
        cov[4] = c[7] - c[1] * c[2];
    }
};

inline void cross3(const double a[3], const double b[3], double c[3]) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

// ComputeEigenvector0 (pointcloud.h:194-222); A = {a00,a01,a02,a11,a12,a22}
inline void eigenvector0(const double A[6], double eval0, double out[3]) {
    const double row0[3] = {A[0] - eval0, A[1], A[2]};
    const double row1[3] = {A[1], A[3] - eval0, A[4]};
    const double row2[3] = {A[2], A[4], A[5] - eval0};
    double r0xr1[3], r0xr2[3], r1xr2[3];
    cross3(row0, row1, r0xr1);
    cross3(row0, row2, r0xr2);
    cross3(row1, row2, r1xr2);
    const double d0 = dot3(r0xr1, r0xr1), d1 = dot3(r0xr2, r0xr2), d2 = dot3(r1xr2, r1xr2);
    double dmax = d0;
    int imax = 0;
    if (d1 > dmax) { dmax = d1; imax = 1; }
    if (d2 > dmax) { imax = 2; }
    const double *src = imax == 0 ? r0xr1 : (imax == 1 ? r0xr2 : r1xr2);
    const double dd = imax == 0 ? d0 : (imax == 1 ? d1 : d2);
    const double sq = std::sqrt(dd);
    for (int i = 0; i < 3; ++i) out[i] = src[i] / sq;
}

// ComputeEigenvector1 (pointcloud.h:224-288)
inline void eigenvector1(const double A[6], const double evec0[3], double eval1, double out[3]) {
    double U[3], V[3];
    if (std::fabs(evec0[0]) > std::fabs(evec0[1])) {
        const double inv_length = 1 / std::sqrt(evec0[0] * evec0[0] + evec0[2] * evec0[2]);
        U[0] = -evec0[2] * inv_length; U[1] = 0; U[2] = evec0[0] * inv_length;
    } else {
        const double inv_length = 1 / std::sqrt(evec0[1] * evec0[1] + evec0[2] * evec0[2]);
        U[0] = 0; U[1] = evec0[2] * inv_length; U[2] = -evec0[1] * inv_length;
    }
    cross3(evec0, U, V);
    const double AU[3] = {A[0] * U[0] + A[1] * U[1] + A[2] * U[2], A[1] * U[0] + A[3] * U[1] + A[4] * U[2],
                          A[2] * U[0] + A[4] * U[1] + A[5] * U[2]};
    const double AV[3] = {A[0] * V[0] + A[1] * V[1] + A[2] * V[2], A[1] * V[0] + A[3] * V[1] + A[4] * V[2],
                          A[2] * V[0] + A[4] * V[1] + A[5] * V[2]};
    double m00 = U[0] * AU[0] + U[1] * AU[1] + U[2] * AU[2] - eval1;
    double m01 = U[0] * AV[0] + U[1] * AV[1] + U[2] * AV[2];
    double m11 = V[0] * AV[0] + V[1] * AV[1] + V[2] * AV[2] - eval1;
    const double absM00 = std::fabs(m00), absM01 = std::fabs(m01), absM11 = std::fabs(m11);
    if (absM00 >= absM11) {
        const double max_abs_comp = std::max(absM00, absM01);
        if (max_abs_comp > 0) {
            if (absM00 >= absM01) { m01 /= m00; m00 = 1 / std::sqrt(1 + m01 * m01); m01 *= m00; }
            else { m00 /= m01; m01 = 1 / std::sqrt(1 + m00 * m00); m00 *= m01; }
            for (int i = 0; i < 3; ++i) out[i] = m01 * U[i] - m00 * V[i];
        } else {
            for (int i = 0; i < 3; ++i) out[i] = U[i];
        }
    } else {
        const double max_abs_comp = std::max(absM11, absM01);
        if (max_abs_comp > 0) {
            if (absM11 >= absM01) { m01 /= m11; m11 = 1 / std::sqrt(1 + m01 * m01); m01 *= m11; }
            else { m11 /= m01; m01 = 1 / std::sqrt(1 + m11 * m11); m11 *= m01; }
            for (int i = 0; i < 3; ++i) out[i] = m11 * U[i] - m01 * V[i];
        } else {
            for (int i = 0; i < 3; ++i) out[i] = U[i];
        }
    }
}

// FastEigen3x3_EV(...).first (pointcloud.h:378-463): eigenvector of the smallest
// eigenvalue.  cov = {xx, xy, xz, yy, yz, zz}.  Returns the un-normalised vector.
inline void smallest_eigenvector(const double cov[6], double n[3]) {
    double A[6];
    double max_coeff = cov[0];
    for (int i = 1; i < 6; ++i) max_coeff = std::max(max_coeff, cov[i]);  // Matrix3d::maxCoeff over all 9 (symmetric)
    if (max_coeff == 0) { n[0] = n[1] = n[2] = 0; return; }
    for (int i = 0; i < 6; ++i) A[i] = cov[i] / max_coeff;
    const double norm = A[1] * A[1] + A[2] * A[2] + A[4] * A[4];
    if (norm > 0) {
        const double q = (A[0] + A[3] + A[5]) / 3;
        const double b00 = A[0] - q, b11 = A[3] - q, b22 = A[5] - q;
        const double p = std::sqrt((b00 * b00 + b11 * b11 + b22 * b22 + norm * 2) / 6);
        const double c00 = b11 * b22 - A[4] * A[4];
        const double c01 = A[1] * b22 - A[4] * A[2];
        const double c02 = A[1] * A[4] - b11 * A[2];
        const double det = (b00 * c00 - A[1] * c01 + A[2] * c02) / (p * p * p);
        double half_det = det * 0.5;
        half_det = std::min(std::max(half_det, -1.0), 1.0);
        const double angle = std::acos(half_det) / (double)3;
        const double two_thirds_pi = 2.09439510239319549;
        const double beta2 = std::cos(angle) * 2;
        const double beta0 = std::cos(angle + two_thirds_pi) * 2;
        const double beta1 = -(beta0 + beta2);
        const double eval[3] = {q + p * beta0, q + p * beta1, q + p * beta2};
        double evec0[3], evec1[3], evec2[3];
        if (half_det >= 0) {
            eigenvector0(A, eval[2], evec2);
            if (eval[2] < eval[0] && eval[2] < eval[1]) { for (int i = 0; i < 3; ++i) n[i] = evec2[i]; return; }
            eigenvector1(A, evec2, eval[1], evec1);
            if (eval[1] < eval[0] && eval[1] < eval[2]) { for (int i = 0; i < 3; ++i) n[i] = evec1[i]; return; }
            cross3(evec1, evec2, n);
        } else {
            eigenvector0(A, eval[0], evec0);
            if (eval[0] < eval[1] && eval[0] < eval[2]) { for (int i = 0; i < 3; ++i) n[i] = evec0[i]; return; }
            eigenvector1(A, evec0, eval[1], evec1);
            if (eval[1] < eval[0] && eval[1] < eval[2]) { for (int i = 0; i < 3; ++i) n[i] = evec1[i]; return; }
            cross3(evec0, evec1, n);
        }
    } else {
        // A *= max_coeff: compare the original diagonal (pointcloud.h:452-460)
        const double a00 = A[0] * max_coeff, a11 = A[3] * max_coeff, a22 = A[5] * max_coeff;
        if (a00 < a11 && a00 < a22) { n[0] = 1; n[1] = 0; n[2] = 0; }
        else if (a11 < a00 && a11 < a22) { n[0] = 0; n[1] = 1; n[2] = 0; }
        else { n[0] = 0; n[1] = 0; n[2] = 1; }
    }
}

// Vector3d::normalize(): v /= sqrt(squaredNorm) if the norm is > 0
inline void normalize3(double v[3]) {
    const double z = dot3(v, v);
    if (z > 0) { const double nn = std::sqrt(z); v[0] /= nn; v[1] /= nn; v[2] /= nn; }
}


// ---------------------------------------------------------------- GPR hyper-parameter objective
// GPRHyperLoss::Evaluate (include/GPR.hpp:154-174) restated matrix by matrix, INCLUDING its gradient expressions as
// written: grad_kernel (GPR.hpp:218-222) returns {2 * sigma * Kff, Kff * Dist * inv_l3} where Kff is the full kernel
// matrix (sigma^2 and the sigma_noise diagonal included) and `Kff * Dist` is Eigen's MATRIX product of two MatrixXd.
// Returns false when the Cholesky factorisation fails (Evaluate returns false, GPR.hpp:159-160).
inline bool GprHyperLoss(const double *X /*[n][2]*/, const double *y, int n, double sigma_noise, double sigma, double l, double *cost,
                         double grad[2]) {
    std::vector<double> Dist((size_t)n * n, 0.0), Kff((size_t)n * n), L((size_t)n * n, 0.0), Kinv((size_t)n * n), alpha(n);
    for (int ri = 0; ri < n - 1; ++ri)  // self_pdist (GPR.hpp:41-54)
        for (int ci = ri + 1; ci < n; ++ci) {
            const double d0 = X[ri * 2] - X[ci * 2], d1 = X[ri * 2 + 1] - X[ci * 2 + 1];
            Dist[(size_t)ri * n + ci] = d0 * d0 + d1 * d1;
            Dist[(size_t)ci * n + ri] = Dist[(size_t)ri * n + ci];
        }
    for (size_t i = 0; i < (size_t)n * n; ++i) Kff[i] = sigma * sigma * std::exp(-0.5 / (l * l) * Dist[i]);  // computeCovariance (GPR.hpp:206-209)
    for (int i = 0; i < n; ++i) Kff[(size_t)i * n + i] += sigma_noise;
    for (int k = 0; k < n; ++k) {  // Eigen::LLT
        double x = Kff[(size_t)k * n + k];
        for (int j = 0; j < k; ++j) x -= L[(size_t)k * n + j] * L[(size_t)k * n + j];
        if (!(x > 0)) return false;
        x = std::sqrt(x);
        L[(size_t)k * n + k] = x;
        for (int i = k + 1; i < n; ++i) {
            double v = Kff[(size_t)i * n + k];
            for (int j = 0; j < k; ++j) v -= L[(size_t)i * n + j] * L[(size_t)k * n + j];
            L[(size_t)i * n + k] = v / x;
        }
    }
    auto llt_solve = [&](std::vector<double> &b) {
        for (int i = 0; i < n; ++i) { double v = b[i]; for (int j = 0; j < i; ++j) v -= L[(size_t)i * n + j] * b[j]; b[i] = v / L[(size_t)i * n + i]; }
        for (int i = n - 1; i >= 0; --i) { double v = b[i]; for (int j = i + 1; j < n; ++j) v -= L[(size_t)j * n + i] * b[j]; b[i] = v / L[(size_t)i * n + i]; }
    };
    for (int i = 0; i < n; ++i) alpha[i] = y[i];
    llt_solve(alpha);
    for (int c = 0; c < n; ++c) {  // Kinv: identity solved in place (GPR.hpp:197-200)
        std::vector<double> e(n, 0.0);
        e[c] = 1.0;
        llt_solve(e);
        for (int i = 0; i < n; ++i) Kinv[(size_t)i * n + c] = e[i];
    }
    double ya = 0, log_diag = 0;
    for (int i = 0; i < n; ++i) { ya += y[i] * alpha[i]; log_diag += std::log(L[(size_t)i * n + i]); }
    *cost = 0.5 * (ya + 2.0 * log_diag + n * std::log(2 * M_PI));  // GPR.hpp:163, LogDet :88-91
    if (grad) {
        const double inv_l3 = 1.0 / (l * l * l);
        std::vector<double> inner((size_t)n * n), dKs((size_t)n * n), dKl((size_t)n * n, 0.0);
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) inner[(size_t)i * n + j] = alpha[i] * alpha[j] - Kinv[(size_t)i * n + j];
        for (size_t i = 0; i < (size_t)n * n; ++i) dKs[i] = 2 * sigma * Kff[i];
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                double sacc = 0;
                for (int k = 0; k < n; ++k) sacc += Kff[(size_t)i * n + k] * Dist[(size_t)k * n + j];
                dKl[(size_t)i * n + j] = sacc * inv_l3;
            }
        double t0 = 0, t1 = 0;  // (inner * dK).trace()
        for (int i = 0; i < n; ++i)
            for (int k = 0; k < n; ++k) { t0 += inner[(size_t)i * n + k] * dKs[(size_t)k * n + i]; t1 += inner[(size_t)i * n + k] * dKl[(size_t)k * n + i]; }
        grad[0] = -0.5 * t0;
        grad[1] = -0.5 * t1;
    }
    return true;
}

}  // namespace orc
