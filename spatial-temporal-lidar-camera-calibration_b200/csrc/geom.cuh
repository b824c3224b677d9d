// geom.cuh — device fp64 plane-fit primitives (compiled with -fmad=false so that the
// plain operators below are not contracted; the reference is built without FMA,
// CMakeLists.txt:2,7).
//
//   smallest_eigvec  <- FastEigen3x3_EV(...).first   include/pointcloud.h:378-463
//                       ComputeEigenvector0/1        include/pointcloud.h:194-288
//   (closed-form symmetric 3x3 eigen-solver after Eberly, "A Robust Eigensolver for
//    3x3 Symmetric Matrices"; acos/cos are CUDA's, not bit-identical to glibc: normals
//    agree to ~1e-15, gates are tolerance-aware in the tests.)
#pragma once
#include "crmath.cuh"
#include "common.cuh"

namespace stl {

struct V3 { double x, y, z; };
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ double dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }

// A = {a00, a01, a02, a11, a12, a22}
__device__ inline V3 eigvec_first(const double *A, double ev) {
    const V3 r0 = {A[0] - ev, A[1], A[2]}, r1 = {A[1], A[3] - ev, A[4]}, r2 = {A[2], A[4], A[5] - ev};
    const V3 c01 = cross(r0, r1), c02 = cross(r0, r2), c12 = cross(r1, r2);
    const double d0 = dot(c01, c01), d1 = dot(c02, c02), d2 = dot(c12, c12);
    double dmax = d0;
    int imax = 0;
    if (d1 > dmax) { dmax = d1; imax = 1; }
    if (d2 > dmax) { imax = 2; }
    const V3 c = imax == 0 ? c01 : (imax == 1 ? c02 : c12);
    const double s = sqrt(imax == 0 ? d0 : (imax == 1 ? d1 : d2));
    return {c.x / s, c.y / s, c.z / s};
}

__device__ inline V3 eigvec_second(const double *A, V3 e0, double ev1) {
    V3 U;
    if (fabs(e0.x) > fabs(e0.y)) {
        const double il = 1 / sqrt(e0.x * e0.x + e0.z * e0.z);
        U = {-e0.z * il, 0.0, e0.x * il};
    } else {
        const double il = 1 / sqrt(e0.y * e0.y + e0.z * e0.z);
        U = {0.0, e0.z * il, -e0.y * il};
    }
    const V3 V = cross(e0, U);
    const V3 AU = {A[0] * U.x + A[1] * U.y + A[2] * U.z, A[1] * U.x + A[3] * U.y + A[4] * U.z, A[2] * U.x + A[4] * U.y + A[5] * U.z};
    const V3 AV = {A[0] * V.x + A[1] * V.y + A[2] * V.z, A[1] * V.x + A[3] * V.y + A[4] * V.z, A[2] * V.x + A[4] * V.y + A[5] * V.z};
    double m00 = U.x * AU.x + U.y * AU.y + U.z * AU.z - ev1;
    double m01 = U.x * AV.x + U.y * AV.y + U.z * AV.z;
    double m11 = V.x * AV.x + V.y * AV.y + V.z * AV.z - ev1;
    const double a00 = fabs(m00), a01 = fabs(m01), a11 = fabs(m11);
    if (a00 >= a11) {
        if (fmax(a00, a01) > 0) {
            if (a00 >= a01) { m01 /= m00; m00 = 1 / sqrt(1 + m01 * m01); m01 *= m00; }
            else { m00 /= m01; m01 = 1 / sqrt(1 + m00 * m00); m00 *= m01; }
            return {m01 * U.x - m00 * V.x, m01 * U.y - m00 * V.y, m01 * U.z - m00 * V.z};
        }
        return U;
    }
    if (fmax(a11, a01) > 0) {
        if (a11 >= a01) { m01 /= m11; m11 = 1 / sqrt(1 + m01 * m01); m01 *= m11; }
        else { m11 /= m01; m01 = 1 / sqrt(1 + m11 * m11); m11 *= m01; }
        return {m11 * U.x - m01 * V.x, m11 * U.y - m01 * V.y, m11 * U.z - m01 * V.z};
    }
    return U;
}

// cov = {xx, xy, xz, yy, yz, zz}; returns the (un-normalised) eigenvector of the smallest eigenvalue
__device__ inline V3 smallest_eigvec(const double *cov) {
    double mx = cov[0];
#pragma unroll
    for (int i = 1; i < 6; ++i) mx = fmax(mx, cov[i]);
    if (mx == 0) return {0.0, 0.0, 0.0};
    double A[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) A[i] = cov[i] / mx;
    const double norm = A[1] * A[1] + A[2] * A[2] + A[4] * A[4];
    if (norm > 0) {
        const double q = (A[0] + A[3] + A[5]) / 3;
        const double b00 = A[0] - q, b11 = A[3] - q, b22 = A[5] - q;
        const double p = sqrt((b00 * b00 + b11 * b11 + b22 * b22 + norm * 2) / 6);
        const double c00 = b11 * b22 - A[4] * A[4];
        const double c01 = A[1] * b22 - A[4] * A[2];
        const double c02 = A[1] * A[4] - b11 * A[2];
        const double det = (b00 * c00 - A[1] * c01 + A[2] * c02) / (p * p * p);
        const double hd = fmin(fmax(det * 0.5, -1.0), 1.0);
        // std::acos / std::cos of the reference, correctly rounded (crmath.cuh): plane normals bit-identical to a glibc evaluation
        const double ang = acos_cr(hd) / 3.0;
        const double beta2 = cos_cr(ang) * 2, beta0 = cos_cr(ang + 2.09439510239319549) * 2, beta1 = -(beta0 + beta2);
        const double e0 = q + p * beta0, e1 = q + p * beta1, e2 = q + p * beta2;
        if (hd >= 0) {
            const V3 v2 = eigvec_first(A, e2);
            if (e2 < e0 && e2 < e1) return v2;
            const V3 v1 = eigvec_second(A, v2, e1);
            if (e1 < e0 && e1 < e2) return v1;
            return cross(v1, v2);
        }
        const V3 v0 = eigvec_first(A, e0);
        if (e0 < e1 && e0 < e2) return v0;
        const V3 v1 = eigvec_second(A, v0, e1);
        if (e1 < e0 && e1 < e2) return v1;
        return cross(v0, v1);
    }
    const double a00 = A[0] * mx, a11 = A[3] * mx, a22 = A[5] * mx;
    if (a00 < a11 && a00 < a22) return {1.0, 0.0, 0.0};
    if (a11 < a00 && a11 < a22) return {0.0, 1.0, 0.0};
    return {0.0, 0.0, 1.0};
}

__device__ __forceinline__ V3 normalized(V3 v) {
    const double z = dot(v, v);
    if (z > 0) { const double n = sqrt(z); return {v.x / n, v.y / n, v.z / n}; }
    return v;
}

}  // namespace stl
