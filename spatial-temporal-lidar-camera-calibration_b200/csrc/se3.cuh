// se3.cuh — device fp64 SE(3) log for the hand-eye term (iba_global.cpp:264-276).
//
// SE3Log (g2o_tools.h:78-82) = g2o::SE3Quat(R, t).log(): rotation matrix ->
// Eigen quaternion (sign-normalised, unit-normalised) -> rotation matrix -> log.
// g2o/Eigen are third-party and absent from the reference tree; this follows
// their published formulas (g2o release 20230223 se3quat.h, Eigen 3.3
// Quaternion.h).  The quaternion round trip matters: Tc is a float32 product and
// only orthonormal to ~1e-7.  acos/tan/sqrt here are CUDA's (not bit-identical
// to glibc); the term is compared at 1e-6 relative.
#pragma once
#include "common.cuh"

namespace stl {

__device__ inline void mat3mul(const double *A, const double *B, double *C) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) C[i * 3 + j] = dot3e(A[i * 3], A[i * 3 + 1], A[i * 3 + 2], B[j], B[3 + j], B[6 + j]);
}

// (A*B) for rigid [R|t]: R = A.R*B.R, t = A.R*B.t + A.t
__device__ inline void rt_compose(const double *AR, const double *At, const double *BR, const double *Bt, double *CR, double *Ct) {
    mat3mul(AR, BR, CR);
#pragma unroll
    for (int i = 0; i < 3; ++i) Ct[i] = dadd(dot3e(AR[i * 3], AR[i * 3 + 1], AR[i * 3 + 2], Bt[0], Bt[1], Bt[2]), At[i]);
}

__device__ inline void se3_log(const double *Rin, const double *t, double *out) {
    double q[4];  // x y z w
    const double tr = (Rin[0] + Rin[4]) + Rin[8];
    if (tr > 0.0) {
        double s = sqrt(tr + 1.0);
        q[3] = 0.5 * s;
        s = 0.5 / s;
        q[0] = (Rin[7] - Rin[5]) * s;
        q[1] = (Rin[2] - Rin[6]) * s;
        q[2] = (Rin[3] - Rin[1]) * s;
    } else {
        // i = index of the largest diagonal entry, (i, j, k) cyclic; written out per case so that
        // nothing is indexed dynamically (keeps q[] in registers)
        int i = 0;
        if (Rin[4] > Rin[0]) i = 1;
        if (Rin[8] > (i == 0 ? Rin[0] : Rin[4])) i = 2;
        if (i == 0) {  // j = 1, k = 2
            double s = sqrt(Rin[0] - Rin[4] - Rin[8] + 1.0);
            q[0] = 0.5 * s;
            s = 0.5 / s;
            q[3] = (Rin[7] - Rin[5]) * s;
            q[1] = (Rin[3] + Rin[1]) * s;
            q[2] = (Rin[6] + Rin[2]) * s;
        } else if (i == 1) {  // j = 2, k = 0
            double s = sqrt(Rin[4] - Rin[8] - Rin[0] + 1.0);
            q[1] = 0.5 * s;
            s = 0.5 / s;
            q[3] = (Rin[2] - Rin[6]) * s;
            q[2] = (Rin[7] + Rin[5]) * s;
            q[0] = (Rin[1] + Rin[3]) * s;
        } else {  // j = 0, k = 1
            double s = sqrt(Rin[8] - Rin[0] - Rin[4] + 1.0);
            q[2] = 0.5 * s;
            s = 0.5 / s;
            q[3] = (Rin[3] - Rin[1]) * s;
            q[0] = (Rin[2] + Rin[6]) * s;
            q[1] = (Rin[5] + Rin[7]) * s;
        }
    }
    if (q[3] < 0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
    const double n = sqrt(((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3]);
    q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
    const double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
    const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
    const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
    const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
    double R[9];
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
    const double d = 0.5 * (((R[0] + R[4]) + R[8]) - 1.0);
    const double dR[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]};
    double w[3], c;
    if (fabs(d) > 0.99999) {
        w[0] = 0.5 * dR[0]; w[1] = 0.5 * dR[1]; w[2] = 0.5 * dR[2];
        c = 1.0 / 12.0;
    } else {
        const double theta = acos(d);
        const double f = theta / (2 * sqrt(1 - d * d));
        w[0] = f * dR[0]; w[1] = f * dR[1]; w[2] = f * dR[2];
        c = (1 - theta / (2 * tan(theta / 2))) / (theta * theta);
    }
    const double O[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    double O2[9];
    mat3mul(O, O, O2);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        double acc[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) acc[j] = (((i == j) ? 1.0 : 0.0) - 0.5 * O[i * 3 + j]) + c * O2[i * 3 + j];
        out[3 + i] = dot3e(acc[0], acc[1], acc[2], t[0], t[1], t[2]);
        out[i] = w[i];
    }
}

}  // namespace stl
