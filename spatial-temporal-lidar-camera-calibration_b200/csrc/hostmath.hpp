// hostmath.hpp — host-side per-candidate preparation (product code; libm on the host).
//
// Sim3Exp (include/g2o_tools.h:106-140) is evaluated ONCE per candidate on the host and
// shipped to the device as R, t, s and the inverse [R^T | -(R^T t)]:  sin/cos/pow come
// from the host libm (the one the reference itself would call), so the device kernels
// only ever need +,-,*,/ and sqrt in fp64, which are IEEE-exact on the GPU.
// Operation order: left-to-right, no FMA (this TU is built with -ffp-contract=off).
#pragma once
#include <cmath>

#include "common.cuh"

namespace stl {

template <class T> inline void h_matmul3(const T *A, const T *B, T *C) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C[i * 3 + j] = (A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j]) + A[i * 3 + 2] * B[6 + j];
}

// [omega, upsilon] -> R, t = V * upsilon; Taylor branch below 1e-4 rad (g2o_tools.h:119-137)
inline void h_se3_exp(const double *x, double *R, double *t) {
    const double w0 = x[0], w1 = x[1], w2 = x[2];
    const double theta = std::sqrt((w0 * w0 + w1 * w1) + w2 * w2);
    const double O[9] = {0.0, -w2, w1, w2, 0.0, -w0, -w1, w0, 0.0};
    double O2[9], V[9];
    h_matmul3(O, O, O2);
    double a, b, c;
    bool taylor = theta < 1e-4;
    if (taylor) {
        a = 1.0; b = 0.5; c = 1.0 / 6.0;
    } else {
        const double costh = std::cos(theta), sinth = std::sin(theta);
        const double invth2 = std::pow(theta, -2.0), invth3 = std::pow(theta, -3.0);
        a = sinth / theta; b = (1.0 - costh) * invth2; c = (theta - sinth) * invth3;
    }
    for (int i = 0; i < 9; ++i) {
        const double I = (i % 4 == 0) ? 1.0 : 0.0;
        // Taylor branch: R = I + Omega + 0.5*Omega2 (no multiplication of Omega by 1)
        R[i] = taylor ? ((I + O[i]) + b * O2[i]) : ((I + a * O[i]) + b * O2[i]);
        V[i] = (I + b * O[i]) + c * O2[i];
    }
    for (int i = 0; i < 3; ++i) t[i] = (V[i * 3] * x[3] + V[i * 3 + 1] * x[4]) + V[i * 3 + 2] * x[5];
}

inline void make_candidate(const double *x, DevCand *c) {
    h_se3_exp(x, c->R, c->t);
    c->s = x[6];  // the scale is a plain multiplier, not exponentiated (g2o_tools.h:138)
    c->sf = (float)x[6];
    c->pad_ = 0.f;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) c->Ri[i * 3 + j] = c->R[j * 3 + i];
    for (int i = 0; i < 3; ++i) c->ti[i] = -((c->Ri[i * 3] * c->t[0] + c->Ri[i * 3 + 1] * c->t[1]) + c->Ri[i * 3 + 2] * c->t[2]);
}

}  // namespace stl
