// calibmath.hpp — host-side preparation for the two optimisation problems that precede the cost evaluation on the
// same 7-parameter block: the hand-eye initialisation (include/NLHECalib.hpp) and the calibration
// bundle adjustment of the ORB-SLAM2 fork (src/orb_slam/src/Optimizer.cc:65-205,1399-1744).
// Everything transcendental is evaluated here with the host libm, once per candidate / per keyframe; the device
// kernels (calib.cu) only multiply, add and divide.
#pragma once
#include <cmath>

#include "dual.cuh"
#include "hostmath.hpp"

namespace stl {

// Eigen::AngleAxisd(R).angle() * axis() (NLHECalib.hpp:40,60; Optimizer.cc:1413-1415): rotation matrix -> quaternion
// (Eigen 3.3 Quaternion.h, as in se3.cuh) -> angle/axis (AngleAxis.h: angle = 2 atan2(|vec|, |w|), axis = vec / (+-|vec|))
inline void h_rotvec_from_R(const double *R, double rv[3]) {
    double q[4];  // x y z w
    const double tr = (R[0] + R[4]) + R[8];
    if (tr > 0.0) {
        double s = std::sqrt(tr + 1.0);
        q[3] = 0.5 * s;
        s = 0.5 / s;
        q[0] = (R[7] - R[5]) * s; q[1] = (R[2] - R[6]) * s; q[2] = (R[3] - R[1]) * s;
    } else {
        int i = 0;
        if (R[4] > R[0]) i = 1;
        if (R[8] > R[i * 4]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        double s = std::sqrt(R[i * 4] - R[j * 4] - R[k * 4] + 1.0);
        q[i] = 0.5 * s;
        s = 0.5 / s;
        q[3] = (R[k * 3 + j] - R[j * 3 + k]) * s;
        q[j] = (R[j * 3 + i] + R[i * 3 + j]) * s;
        q[k] = (R[k * 3 + i] + R[i * 3 + k]) * s;
    }
    double n = std::sqrt((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]);
    if (n != 0.0) {
        const double angle = 2.0 * std::atan2(n, std::fabs(q[3]));
        if (q[3] < 0.0) n = -n;
        for (int a = 0; a < 3; ++a) rv[a] = angle * (q[a] / n);
    } else {
        rv[0] = rv[1] = rv[2] = 0.0;  // angle 0, axis (1, 0, 0)
    }
}

// One EdgeHE (NLHECalib.hpp:27-86): candidate-independent constants
struct HeEdge {
    double ra[3], rb[3];  // rotation vectors of Ta, Tb
    double Ra[9], ta[3], tb[3];
    double info;          // information = info * I
};

inline void make_he_edge(const double *Ta, const double *Tb, double info, HeEdge *e) {
    double Rb[9];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) { e->Ra[i * 3 + j] = Ta[i * 4 + j]; Rb[i * 3 + j] = Tb[i * 4 + j]; }
        e->ta[i] = Ta[i * 4 + 3];
        e->tb[i] = Tb[i * 4 + 3];
    }
    h_rotvec_from_R(e->Ra, e->ra);
    h_rotvec_from_R(Rb, e->rb);
    e->info = info;
}

// calibEdge::operator() (Optimizer.cc:83-196) — the parts that depend on the candidate only, as duals
struct CalibCand {
    D7 scale;
    D7 v[3], cth, sth, omc, p[3];       // Tlc: axis of -omega, cos / sin / 1 - cos of its norm, translation p
    D7 a3[3], cth3, sth3, omc3, t3[3];  // calib: axis of omega, ..., translation
    int small1, small3;                 // theta == 0 branches (first-order formulas)
    D7 w1[3], w3[3];                    // the rotation vectors themselves (used by the theta == 0 branches)
};

inline D7 d7_norm3(const D7 *a) { return d7_sqrt((a[0] * a[0] + a[1] * a[1]) + a[2] * a[2]); }
inline void d7_cross(const D7 *a, const D7 *b, D7 *o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

inline void make_calib_candidate(const double *x, CalibCand *c) {
    D7 calib[7];
    for (int i = 0; i < 7; ++i) calib[i] = d7_var(x[i], i);
    c->scale = calib[6];
    D7 t[3];
    for (int i = 0; i < 3; ++i) { c->w1[i] = -calib[i]; t[i] = -calib[3 + i]; c->w3[i] = calib[i]; c->t3[i] = calib[3 + i]; }
    // norm() of a vector whose entries are exactly zero has no derivative; the reference's `theta > T(0)` test takes
    // the first-order branch there (Optimizer.cc:100-113)
    const bool zero = x[0] == 0.0 && x[1] == 0.0 && x[2] == 0.0;
    c->small1 = c->small3 = zero ? 1 : 0;
    if (!zero) {
        const D7 theta = d7_norm3(c->w1);
        for (int i = 0; i < 3; ++i) c->v[i] = c->w1[i] / theta;
        c->cth = d7_cos(theta); c->sth = d7_sin(theta);
        D7 vXt[3];
        d7_cross(c->v, t, vXt);
        const D7 vDott = (c->v[0] * t[0] + c->v[1] * t[1]) + c->v[2] * t[2];
        c->omc = d7_const(1.0) - c->cth;
        for (int i = 0; i < 3; ++i) c->p[i] = (t[i] * c->cth + vXt[i] * c->sth) + (c->v[i] * vDott) * c->omc;
        const D7 theta3 = d7_norm3(c->w3);
        for (int i = 0; i < 3; ++i) c->a3[i] = c->w3[i] / theta3;
        c->cth3 = d7_cos(theta3); c->sth3 = d7_sin(theta3);
        c->omc3 = d7_const(1.0) - c->cth3;
    } else {
        D7 wXt[3];
        d7_cross(c->w1, t, wXt);
        for (int i = 0; i < 3; ++i) { c->p[i] = t[i] + wXt[i]; c->v[i] = c->a3[i] = d7_const(0.0); }
        c->cth = c->cth3 = d7_const(1.0); c->sth = c->sth3 = c->omc = c->omc3 = d7_const(0.0);
    }
}

// per keyframe: the LiDAR pose of calibEdge::Tlw_quat (Optimizer.cc:139-157), constants
struct CalibKf {
    double a2[3], cth2, sth2, omc2, w2[3], t2[3];
    double fx, fy, cx, cy;
    int small2, pad_;
};

inline void make_calib_kf(const double *Tlw_quat, const float *intr, CalibKf *k) {
    const double th = std::sqrt((Tlw_quat[0] * Tlw_quat[0] + Tlw_quat[1] * Tlw_quat[1]) + Tlw_quat[2] * Tlw_quat[2]);
    k->small2 = th > 0.0 ? 0 : 1;
    const double inv = th > 0.0 ? 1.0 / th : 0.0;  // Jet division multiplies by the reciprocal of the denominator's value
    for (int i = 0; i < 3; ++i) { k->w2[i] = Tlw_quat[i]; k->t2[i] = Tlw_quat[3 + i]; k->a2[i] = Tlw_quat[i] * inv; }
    k->cth2 = std::cos(th); k->sth2 = std::sin(th); k->omc2 = 1.0 - k->cth2;
    k->fx = intr[0]; k->fy = intr[1]; k->cx = intr[2]; k->cy = intr[3];
    k->pad_ = 0;
}

}  // namespace stl
