// oracle_calib.hpp — CPU ORACLE (test infrastructure, NOT product code).
// Restatement of the two g2o problems that precede the cost evaluation:
//   EdgeHE / EdgeRegulation            include/NLHECalib.hpp:27-116   (error and Jacobian exactly as coded)
//   calibEdge::operator()<T>           src/orb_slam/src/Optimizer.cc:83-196 (autodiff on Dual<7>, the reference uses g2o's Jet)
//   robust kernel + quadratic form     g2o::RobustKernelHuber::robustify, BaseUnaryEdge::constructQuadraticForm
//                                      (third-party, restated from g2o 20230223: H += J^T rho' Omega J, b += -J^T rho' Omega e)
// Eigen::AngleAxisd(Matrix3d) is restated from Eigen 3.3 (Quaternion.h / AngleAxis.h) — parity unpinned like SE3Log.
#pragma once
#include <cmath>
#include <cstring>

#include "../include/stlcalib.h"
#include "oracle_math.hpp"

namespace orc {

inline void RotVecFromMatrix(const double R[9], double rv[3]) {
    double q[4];
    const double t = R[0] + R[4] + R[8];
    if (t > 0) {
        double s = std::sqrt(t + 1.0);
        q[3] = 0.5 * s; s = 0.5 / s;
        q[0] = (R[7] - R[5]) * s; q[1] = (R[2] - R[6]) * s; q[2] = (R[3] - R[1]) * s;
    } else {
        int i = 0;
        if (R[4] > R[0]) i = 1;
        if (R[8] > R[i * 3 + i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        double s = std::sqrt(R[i * 3 + i] - R[j * 3 + j] - R[k * 3 + k] + 1.0);
        q[i] = 0.5 * s; s = 0.5 / s;
        q[3] = (R[k * 3 + j] - R[j * 3 + k]) * s;
        q[j] = (R[j * 3 + i] + R[i * 3 + j]) * s;
        q[k] = (R[k * 3 + i] + R[i * 3 + k]) * s;
    }
    double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
    if (n != 0) {
        const double angle = 2.0 * std::atan2(n, std::fabs(q[3]));
        if (q[3] < 0) n = -n;
        for (int a = 0; a < 3; ++a) rv[a] = angle * (q[a] / n);
    } else {
        rv[0] = rv[1] = rv[2] = 0;
    }
}

// EdgeHE::computeError + linearizeOplus (weight = 1)
inline void HeEdgeEval(const double Ta[12], const double Tb[12], const double x[7], double e[3], double J[21]) {
    double R[9], t[3], s;
    Sim3Exp<double>(x, R, t, s);
    double Ra[9], Rb[9], ta[3], tb[3], ra[3], rb[3];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) { Ra[i * 3 + j] = Ta[i * 4 + j]; Rb[i * 3 + j] = Tb[i * 4 + j]; }
        ta[i] = Ta[i * 4 + 3]; tb[i] = Tb[i * 4 + 3];
    }
    RotVecFromMatrix(Ra, ra);
    RotVecFromMatrix(Rb, rb);
    double Rrb[3], Rtb[3];
    matvec3(R, rb, Rrb);
    matvec3(R, tb, Rtb);
    for (int r = 0; r < 3; ++r) {
        const double A[4] = {Ra[r * 3] - (r == 0), Ra[r * 3 + 1] - (r == 1), Ra[r * 3 + 2] - (r == 2), ta[r]};
        const double errTran = (((A[0] * t[0] + A[1] * t[1]) + A[2] * t[2]) + A[3] * s) - Rtb[r];
        e[r] = (Rrb[r] - ra[r]) + errTran;
    }
    // -1.0 * Eigen::skew(v) with Eigen::skew(v) = [0 v2 -v1; -v2 0 v0; v1 -v0 0] (NLHECalib.hpp:20-24)
    const double Jr[9] = {0, -Rrb[2], Rrb[1], Rrb[2], 0, -Rrb[0], -Rrb[1], Rrb[0], 0};
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) { J[r * 7 + c] = Jr[r * 3 + c]; J[r * 7 + 3 + c] = Ra[r * 3 + c] - (r == c); }
        J[r * 7 + 6] = ta[r];
    }
}

template <class T> inline void Cross3(const T a[3], const T b[3], T o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
template <class T> inline void Rodrigues(const T w[3], const T X[3], T out[3]) {  // the block repeated three times in calibEdge
    const T theta = sqrt((w[0] * w[0] + w[1] * w[1]) + w[2] * w[2]);
    if (value_of(theta) > 0) {
        const T axis[3] = {w[0] / theta, w[1] / theta, w[2] / theta};
        const T cth = cos(theta), sth = sin(theta);
        T aX[3];
        Cross3(axis, X, aX);
        const T d = (axis[0] * X[0] + axis[1] * X[1]) + axis[2] * X[2];
        const T omc = T(1.0) - cth;
        for (int i = 0; i < 3; ++i) out[i] = (X[i] * cth + aX[i] * sth) + (axis[i] * d) * omc;
    } else {
        T wX[3];
        Cross3(w, X, wX);
        for (int i = 0; i < 3; ++i) out[i] = X[i] + wX[i];
    }
}

// calibEdge::operator() (Optimizer.cc:83-196)
template <class T>
inline void CalibEdgeEval(const T calib[7], const double Xw[3], const double Tlw_quat[6], double fx, double fy, double cx, double cy,
                          const double obs[2], T err[2]) {
    const T scale = calib[6];
    const T Xc0[3] = {scale * T(Xw[0]), scale * T(Xw[1]), scale * T(Xw[2])};
    const T w1[3] = {-calib[0], -calib[1], -calib[2]};
    const T t[3] = {-calib[3], -calib[4], -calib[5]};
    T p[3];
    Rodrigues(w1, t, p);
    T Xl0[3], Xli[3], Xci[3];
    Rodrigues(w1, Xc0, Xl0);
    for (int i = 0; i < 3; ++i) Xl0[i] = Xl0[i] + p[i];
    const T w2[3] = {T(Tlw_quat[0]), T(Tlw_quat[1]), T(Tlw_quat[2])};
    Rodrigues(w2, Xl0, Xli);
    for (int i = 0; i < 3; ++i) Xli[i] = Xli[i] + T(Tlw_quat[3 + i]);
    const T w3[3] = {calib[0], calib[1], calib[2]};
    Rodrigues(w3, Xli, Xci);
    for (int i = 0; i < 3; ++i) Xci[i] = Xci[i] + calib[3 + i];
    const T pre0 = T(fx) * Xci[0] / Xci[2] + T(cx), pre1 = T(fy) * Xci[1] / Xci[2] + T(cy);
    err[0] = T(obs[0]) - pre0;
    err[1] = T(obs[1]) - pre1;
}

struct G2oAcc {
    stl_lin_sums_t *o;
    void add(const double *e, const double *J /*[nr][7]*/, int nr, double info, double delta, double *chi2) {
        double e2 = 0;
        for (int r = 0; r < nr; ++r) e2 += e[r] * e[r];
        e2 *= info;
        if (chi2) *chi2 = e2;
        double rho0 = e2, rho1 = 1.0;
        if (delta > 0 && e2 > delta * delta) { const double sq = std::sqrt(e2); rho0 = 2 * sq * delta - delta * delta; rho1 = delta / sq; }
        o->cost += rho0;
        for (int r = 0; r < nr; ++r)
            for (int a = 0; a < 7; ++a) {
                o->g[a] += J[r * 7 + a] * (rho1 * info) * e[r];
                for (int c = 0; c < 7; ++c) o->H[a * 7 + c] += J[r * 7 + a] * (rho1 * info) * J[r * 7 + c];
            }
        o->n_blocks_2d += 1;
        o->n_residuals += nr;
    }
};

inline void HeLinearize(const stl_he_edges_t &ed, const double x[7], stl_lin_sums_t *out, double *chi2) {
    std::memset(out, 0, sizeof(*out));
    G2oAcc acc{out};
    for (int i = 0; i < ed.n; ++i) {
        double e[3], J[21];
        HeEdgeEval(ed.Ta + (size_t)i * 12, ed.Tb + (size_t)i * 12, x, e, J);
        acc.add(e, J, 3, ed.info ? ed.info[i] : 1.0, ed.huber_delta, chi2 ? chi2 + i : nullptr);
    }
    if (ed.regulation > 0) {  // EdgeRegulation: error = params[3..5], no robust kernel
        const double e[3] = {x[3], x[4], x[5]};
        double J[21] = {0};
        J[0 * 7 + 3] = J[1 * 7 + 4] = J[2 * 7 + 5] = 1.0;
        acc.add(e, J, 3, ed.regulation, 0.0, nullptr);
        out->n_blocks_2d -= 1;  // counted apart from the motion edges
    }
}

inline void CalibLinearize(const stl_calib_edges_t &ed, const double x[7], stl_lin_sums_t *out, double *chi2) {
    std::memset(out, 0, sizeof(*out));
    G2oAcc acc{out};
    typedef Dual<7> D;
    D xd[7];
    for (int i = 0; i < 7; ++i) xd[i] = D::var(x[i], i);
    for (int f = 0; f < ed.n_kf; ++f)
        for (int64_t i = ed.edge_offset[f]; i < ed.edge_offset[f + 1]; ++i) {
            D err[2];
            CalibEdgeEval<D>(xd, ed.Xw + i * 3, ed.Tlw_quat + (size_t)f * 6, ed.intrinsics[f * 4], ed.intrinsics[f * 4 + 1], ed.intrinsics[f * 4 + 2],
                             ed.intrinsics[f * 4 + 3], ed.obs + i * 2, err);
            const double e[2] = {err[0].a, err[1].a};
            double J[14];
            for (int a = 0; a < 7; ++a) { J[a] = err[0].v[a]; J[7 + a] = err[1].v[a]; }
            const double info = (double)ed.inv_sigma2[i];
            if (ed.level && ed.level[i]) {
                if (chi2) chi2[i] = info * (e[0] * e[0] + e[1] * e[1]);
                continue;
            }
            acc.add(e, J, 2, info, ed.huber_delta, chi2 ? chi2 + i : nullptr);
        }
}

}  // namespace orc
