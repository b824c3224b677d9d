// oracle_refmath.cpp — CPU ORACLE support (test infrastructure, NOT product code).
// Compiles slices of the REFERENCE's own sources, verbatim, against ref_shim_eigen.hpp.  The slices are cut out of
// /root/reference at build time (oracle/Makefile: sed line ranges -> oracle/_ref/*.inc, git-ignored); no reference source
// is copied into the repository.
#include <algorithm>
#include <cstdint>
#include <tuple>
#include <vector>

#include "ref_shim_eigen.hpp"
#include "oracle_math.hpp"  // orc::Dual<7> with cos / sin / pow / sqrt, found by ADL inside the templated slice

typedef std::vector<Eigen::Vector3d> VecVector3d;
typedef std::uint32_t IndexType;

#include "_ref/ref_pointcloud_covariance.inc"   // include/pointcloud.h:126-158  ComputeCovariance
#include "_ref/ref_pointcloud_eigvec.inc"       // include/pointcloud.h:194-288  ComputeEigenvector0 / 1
#include "_ref/ref_pointcloud_fasteigen.inc"    // include/pointcloud.h:378-463  FastEigen3x3_EV
#include "_ref/ref_g2o_tools_skew.inc"          // include/g2o_tools.h:58-69     skew
#include "_ref/ref_g2o_tools_sim3exp.inc"       // include/g2o_tools.h:105-140   Sim3Exp<T>
#include "_ref/ref_g2o_tools_se3exp.inc"        // include/g2o_tools.h:149-183   SE3Exp<T>

extern "C" {
// covariance of pts[idx[0..n)] -> cov[9] row-major
void refm_covariance(const double *pts, int npts, const uint32_t *idx, int n, double cov[9]) {
    VecVector3d P((size_t)npts);
    for (int i = 0; i < npts; ++i) P[i] = Eigen::Vector3d(pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2]);
    std::vector<IndexType> I(idx, idx + n);
    const Eigen::Matrix3d C = ComputeCovariance<IndexType>(P, I);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) cov[i * 3 + j] = C(i, j);
}
void refm_fast_eigen(const double cov[9], double evec[3], double eval[3]) {
    Eigen::Matrix3d C;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) C(i, j) = cov[i * 3 + j];
    Eigen::Vector3d v, e;
    std::tie(v, e) = FastEigen3x3_EV(C);
    for (int i = 0; i < 3; ++i) { evec[i] = v(i); eval[i] = e(i); }
}
void refm_sim3exp(const double x[7], double R[9], double t[3], double *s) {
    auto [Rm, tv, sc] = Sim3Exp<double>(x);
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) R[i * 3 + j] = Rm(i, j); t[i] = tv(i); }
    *s = sc;
}
// the same on duals: value and the 7 partials of every entry of R (9), t (3), s — out[13][8]
void refm_sim3exp_dual(const double x[7], double out[13 * 8]) {
    typedef orc::Dual<7> D;
    D xd[7];
    for (int i = 0; i < 7; ++i) xd[i] = D::var(x[i], i);
    auto [Rm, tv, sc] = Sim3Exp<D>(xd);
    int k = 0;
    auto put = [&](const D &d) { out[k * 8] = d.a; for (int a = 0; a < 7; ++a) out[k * 8 + 1 + a] = d.v[a]; ++k; };
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) put(Rm(i, j));
    for (int i = 0; i < 3; ++i) put(tv(i));
    put(sc);
}
void refm_se3exp(const double x[6], double R[9], double t[3]) {
    auto [Rm, tv] = SE3Exp<double>(x);
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) R[i * 3 + j] = Rm(i, j); t[i] = tv(i); }
}
}
