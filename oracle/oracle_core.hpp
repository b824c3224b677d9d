// oracle_core.hpp — CPU ORACLE (test infrastructure, NOT product code).
//
// Restatement of the reference's cost evaluation over a KeyFramePack:
//   BAError                      src/examples/iba_global.cpp:169-344
//   FindProjectCorrespondences   src/examples/iba_global.cpp:55-96
//   ComputeAlignmentDist         src/examples/iba_global.cpp:111-156
//   BuildProblem (LM path)       src/examples/iba_local.cpp:145-323
//   IBA_PlaneFactor              include/IBACalib2.hpp:152-184
//   Point2Point/Point2Plane      include/IBACalib2.hpp:570-584,611-625
//   Huber loss + normal equations: Ceres semantics (third-party, unpinned;
//   SURVEY.md Appendix A16), restated.
// Templated on the KD-tree backend: the in-repo port (oracle_kdtree.hpp) or
// the reference's real vendored nanoflann (oracle_ref.cpp -> oracle/_ref/).
//
// PARITY PINNING: the reference ships no tests, golden vectors or fixtures for
// this path and cannot be compiled here (Eigen/g2o/Ceres/OpenCV/Nomad absent,
// SURVEY.md F6/F7).  What IS pinned: every KNN result against the reference's
// own nanoflann compiled from /root/reference/include (oracle/_ref), and three
// analytic known-answer tests (tests/test_oracle_kat.py).  The Eigen operation
// order and g2o's SE3Quat::log are restated from their published sources and
// are "parity unpinned" (see DESIGN.md §Oracle).
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include "../include/stlcalib.h"
#include "oracle_math.hpp"
#include "oracle_calib.hpp"

namespace orc {

struct Corr { uint32_t kp, pt; };

struct AlignResult {   // one ComputeAlignmentDist call
    uint32_t kp;       // keypoint index of the correspondence
    uint32_t nn;       // 1-NN of the map point in the scan
    int32_t m;         // neighbours kept after the radius truncation
    int32_t is_plane;
    double dist;
    uint32_t knn[32];  // first m neighbour indices in distance order
    double normal[3];
};

struct TieStats {  // queries whose reference result depends on KD-tree visit order
    int64_t nn2d = 0, nn3d = 0, knn3d = 0;
};

struct FrameDebug {
    std::vector<Corr> corr;
    std::vector<AlignResult> align;
};

// Residual block frozen by BuildProblem (iba_local.cpp:263-309)
struct Block {
    int type;  // 0 = IBA_PlaneFactor, 1 = Point2Point_Factor, 2 = Point2Plane_Factor, 3 = IBA_GPRFactor
    int kf;
    uint32_t kp;
    // plane factor
    double fx, fy, cx, cy, u0, v0, p0[3], n0[3];
    int ncov;
    double R[STL_MAX_COVIS][9], t[STL_MAX_COVIS][3], u1[STL_MAX_COVIS], v1[STL_MAX_COVIS];
    // 3-D factors
    double map_pt[3], query_pt[3], normal[3];
    // GPR factor: neighbour points (LiDAR frame, distance order, the scan point itself first) and hyper-parameters
    int gm;
    double gpts[32][3];
    double sigma, l, sigma_noise;
};

template <class Tree2, class Tree3>
class Oracle {
  public:
    Oracle(const stl_pack_t *pack, const stl_params_t *params, int leaf2d, int leaf3d, int nthreads)
        : pk_(*pack), pr_(*params), leaf2d_(leaf2d), nthreads_(nthreads) {
        const int F = pk_.n_kf;
        scans_.resize(F);
        trees_.resize(F);
        // Scans widened float32 -> fp64 (io_tools.h:170-187); one KDTree3D per scan (iba_global.cpp:362-367)
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads_)
        for (int f = 0; f < F; ++f) {
            const int64_t n = pk_.scan_offset[f + 1] - pk_.scan_offset[f];
            const float *src = pk_.scan_xyz + pk_.scan_offset[f] * 3;
            scans_[f].resize((size_t)n * 3);
            for (int64_t i = 0; i < n * 3; ++i) scans_[f][i] = (double)src[i];
            trees_[f].reset(new Tree3(scans_[f].data(), (size_t)n, leaf3d));
        }
    }

    // ------------------------------------------------------------------ KNN with tie report
    // Returns the k nearest of tree in (d2, index) order.  Asks the tree for k+1
    // so that an exact tie at any rank (incl. the k-boundary) is detected; without
    // ties the first k equal the reference's own k-search.  strict = the
    // reference's exact call (k results, visit-order ties), used for timing.
    template <class Tree>
    static size_t knn(const Tree &tree, const double *q, size_t k, uint32_t *idx, double *d2, bool strict, bool *tie) {
        if (strict) return tree.knn(q, k, idx, d2);
        uint32_t ti[34]; double td[34];
        const size_t got = tree.knn(q, k + 1, ti, td);
        bool t = false;
        // insertion sort by (d2, idx); entries arrive sorted by d2 already
        for (size_t i = 1; i < got; ++i) {
            if (td[i] == td[i - 1]) t = true;
            size_t j = i;
            while (j > 0 && td[j - 1] == td[j] && ti[j - 1] > ti[j]) { std::swap(ti[j - 1], ti[j]); --j; }
        }
        if (tie) *tie = t;
        const size_t out = got < k ? got : k;
        for (size_t i = 0; i < out; ++i) { idx[i] = ti[i]; d2[i] = td[i]; }
        return out;
    }

    // ------------------------------------------------------------------ FindProjectCorrespondences
    // iba_global.cpp:55-96.  PC = scan in the camera frame (fp64, [n][3]).
    void find_corr(int f, const std::vector<double> &PC, std::vector<Corr> &corrset, bool strict, TieStats *ties) const {
        const double fx = pk_.intrinsics[f * 4], cx = pk_.intrinsics[f * 4 + 2], cy = pk_.intrinsics[f * 4 + 3];
        const double W = pk_.image_wh[f * 2], H = pk_.image_wh[f * 2 + 1];
        const size_t n = PC.size() / 3;
        std::vector<double> proj;
        std::vector<uint32_t> pidx;
        for (size_t i = 0; i < n; ++i) {
            const double x = PC[i * 3], y = PC[i * 3 + 1], z = PC[i * 3 + 2];
            if (z > 0) {
                const double u = (fx * x + cx * z) / z;
                const double v = (fx * y + cy * z) / z;  // fx, not fy (iba_global.cpp:73)
                bool in;
                if (pr_.variant == 1) {  // iba_global_stable.cpp:92-94: the ROUNDED pixel must be inside the image
                    const double ru = std::round(u), rv = std::round(v);
                    in = 0 <= ru && ru < W && 0 <= rv && rv < H;
                } else {
                    in = 0 <= u && u < W && 0 <= v && v < H;
                }
                if (in) { proj.push_back(u); proj.push_back(v); pidx.push_back((uint32_t)i); }
            }
        }
        if (pidx.empty()) return;
        Tree2 tree(proj.data(), pidx.size(), leaf2d_);
        const double maxd2 = pr_.max_pixel_dist * pr_.max_pixel_dist;
        const int64_t k0 = pk_.kp_offset[f], k1 = pk_.kp_offset[f + 1];
        for (int64_t k = k0; k < k1; ++k) {
            double q[2] = {(double)pk_.kp_xy[k * 2], (double)pk_.kp_xy[k * 2 + 1]};
            if (pr_.variant == 1 && !stable_keypoint(f, k, q)) continue;  // only keypoints that observe a map point are queried
            uint32_t idx[2]; double d2[2]; bool tie = false;
            const size_t got = knn(tree, q, 1, idx, d2, strict, &tie);
            if (got > 0 && d2[0] <= maxd2) {
                corrset.push_back({(uint32_t)(k - k0), pidx[idx[0]]});
                if (tie && ties) ties->nn2d++;
            }
        }
    }

    // iba_global_stable.cpp:67-80: the query pixel of a keypoint is the re-projection of its map point
    // with the SLAM pose (float32 widened; fx, fy, cx, cy as stored), in keypoint-index order here
    // (the reference walks an unordered_map; only the summation order depends on it).  The 3x3 * 3x1
    // product is summed (R0 X + R1 Y) + R2 Z, then + t (Eigen's order for this expression is unverified).
    bool stable_keypoint(int f, int64_t k, double q[2]) const {
        const float *mp = pk_.kp_mappoint + (size_t)k * 3;
        if (std::isnan(mp[0])) return false;
        const float *Tcw = pk_.Tcw + (size_t)f * 12;
        const double X = mp[0], Y = mp[1], Z = mp[2];
        double P[3];
        for (int i = 0; i < 3; ++i)
            P[i] = (((double)Tcw[i * 4] * X + (double)Tcw[i * 4 + 1] * Y) + (double)Tcw[i * 4 + 2] * Z) + (double)Tcw[i * 4 + 3];
        const double fx = pk_.intrinsics[f * 4], fy = pk_.intrinsics[f * 4 + 1], cx = pk_.intrinsics[f * 4 + 2], cy = pk_.intrinsics[f * 4 + 3];
        q[0] = fx * P[0] / P[2] + cx;
        q[1] = fy * P[1] / P[2] + cy;
        return true;
    }

    // ------------------------------------------------------------------ ComputeAlignmentDist
    // iba_global.cpp:111-156; variant 1 = iba_global_stable.cpp:130-176
    void align_dist(int f, const double q[3], AlignResult &r, bool strict, TieStats *ties) const {
        const std::vector<double> &P = scans_[f];
        const Tree3 &tree = *trees_[f];
        uint32_t nn_idx[2]; double nn_d2[2]; bool tie = false;
        knn(tree, q, 1, nn_idx, nn_d2, strict, &tie);
        if (tie && ties) ties->nn3d++;
        r.nn = nn_idx[0];
        const double *nn_pt = &P[(size_t)r.nn * 3];
        const double dq[3] = {nn_pt[0] - q[0], nn_pt[1] - q[1], nn_pt[2] - q[2]};
        const double pt2pt = std::sqrt(dot3(dq, dq));
        r.m = 0; r.is_plane = 0; r.dist = pt2pt;
        r.normal[0] = r.normal[1] = r.normal[2] = 0;
        if (!pr_.use_plane) return;
        uint32_t idx[33]; double d2[33];
        tie = false;
        size_t k = knn(tree, nn_pt, (size_t)pr_.norm_max_pts, idx, d2, strict, &tie);
        const double r2 = pr_.norm_radius * pr_.norm_radius;
        size_t m = 0;
        while (m < k && d2[m] < r2) ++m;  // std::lower_bound over the sorted list
        if (tie && ties) {
            bool t = false;  // only ties that touch the kept set matter
            for (size_t i = 1; i < k && i <= m; ++i) if (d2[i] == d2[i - 1]) t = true;
            if (t) ties->knn3d++;
        }
        r.m = (int32_t)m;
        for (size_t i = 0; i < m; ++i) r.knn[i] = idx[i];
        if (m == 0) return;  // unreachable in the reference (nn_pt is a data point); guards sq_dist[k-1]
        if (pr_.variant == 1) {
            if (m < 3) return;  // iba_global_stable.cpp:154
        } else {
            if (d2[m - 1] < pr_.min_diff_dist * pr_.min_diff_dist) return;
            if ((int)m < pr_.norm_min_pts) return;
        }
        Cumulants cu;
        for (size_t i = 0; i < m; ++i) cu.add(P[(size_t)idx[i] * 3], P[(size_t)idx[i] * 3 + 1], P[(size_t)idx[i] * 3 + 2]);
        double cov[6], nrm[3];
        cu.finish((int)m, cov);
        smallest_eigenvector(cov, nrm);
        normalize3(nrm);
        for (int i = 0; i < 3; ++i) r.normal[i] = nrm[i];
        double reg_err = 0;
        for (size_t i = 0; i < m; ++i) {
            const double *p = &P[(size_t)idx[i] * 3];
            const double d[3] = {p[0] - nn_pt[0], p[1] - nn_pt[1], p[2] - nn_pt[2]};
            reg_err += std::fabs(dot3(d, nrm));
        }
        if (reg_err / (double)(m - 1) > pr_.norm_reg_threshold) return;
        if (pr_.variant == 1) {  // iba_global_stable.cpp:167-171: extent gate after the fit, on the norm
            double max_dist = 0;
            for (size_t i = 0; i < m; ++i) {
                const double *p = &P[(size_t)idx[i] * 3];
                const double d[3] = {p[0] - nn_pt[0], p[1] - nn_pt[1], p[2] - nn_pt[2]};
                max_dist = std::max(std::sqrt(dot3(d, d)), max_dist);
            }
            if (max_dist < pr_.min_diff_dist) return;
        }
        r.is_plane = 1;
        r.dist = std::fabs(dot3(dq, nrm));
    }

    // ------------------------------------------------------------------ one keyframe of BAError
    struct FrameSums {
        double s2d = 0, s3d = 0, she = 0;
        int64_t che = 0, c2d = 0, v2d = 0, c3d = 0, v3d = 0, vpl = 0, vpt = 0, kept = 0, ncorr = 0;
        int64_t q2d = 0, q3d_nn = 0, q3d_knn = 0;  // findNeighbors calls
    };

    // Accumulates keyframe f into acc exactly in the reference's order
    // (iba_global.cpp:194-329).
    void eval_frame(int f, const Rt &Tcl, const Rt &Tlc, double s, FrameSums &acc, bool strict, TieStats *ties,
                    FrameDebug *dbg) const {
        const std::vector<double> &PL = scans_[f];
        const size_t n = PL.size() / 3;
        std::vector<double> PC(n * 3);
        for (size_t i = 0; i < n; ++i) apply(Tcl, &PL[i * 3], &PC[i * 3]);  // TransformPointCloud (pointcloud.h:82-86)
        std::vector<Corr> corrset;
        find_corr(f, PC, corrset, strict, ties);
        if (pr_.variant == 1) {
            for (int64_t k = pk_.kp_offset[f]; k < pk_.kp_offset[f + 1]; ++k) acc.q2d += !std::isnan(pk_.kp_mappoint[(size_t)k * 3]);
        } else {
            acc.q2d += pk_.kp_offset[f + 1] - pk_.kp_offset[f];
        }
        if (dbg) dbg->corr = corrset;
        if ((int)corrset.size() < pr_.num_min_corr) return;  // iba_global.cpp:203
        acc.kept++;
        acc.ncorr += (int64_t)corrset.size();
        const int64_t k0 = pk_.kp_offset[f];
        const float *Tcw = pk_.Tcw + (size_t)f * 12;
        Rt TcwRS;  // Tcw with real size: float32 widened, translation * s (iba_global.cpp:206-208)
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) TcwRS.R[i * 3 + j] = (double)Tcw[i * 4 + j];
            TcwRS.t[i] = (double)Tcw[i * 4 + 3] * s;
        }
        if (pr_.err_weight[1] <= 1e-10) {  // iba_global.cpp:214-219
            acc.s3d = 0;
            acc.c3d++;
            acc.v3d++;
        } else {
            for (const Corr &c : corrset) {
                const float *mp = pk_.kp_mappoint + (size_t)(k0 + c.kp) * 3;
                if (std::isnan(mp[0])) continue;  // mapKpt2Mpt.count == 0
                // GetWorldPos()*scale: CV_32F Mat scaled in float32, then widened (unverified vs OpenCV)
                const float sf = (float)s;
                const double Pw[3] = {(double)(float)(mp[0] * sf), (double)(float)(mp[1] * sf), (double)(float)(mp[2] * sf)};
                double Pc[3], Pl[3];
                for (int i = 0; i < 3; ++i)
                    Pc[i] = ((TcwRS.R[i * 3] * Pw[0] + TcwRS.R[i * 3 + 1] * Pw[1]) + TcwRS.R[i * 3 + 2] * Pw[2]) + TcwRS.t[i];
                apply(Tlc, Pc, Pl);  // Tcl.inverse() * P (iba_global.cpp:234)
                AlignResult r;
                r.kp = c.kp;
                align_dist(f, Pl, r, strict, ties);
                acc.q3d_nn++;
                if (pr_.use_plane) acc.q3d_knn++;
                if (dbg) dbg->align.push_back(r);
                if (r.dist < pr_.corr_3d_3d_threshold) {
                    acc.s3d += r.dist;
                    acc.v3d++;
                    if (r.is_plane) acc.vpl++; else acc.vpt++;
                }
                acc.c3d++;
            }
        }
        // hand-eye term (iba_global.cpp:264-276)
        if (pk_.he_valid[f]) {
            Rt Tc, Tl;
            const float *tc = pk_.he_Tc + (size_t)f * 12;
            const double *tl = pk_.he_Tl + (size_t)f * 12;
            for (int i = 0; i < 3; ++i) {
                for (int j = 0; j < 3; ++j) { Tc.R[i * 3 + j] = (double)tc[i * 4 + j]; Tl.R[i * 3 + j] = tl[i * 4 + j]; }
                Tc.t[i] = (double)tc[i * 4 + 3] * s;
                Tl.t[i] = tl[i * 4 + 3];
            }
            const Rt C1 = compose(Tcl, Tl), C2 = compose(Tc, Tcl);
            double l1[6], l2[6];
            SE3Log(C1.R, C1.t, l1);
            SE3Log(C2.R, C2.t, l2);
            double ss = 0;
            for (int i = 0; i < 6; ++i) { const double d = l1[i] - l2[i]; ss += d * d; }
            acc.she += std::sqrt(ss);
            acc.che++;
        }
        // 3-D/2-D term (iba_global.cpp:291-328)
        const int C = pk_.n_covis;
        const double fx = pk_.intrinsics[f * 4], fy = pk_.intrinsics[f * 4 + 1], cx = pk_.intrinsics[f * 4 + 2], cy = pk_.intrinsics[f * 4 + 3];
        const double W = pk_.image_wh[f * 2], H = pk_.image_wh[f * 2 + 1];
        for (const Corr &c : corrset) {
            const double *p0 = &PC[(size_t)c.pt * 3];
            for (int j = 0; j < C; ++j) {
                if (!pk_.covis_valid[(size_t)f * C + j]) continue;
                const float *uv = pk_.covis_uv + ((size_t)(k0 + c.kp) * C + j) * 2;
                if (std::isnan(uv[0])) continue;  // KptMapList[j].count(kp) == 0
                const double u1 = uv[0], v1 = uv[1];
                const float *rp = pk_.covis_relpose + ((size_t)f * C + j) * 12;
                double p1[3];
                for (int i = 0; i < 3; ++i)
                    p1[i] = (((double)rp[i * 4] * p0[0] + (double)rp[i * 4 + 1] * p0[1]) + (double)rp[i * 4 + 2] * p0[2]) +
                            (double)rp[i * 4 + 3] * s;  // relCVPose.t *= scale (iba_global.cpp:283)
                const double ou = fx * p1[0] / p1[2] + cx;
                const double ov = fy * p1[1] / p1[2] + cy;
                if (!(ou >= 0 && ou < W && ov >= 0 && ov < H)) continue;
                const double err = (ou - u1) * (ou - u1) + (ov - v1) * (ov - v1);
                const double dist = std::sqrt(err);
                if (dist < pr_.corr_3d_2d_threshold) { acc.s2d += dist; acc.v2d++; }
                acc.c2d++;
            }
        }
    }

    // ------------------------------------------------------------------ BAError
    // mode 0: serial over keyframes, one accumulator (Nomad mode, iba_global.cpp:385);
    // mode 1: OpenMP over keyframes (iba_func.cpp:203,463), per-frame partials summed in keyframe order.
    void ba_error(const double x[7], int mode, bool strict, stl_eval_sums_t *out, TieStats *ties, double counters[3]) const {
        double R[9], t[3], s;
        Sim3Exp<double>(x, R, t, s);
        Rt Tcl;
        std::memcpy(Tcl.R, R, sizeof(R));
        std::memcpy(Tcl.t, t, sizeof(t));
        const Rt Tlc = inverse(Tcl);
        const int F = pk_.n_kf;
        FrameSums acc;
        if (mode == 0) {
            for (int f = 0; f < F; ++f) eval_frame(f, Tcl, Tlc, s, acc, strict, ties, nullptr);
        } else {
            std::vector<FrameSums> part(F);
            std::vector<TieStats> tpart(F);
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads_)
            for (int f = 0; f < F; ++f) eval_frame(f, Tcl, Tlc, s, part[f], strict, ties ? &tpart[f] : nullptr, nullptr);
            for (int f = 0; f < F; ++f) {
                const FrameSums &p = part[f];
                if (pr_.err_weight[1] <= 1e-10 && p.kept) acc.s3d = 0; else acc.s3d += p.s3d;
                acc.s2d += p.s2d; acc.she += p.she; acc.che += p.che; acc.c2d += p.c2d; acc.v2d += p.v2d;
                acc.c3d += p.c3d; acc.v3d += p.v3d; acc.vpl += p.vpl; acc.vpt += p.vpt; acc.kept += p.kept; acc.ncorr += p.ncorr;
                acc.q2d += p.q2d; acc.q3d_nn += p.q3d_nn; acc.q3d_knn += p.q3d_knn;
                if (ties) { ties->nn2d += tpart[f].nn2d; ties->nn3d += tpart[f].nn3d; ties->knn3d += tpart[f].knn3d; }
            }
        }
        out->sum_3d2d = acc.s2d; out->sum_3d3d = acc.s3d; out->sum_he = acc.she; out->cnt_he = (double)acc.che;
        out->cnt_3d2d = (double)acc.c2d; out->valid_3d2d = (double)acc.v2d; out->cnt_3d3d = (double)acc.c3d;
        out->valid_3d3d = (double)acc.v3d; out->valid_pl = (double)acc.vpl; out->valid_pt = (double)acc.vpt;
        out->n_frames = (double)acc.kept; out->n_corr = (double)acc.ncorr;
        if (counters) { counters[0] = (double)acc.q2d; counters[1] = (double)acc.q3d_nn; counters[2] = (double)acc.q3d_knn; }
    }

    // Debug: one keyframe, full detail (for index-parity tests).
    void frame_debug(const double x[7], int f, FrameDebug &dbg, TieStats *ties) const {
        double R[9], t[3], s;
        Sim3Exp<double>(x, R, t, s);
        Rt Tcl;
        std::memcpy(Tcl.R, R, sizeof(R));
        std::memcpy(Tcl.t, t, sizeof(t));
        const Rt Tlc = inverse(Tcl);
        FrameSums acc;
        eval_frame(f, Tcl, Tlc, s, acc, false, ties, &dbg);
    }

    void frame_sums(const double x[7], int f, double out[13]) const {
        double R[9], t[3], s;
        Sim3Exp<double>(x, R, t, s);
        Rt Tcl;
        std::memcpy(Tcl.R, R, sizeof(R));
        std::memcpy(Tcl.t, t, sizeof(t));
        const Rt Tlc = inverse(Tcl);
        FrameSums a;
        eval_frame(f, Tcl, Tlc, s, a, false, nullptr, nullptr);
        const double v[13] = {a.s2d, (double)a.v2d, (double)a.c2d, a.she, (double)a.che, (double)a.kept, (double)a.ncorr, (double)a.q3d_nn,
                              a.s3d, (double)a.v3d, (double)a.c3d, (double)a.vpl, (double)a.vpt};
        for (int i = 0; i < 13; ++i) out[i] = v[i];
    }

    // Stand-alone 3-D k-NN on scan f (for the KNN parity tests); (d2, idx) order.
    size_t knn3d(int f, const double q[3], size_t k, uint32_t *idx, double *d2, bool strict, bool *tie) const {
        return knn(*trees_[f], q, k, idx, d2, strict, tie);
    }

    // ------------------------------------------------------------------ LM path: BuildProblem
    // iba_local.cpp:145-323.  Blocks are appended in keyframe order, per keyframe in
    // correspondence order (the reference's order depends on OpenMP scheduling; sums
    // are order-insensitive up to rounding).
    void associate(const double x0[7], bool strict, TieStats *ties) {
        blocks_.clear();
        double R[9], t[3], s0;
        Sim3Exp<double>(x0, R, t, s0);
        Rt T0;
        std::memcpy(T0.R, R, sizeof(R));
        std::memcpy(T0.t, t, sizeof(t));
        const Rt T0inv = inverse(T0);
        const int F = pk_.n_kf, C = pk_.n_covis;
        const double max_3d_dist2 = pr_.max_3d_dist * pr_.max_3d_dist;
        const double r2 = pr_.norm_radius * pr_.norm_radius, mind2 = pr_.min_diff_dist * pr_.min_diff_dist;
        std::vector<std::vector<Block>> per(F);
        std::vector<TieStats> tpart(F);
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads_)
        for (int f = 0; f < F; ++f) {
            TieStats *tf = ties ? &tpart[f] : nullptr;
            const std::vector<double> &PL = scans_[f];
            const size_t n = PL.size() / 3;
            std::vector<double> PC(n * 3);
            for (size_t i = 0; i < n; ++i) apply(T0, &PL[i * 3], &PC[i * 3]);  // iba_local.cpp:168
            std::vector<Corr> corrset;
            find_corr(f, PC, corrset, strict, tf);  // iba_local.cpp:191 (same formulas, incl. fx-for-v)
            if ((int)corrset.size() < pr_.num_min_corr) continue;  // iba_local.cpp:192
            const int64_t k0 = pk_.kp_offset[f];
            const float *Tcwf = pk_.Tcw + (size_t)f * 12;
            const Tree3 &tree = *trees_[f];
            for (const Corr &c : corrset) {
                // ComputeLocalNeighbor (pointcloud.h:733-760) around the scan point
                const double *nn_pt = &PL[(size_t)c.pt * 3];
                uint32_t idx[33]; double d2[33]; bool tie = false;
                size_t k = knn(tree, nn_pt, (size_t)pr_.norm_max_pts, idx, d2, strict, &tie);
                size_t m = 0;
                while (m < k && d2[m] < r2) ++m;
                if (tie && tf) { bool tt = false; for (size_t i = 1; i < k && i <= m; ++i) if (d2[i] == d2[i - 1]) tt = true; if (tt) tf->knn3d++; }
                if ((int)m < pr_.norm_min_pts || m == 0 || d2[m - 1] < mind2) continue;
                const float *mp = pk_.kp_mappoint + (size_t)(k0 + c.kp) * 3;
                if (std::isnan(mp[0])) continue;  // iba_local.cpp:213
                // plane fit over the neighbour points (iba_local.cpp:218-231)
                Cumulants cu;
                for (size_t i = 0; i < m; ++i) cu.add(PL[(size_t)idx[i] * 3], PL[(size_t)idx[i] * 3 + 1], PL[(size_t)idx[i] * 3 + 2]);
                double cov[6], nrm[3];
                cu.finish((int)m, cov);
                smallest_eigenvector(cov, nrm);
                normalize3(nrm);
                double reg_err = 0;
                for (size_t i = 0; i < m; ++i) {
                    const double *p = &PL[(size_t)idx[i] * 3];
                    const double d[3] = {p[0] - nn_pt[0], p[1] - nn_pt[1], p[2] - nn_pt[2]};
                    reg_err += std::fabs(dot3(d, nrm));
                }
                reg_err /= (double)(m - 1);
                const bool valid_plane = reg_err < pr_.norm_reg_threshold;  // strict '<' (iba_local.cpp:231)
                // MapPoint in the reference camera frame, no scale (iba_local.cpp:239-240)
                double MapPoint[3];
                for (int i = 0; i < 3; ++i)
                    MapPoint[i] = (((double)Tcwf[i * 4] * (double)mp[0] + (double)Tcwf[i * 4 + 1] * (double)mp[1]) +
                                   (double)Tcwf[i * 4 + 2] * (double)mp[2]) + (double)Tcwf[i * 4 + 3];
                Block b;
                std::memset(&b, 0, sizeof(b));
                b.kf = f; b.kp = c.kp;
                b.ncov = 0;
                for (int j = 0; j < C; ++j) {
                    if (!pk_.covis_valid[(size_t)f * C + j]) continue;
                    const float *uv = pk_.covis_uv + ((size_t)(k0 + c.kp) * C + j) * 2;
                    if (std::isnan(uv[0])) continue;
                    const float *rp = pk_.covis_relpose + ((size_t)f * C + j) * 12;
                    for (int a = 0; a < 3; ++a) {
                        for (int bb = 0; bb < 3; ++bb) b.R[b.ncov][a * 3 + bb] = (double)rp[a * 4 + bb];
                        b.t[b.ncov][a] = (double)rp[a * 4 + 3];  // unscaled (iba_local.cpp:184-188)
                    }
                    b.u1[b.ncov] = uv[0]; b.v1[b.ncov] = uv[1];
                    b.ncov++;
                }
                if (b.ncov == 0) continue;  // iba_local.cpp:259
                if (valid_plane) {
                    b.type = 0;
                    b.fx = pk_.intrinsics[f * 4]; b.fy = pk_.intrinsics[f * 4 + 1]; b.cx = pk_.intrinsics[f * 4 + 2]; b.cy = pk_.intrinsics[f * 4 + 3];
                    b.u0 = pk_.kp_xy[(k0 + c.kp) * 2]; b.v0 = pk_.kp_xy[(k0 + c.kp) * 2 + 1];
                    for (int i = 0; i < 3; ++i) { b.p0[i] = nn_pt[i]; b.n0[i] = nrm[i]; }
                    per[f].push_back(b);
                } else if (pr_.use_gpr) {
                    // the branch the reference keeps commented out (iba_local.cpp:272-280): depth of the
                    // keypoint by GP regression over the neighbour points, fixed hyper-parameters
                    b.type = 3;
                    b.fx = pk_.intrinsics[f * 4]; b.fy = pk_.intrinsics[f * 4 + 1]; b.cx = pk_.intrinsics[f * 4 + 2]; b.cy = pk_.intrinsics[f * 4 + 3];
                    b.u0 = pk_.kp_xy[(k0 + c.kp) * 2]; b.v0 = pk_.kp_xy[(k0 + c.kp) * 2 + 1];
                    b.gm = (int)m;
                    for (size_t i = 0; i < m; ++i)
                        for (int a = 0; a < 3; ++a) b.gpts[i][a] = PL[(size_t)idx[i] * 3 + a];
                    b.sigma = pr_.gpr_sigma; b.l = pr_.gpr_l; b.sigma_noise = pr_.gpr_sigma_noise;
                    per[f].push_back(b);
                }
                // 3-D term (iba_local.cpp:283-309)
                double Ms[3] = {MapPoint[0] * s0, MapPoint[1] * s0, MapPoint[2] * s0}, Ml[3];
                apply(T0inv, Ms, Ml);
                uint32_t ni[2]; double nd[2]; tie = false;
                knn(tree, Ml, 1, ni, nd, strict, &tie);
                if (tie && tf) tf->nn3d++;
                if (nd[0] > max_3d_dist2) continue;
                const double *NN = &PL[(size_t)ni[0] * 3];
                // ComputeLocalNormalSingleThre (pointcloud.h:699-717,651-666)
                tie = false;
                k = knn(tree, NN, (size_t)pr_.norm_max_pts, idx, d2, strict, &tie);
                m = 0;
                while (m < k && d2[m] < r2) ++m;
                if (tie && tf) { bool tt = false; for (size_t i = 1; i < k && i <= m; ++i) if (d2[i] == d2[i - 1]) tt = true; if (tt) tf->knn3d++; }
                bool state = false;
                double n3[3] = {0, 0, 1};
                if (!((int)m < pr_.norm_min_pts || m == 0 || d2[m - 1] < mind2)) {
                    Cumulants c2;
                    for (size_t i = 0; i < m; ++i) c2.add(PL[(size_t)idx[i] * 3], PL[(size_t)idx[i] * 3 + 1], PL[(size_t)idx[i] * 3 + 2]);
                    c2.finish((int)m, cov);
                    smallest_eigenvector(cov, n3);
                    normalize3(n3);
                    double re = 0;
                    for (size_t i = 0; i < m; ++i) {
                        const double *p = &PL[(size_t)idx[i] * 3];
                        const double d[3] = {p[0] - NN[0], p[1] - NN[1], p[2] - NN[2]};
                        re += std::fabs(dot3(d, n3));
                    }
                    re /= (double)(m - 1);
                    state = re < pr_.norm_reg_threshold;
                }
                Block b3;
                std::memset(&b3, 0, sizeof(b3));
                b3.kf = f; b3.kp = c.kp;
                b3.type = state ? 2 : 1;
                for (int i = 0; i < 3; ++i) { b3.map_pt[i] = MapPoint[i]; b3.query_pt[i] = NN[i]; b3.normal[i] = n3[i]; }
                per[f].push_back(b3);
            }
        }
        for (int f = 0; f < F; ++f) {
            blocks_.insert(blocks_.end(), per[f].begin(), per[f].end());
            if (ties) { ties->nn2d += tpart[f].nn2d; ties->nn3d += tpart[f].nn3d; ties->knn3d += tpart[f].knn3d; }
        }
    }

    const std::vector<Block> &blocks() const { return blocks_; }
    // per-factor hyper-parameters of the GPR blocks (block order), e.g. the ones a GPR::fit produced
    void set_gpr_hyper(const double *sigma_l, size_t n) {
        size_t g = 0;
        for (Block &b : blocks_)
            if (b.type == 3 && g < n) { b.sigma = sigma_l[g * 2]; b.l = sigma_l[g * 2 + 1]; ++g; }
    }
    // training data of GPR::fit for GPR block g at the association extrinsic x0 (IBACalib2.hpp:449-457)
    int gpr_train(size_t g, const double x0[7], double *X /*[32][2]*/, double *y) const {
        size_t k = 0;
        for (const Block &b : blocks_) {
            if (b.type != 3) continue;
            if (k++ != g) continue;
            double R[9], t[3], s;
            Sim3Exp<double>(x0, R, t, s);
            for (int j = 0; j < b.gm; ++j) {
                double pt[3];
                matvec3(R, b.gpts[j], pt);
                for (int a = 0; a < 3; ++a) pt[a] += t[a];
                X[j * 2] = b.fx * pt[0] / pt[2] + b.cx;
                X[j * 2 + 1] = b.fy * pt[1] / pt[2] + b.cy;
                y[j] = pt[2];
            }
            return b.gm;
        }
        return 0;
    }

    // Residuals of one block as Duals (the functors' operator()).
    template <class T>
    static int block_residuals(const Block &b, const T x[7], T *e) {
        if (b.type == 0) {  // IBA_PlaneFactor::operator() (IBACalib2.hpp:152-184)
            T R[9], t[3], s;
            Sim3Exp<T>(x, R, t, s);
            const T fx(b.fx), fy(b.fy), cx(b.cx), cy(b.cy), u0(b.u0), v0(b.v0);
            const T p0[3] = {T(b.p0[0]), T(b.p0[1]), T(b.p0[2])}, n0[3] = {T(b.n0[0]), T(b.n0[1]), T(b.n0[2])};
            T p0c[3], n0c[3];
            matvec3(R, p0, p0c);
            for (int i = 0; i < 3; ++i) p0c[i] = p0c[i] + t[i];
            matvec3(R, n0, n0c);
            const T Cxz = (u0 - cx) / fx, Cyz = (v0 - cy) / fy;
            const T Z0 = dot3(n0c, p0c) / ((Cxz * n0c[0] + Cyz * n0c[1]) + n0c[2]);
            const T P0[3] = {Cxz * Z0, Cyz * Z0, Z0};
            for (int i = 0; i < b.ncov; ++i) {
                T Rm[9], tv[3], P1[3];
                for (int a = 0; a < 9; ++a) Rm[a] = T(b.R[i][a]);
                for (int a = 0; a < 3; ++a) tv[a] = T(b.t[i][a]) * s;  // _t *= _s
                matvec3(Rm, P0, P1);
                for (int a = 0; a < 3; ++a) P1[a] = P1[a] + tv[a];
                e[2 * i] = (fx * P1[0] / P1[2] + cx) - T(b.u1[i]);
                e[2 * i + 1] = (fy * P1[1] / P1[2] + cy) - T(b.v1[i]);
            }
            return 2 * b.ncov;
        }
        if (b.type == 3) {  // IBA_GPRFactor::operator() (IBACalib2.hpp:472-507) + TGPR::fit_predict (GPR.hpp:449-491)
            T R[9], t[3], s;
            Sim3Exp<T>(x, R, t, s);
            const T fx(b.fx), fy(b.fy), cx(b.cx), cy(b.cy), u0(b.u0), v0(b.v0);
            const int n = b.gm;
            std::vector<T> X(2 * n), y(n), K((size_t)n * n), alpha(n), ks(n);
            for (int j = 0; j < n; ++j) {
                const T p[3] = {T(b.gpts[j][0]), T(b.gpts[j][1]), T(b.gpts[j][2])};
                T tf[3];
                matvec3(R, p, tf);
                for (int a = 0; a < 3; ++a) tf[a] = tf[a] + t[a];
                X[2 * j] = fx * tf[0] / tf[2] + cx;
                X[2 * j + 1] = fy * tf[1] / tf[2] + cy;
                y[j] = tf[2];
            }
            const T sigma(b.sigma), l(b.l);
            const T sigma2 = sigma * sigma, coef = T(-0.5) / (l * l);
            // self_pdist (GPR.hpp:41-54) + computeCovariance (GPR.hpp:464-467) + sigma_noise * I
            for (int r = 0; r < n; ++r) {
                K[(size_t)r * n + r] = sigma2 * exp(coef * T(0.0)) + T(b.sigma_noise);
                for (int c2 = r + 1; c2 < n; ++c2) {
                    const T dx = X[2 * r] - X[2 * c2], dy = X[2 * r + 1] - X[2 * c2 + 1];
                    const T kv = sigma2 * exp(coef * (dx * dx + dy * dy));
                    K[(size_t)r * n + c2] = kv;
                    K[(size_t)c2 * n + r] = kv;
                }
            }
            // Eigen::LLT, unblocked lower Cholesky (size < 32), then the two triangular solves
            for (int k = 0; k < n; ++k) {
                T xk = K[(size_t)k * n + k];
                for (int j = 0; j < k; ++j) xk = xk - K[(size_t)k * n + j] * K[(size_t)k * n + j];
                xk = sqrt(xk);
                K[(size_t)k * n + k] = xk;
                for (int i = k + 1; i < n; ++i) {
                    T v = K[(size_t)i * n + k];
                    for (int j = 0; j < k; ++j) v = v - K[(size_t)i * n + j] * K[(size_t)k * n + j];
                    K[(size_t)i * n + k] = v / xk;
                }
            }
            for (int i = 0; i < n; ++i) {
                T v = y[i];
                for (int j = 0; j < i; ++j) v = v - K[(size_t)i * n + j] * alpha[j];
                alpha[i] = v / K[(size_t)i * n + i];
            }
            for (int i = n - 1; i >= 0; --i) {
                T v = alpha[i];
                for (int j = i + 1; j < n; ++j) v = v - K[(size_t)j * n + i] * alpha[j];
                alpha[i] = v / K[(size_t)i * n + i];
            }
            // Kstar = rbf_kernel_2d(train_x, {test_x}) (GPR.hpp:57-63), mu = Kstar^T alpha
            const T inv_l2 = T(1.0) / (l * l);
            T z(0.0);
            for (int j = 0; j < n; ++j) {
                const T dx = X[2 * j] - u0, dy = X[2 * j + 1] - v0;
                ks[j] = sigma2 * exp(T(-0.5) * inv_l2 * (dx * dx + dy * dy));
                z = z + ks[j] * alpha[j];
            }
            const T P0[3] = {z * (u0 - cx) / fx, z * (v0 - cy) / fy, z};
            for (int i = 0; i < b.ncov; ++i) {
                T Rm[9], tv[3], P1[3];
                for (int a = 0; a < 9; ++a) Rm[a] = T(b.R[i][a]);
                for (int a = 0; a < 3; ++a) tv[a] = T(b.t[i][a]) * s;
                matvec3(Rm, P0, P1);
                for (int a = 0; a < 3; ++a) P1[a] = P1[a] + tv[a];
                e[2 * i] = (fx * P1[0] / P1[2] + cx) - T(b.u1[i]);
                e[2 * i + 1] = (fy * P1[1] / P1[2] + cy) - T(b.v1[i]);
            }
            return 2 * b.ncov;
        }
        // Point2Point_Factor / Point2Plane_Factor (IBACalib2.hpp:570-584,611-625)
        const T inv[6] = {-x[0], -x[1], -x[2], -x[3], -x[4], -x[5]};
        T Rlc[9], tlc[3];
        SE3Exp<T>(inv, Rlc, tlc);
        const T s = x[6];
        const T Ms[3] = {T(b.map_pt[0]) * s, T(b.map_pt[1]) * s, T(b.map_pt[2]) * s};
        T M[3];
        matvec3(Rlc, Ms, M);
        for (int i = 0; i < 3; ++i) M[i] = M[i] + tlc[i];
        if (b.type == 1) {
            for (int i = 0; i < 3; ++i) e[i] = M[i] - T(b.query_pt[i]);
            return 3;
        }
        const T d[3] = {M[0] - T(b.query_pt[0]), M[1] - T(b.query_pt[1]), M[2] - T(b.query_pt[2])};
        const T nn[3] = {T(b.normal[0]), T(b.normal[1]), T(b.normal[2])};
        e[0] = dot3(d, nn);
        return 1;
    }

    // Evaluate all frozen blocks at x: Huber-corrected cost, g = J^T r, H = J^T J
    // (Ceres Corrector with rho'' <= 0: residual and Jacobian scaled by sqrt(rho'); A16).
    // Adds blocks [i0, i1) into *out in block order.
    void linearize_range(const double x[7], size_t i0, size_t i1, stl_lin_sums_t *out) const {
        typedef Dual<7> D;
        D xd[7];
        for (int i = 0; i < 7; ++i) xd[i] = D::var(x[i], i);
        for (size_t bi = i0; bi < i1; ++bi) {
            const Block &b = blocks_[bi];
            D e[2 * STL_MAX_COVIS];
            const int nr = block_residuals<D>(b, xd, e);
            double sq = 0;
            for (int i = 0; i < nr; ++i) sq += e[i].a * e[i].a;
            const double delta = (b.type == 0 || b.type == 3) ? pr_.robust_kernel_delta : pr_.robust_kernel_3ddelta;
            double rho0, rho1;  // ceres::HuberLoss::Evaluate
            if (sq > delta * delta) { const double r = std::sqrt(sq); rho0 = 2 * delta * r - delta * delta; rho1 = std::max(std::numeric_limits<double>::min(), delta / r); }
            else { rho0 = sq; rho1 = 1.0; }
            const double sr = std::sqrt(rho1);
            out->cost += 0.5 * rho0;
            for (int i = 0; i < nr; ++i) {
                const double ri = sr * e[i].a;
                double J[7];
                for (int a = 0; a < 7; ++a) J[a] = sr * e[i].v[a];
                for (int a = 0; a < 7; ++a) {
                    out->g[a] += J[a] * ri;
                    for (int c = 0; c < 7; ++c) out->H[a * 7 + c] += J[a] * J[c];
                }
            }
            out->n_residuals += nr;
            if (b.type == 0) out->n_blocks_2d += 1; else if (b.type == 1) out->n_blocks_pt += 1; else if (b.type == 2) out->n_blocks_pl += 1; else out->n_blocks_gpr += 1;
        }
    }

    // serial, one accumulator in block order (the order the golden vectors were written in)
    void linearize(const double x[7], stl_lin_sums_t *out) const {
        std::memset(out, 0, sizeof(*out));
        linearize_range(x, 0, blocks_.size(), out);
    }

    // Residual blocks evaluated by `nthreads` threads, as ceres::Problem::Evaluate does with
    // options.num_threads = hardware_concurrency() (iba_local.cpp:439).  Fixed chunks of 256 blocks, partials
    // added in chunk order: the result does not depend on the thread count (it differs from the serial
    // order above only by the re-association of the fp64 sums).
    void linearize_mt(const double x[7], int nthreads, stl_lin_sums_t *out) const {
        const size_t nb = blocks_.size(), chunk = 256, nc = (nb + chunk - 1) / chunk;
        std::vector<stl_lin_sums_t> part(nc);
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads > 0 ? nthreads : nthreads_)
        for (long long c = 0; c < (long long)nc; ++c) {
            std::memset(&part[c], 0, sizeof(stl_lin_sums_t));
            linearize_range(x, (size_t)c * chunk, std::min(nb, (size_t)(c + 1) * chunk), &part[c]);
        }
        std::memset(out, 0, sizeof(*out));
        double *o = reinterpret_cast<double *>(out);
        for (size_t c = 0; c < nc; ++c) {
            const double *p = reinterpret_cast<const double *>(&part[c]);
            for (int i = 0; i < STL_LIN_NSUMS; ++i) o[i] += p[i];
        }
    }

    const stl_pack_t &pack() const { return pk_; }
    const stl_params_t &params() const { return pr_; }
    const std::vector<double> &scan(int f) const { return scans_[f]; }

  private:
    stl_pack_t pk_;
    stl_params_t pr_;
    int leaf2d_, nthreads_;
    std::vector<std::vector<double>> scans_;
    std::vector<std::unique_ptr<Tree3>> trees_;
    std::vector<Block> blocks_;
};

// BAError epilogue (iba_global.cpp:330-343)
inline void finalize(const stl_params_t &pr, const stl_eval_sums_t &s, stl_ba_error_t *o) {
    if (s.valid_3d2d == 0 && pr.err_weight[0] > 1e-10) o->f1 = DBL_MAX; else o->f1 = s.sum_3d2d / s.valid_3d2d;
    if (s.valid_3d3d == 0 && pr.err_weight[1] > 1e-10) o->f2 = DBL_MAX; else o->f2 = s.sum_3d3d / s.valid_3d3d;
    o->C = s.sum_he / s.cnt_he;
    o->valid_cnt_3d_2d = (int32_t)s.valid_3d2d;
    o->cnt_3d_2d = (int32_t)s.cnt_3d2d;
}

}  // namespace orc
