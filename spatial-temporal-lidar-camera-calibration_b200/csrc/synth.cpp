// synth.cpp — seeded synthetic KITTI-shaped KeyFramePack generator (host only).
//
// Test/bench infrastructure shared by the CPU oracle and the CUDA path: both
// consume the SAME bytes.  It replaces, for measurement, everything the
// reference loads from disk before the hot path starts (KITTI .bin scans,
// ORB-SLAM2 keyframes/map, F-LOAM poses: src/examples/iba_global.cpp:473-511)
// following the recipe of SURVEY.md §8(d):
//   * a procedural street scene (ground + lattice of box buildings + parked
//     boxes) ray-cast by a 64-beam spinning LiDAR from every keyframe pose,
//     2 cm range noise, float32 storage (io_tools.h:170-187);
//   * KITTI-00 camera (config/orb_ori/KITTI00-02.yaml:8-11), a KITTI-like GT
//     extrinsic and a monocular scale;
//   * keypoints anchored on projected scan points (+0.5 px noise) or uniform,
//     map points for part of the anchored ones, covisible matches by
//     re-projecting the shared anchor, float32 poses, fp64 LiDAR odometry.
// Every value of keyframe i depends only on (seed, global index i), so a
// keyframe shard generated on another rank is bit-identical to the same slice
// of the full pack.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

#include "../../include/stlcalib.h"
#include "../../include/stlsynth.h"

namespace {

inline uint64_t mix64(uint64_t z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

struct Rng {
    uint64_t s;
    Rng(uint64_t seed, uint64_t a, uint64_t b) { s = mix64(mix64(seed ^ mix64(a)) ^ (b * 0x2545f4914f6cdd1dull)); }
    uint64_t next() {
        s += 0x9e3779b97f4a7c15ull;
        uint64_t z = s;
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        return z ^ (z >> 31);
    }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }  // [0,1)
    double gauss() {
        double u1 = uni(), u2 = uni();
        if (u1 < 1e-300) u1 = 1e-300;
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
    uint32_t below(uint32_t n) { return (uint32_t)(uni() * n); }
};

struct M34 {  // row-major [R|t]
    double m[12];
};
inline M34 mul(const M34 &a, const M34 &b) {
    M34 c;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j)
            c.m[i * 4 + j] = a.m[i * 4] * b.m[j] + a.m[i * 4 + 1] * b.m[4 + j] + a.m[i * 4 + 2] * b.m[8 + j];
        c.m[i * 4 + 3] = a.m[i * 4] * b.m[3] + a.m[i * 4 + 1] * b.m[7] + a.m[i * 4 + 2] * b.m[11] + a.m[i * 4 + 3];
    }
    return c;
}
inline M34 inv(const M34 &a) {
    M34 c;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) c.m[i * 4 + j] = a.m[j * 4 + i];
    for (int i = 0; i < 3; ++i)
        c.m[i * 4 + 3] = -(c.m[i * 4] * a.m[3] + c.m[i * 4 + 1] * a.m[7] + c.m[i * 4 + 2] * a.m[11]);
    return c;
}
inline void apply(const M34 &a, const double p[3], double q[3]) {
    for (int i = 0; i < 3; ++i) q[i] = a.m[i * 4] * p[0] + a.m[i * 4 + 1] * p[1] + a.m[i * 4 + 2] * p[2] + a.m[i * 4 + 3];
}
// float32 product of two float32 poses (what cv::Mat CV_32F operator* yields)
inline void mul_f32(const float a[12], const float b[12], float c[12]) {
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
            float acc = a[i * 4] * b[j];
            acc += a[i * 4 + 1] * b[4 + j];
            acc += a[i * 4 + 2] * b[8 + j];
            c[i * 4 + j] = acc;
        }
        float acc = a[i * 4] * b[3];
        acc += a[i * 4 + 1] * b[7];
        acc += a[i * 4 + 2] * b[11];
        acc += a[i * 4 + 3];
        c[i * 4 + 3] = acc;
    }
}
inline void inv_f32(const float a[12], float c[12]) {  // KeyFrame::SetPose builds Twc = [Rcw^T | -Rcw^T tcw] in float32
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) c[i * 4 + j] = a[j * 4 + i];
    for (int i = 0; i < 3; ++i) {
        float acc = c[i * 4] * a[3];
        acc += c[i * 4 + 1] * a[7];
        acc += c[i * 4 + 2] * a[11];
        c[i * 4 + 3] = -acc;
    }
}

M34 rot_zyx(double yaw, double pitch, double roll, double tx, double ty, double tz) {
    double cy = std::cos(yaw), sy = std::sin(yaw), cp = std::cos(pitch), sp = std::sin(pitch), cr = std::cos(roll),
           sr = std::sin(roll);
    M34 T;
    T.m[0] = cy * cp; T.m[1] = cy * sp * sr - sy * cr; T.m[2] = cy * sp * cr + sy * sr; T.m[3] = tx;
    T.m[4] = sy * cp; T.m[5] = sy * sp * sr + cy * cr; T.m[6] = sy * sp * cr - cy * sr; T.m[7] = ty;
    T.m[8] = -sp;     T.m[9] = cp * sr;                T.m[10] = cp * cr;               T.m[11] = tz;
    return T;
}

// exp of [omega, upsilon] (generator's own; the oracle restates the reference's Sim3Exp separately)
M34 se3_exp(const double x[6]) {
    double th = std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    double O[9] = {0, -x[2], x[1], x[2], 0, -x[0], -x[1], x[0], 0}, O2[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) O2[i * 3 + j] = O[i * 3] * O[j] + O[i * 3 + 1] * O[3 + j] + O[i * 3 + 2] * O[6 + j];
    double a, b, c;
    if (th < 1e-6) { a = 1; b = 0.5; c = 1.0 / 6; }
    else { a = std::sin(th) / th; b = (1 - std::cos(th)) / (th * th); c = (th - std::sin(th)) / (th * th * th); }
    M34 T;
    double V[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double I = (i == j) ? 1.0 : 0.0;
            T.m[i * 4 + j] = I + a * O[i * 3 + j] + b * O2[i * 3 + j];
            V[i * 3 + j] = I + b * O[i * 3 + j] + c * O2[i * 3 + j];
        }
    for (int i = 0; i < 3; ++i) T.m[i * 4 + 3] = V[i * 3] * x[3] + V[i * 3 + 1] * x[4] + V[i * 3 + 2] * x[5];
    return T;
}

// ----------------------------------------------------------------- scene
constexpr double kPitch = 24.0;  // lattice pitch (m)

struct Box { double lo[3], hi[3]; };

inline bool cell_box(uint64_t seed, int ci, int cj, Box &b) {
    if (cj == 0) return false;  // the street
    uint64_t h = mix64(seed ^ mix64((uint64_t)(uint32_t)ci * 0x100000001b3ull + (uint32_t)cj));
    if ((h & 7) == 0) return false;  // empty lot
    double u1 = ((h >> 8) & 0xffff) / 65536.0, u2 = ((h >> 24) & 0xffff) / 65536.0, u3 = ((h >> 40) & 0xffff) / 65536.0;
    double u4 = ((h >> 3) & 0x1f) / 32.0, u5 = ((h >> 56) & 0xff) / 256.0;
    double cx = (ci + 0.5) * kPitch + (u4 - 0.5) * 2.0, cy = cj * kPitch + (u5 - 0.5) * 2.0;
    double hx = 4.0 + 6.0 * u1, hy = 4.0 + 6.0 * u2, hz = 3.0 + 12.0 * u3;
    b.lo[0] = cx - hx; b.hi[0] = cx + hx;
    b.lo[1] = cy - hy; b.hi[1] = cy + hy;
    b.lo[2] = 0; b.hi[2] = hz;
    return true;
}
inline bool cell_car(uint64_t seed, int ci, int cj, Box &b) {
    if (cj != 0) return false;
    uint64_t h = mix64(seed ^ 0xabcdef12345ull ^ mix64((uint64_t)(uint32_t)ci));
    if ((h & 3) == 0) return false;
    double u1 = ((h >> 8) & 0xffff) / 65536.0, u2 = ((h >> 24) & 0xffff) / 65536.0;
    double side = (h & 4) ? 1.0 : -1.0;
    double cx = (ci + 0.2 + 0.6 * u1) * kPitch, cy = side * (5.0 + 2.5 * u2);
    b.lo[0] = cx - 2.2; b.hi[0] = cx + 2.2;
    b.lo[1] = cy - 0.9; b.hi[1] = cy + 0.9;
    b.lo[2] = 0; b.hi[2] = 1.5;
    return true;
}
inline double ray_box(const double o[3], const double inv[3], const Box &b, double tmax) {
    double t0 = 1e-6, t1 = tmax;
    for (int a = 0; a < 3; ++a) {
        double ta = (b.lo[a] - o[a]) * inv[a], tb = (b.hi[a] - o[a]) * inv[a];
        if (ta > tb) { double t = ta; ta = tb; tb = t; }
        if (ta > t0) t0 = ta;
        if (tb < t1) t1 = tb;
        if (t0 > t1) return -1.0;
    }
    return t0;
}

// nearest hit distance along (o, d), or < 0 if none within rmax
double cast(uint64_t seed, const double o[3], const double d[3], double rmax) {
    double best = rmax;
    bool hit = false;
    if (d[2] < -1e-9) {
        double t = -o[2] / d[2];
        if (t < best) { best = t; hit = true; }
    }
    double inv[3];
    for (int a = 0; a < 3; ++a) inv[a] = 1.0 / (std::fabs(d[a]) < 1e-12 ? (d[a] < 0 ? -1e-12 : 1e-12) : d[a]);
    // 2-D DDA over lattice cells; cell (ci, cj) spans x in [ci*P,(ci+1)*P), y in [(cj-0.5)*P,(cj+0.5)*P)
    int ci = (int)std::floor(o[0] / kPitch), cj = (int)std::floor(o[1] / kPitch + 0.5);
    int sx = d[0] > 0 ? 1 : -1, sy = d[1] > 0 ? 1 : -1;
    double nx = ((d[0] > 0 ? ci + 1 : ci) * kPitch - o[0]) * inv[0];
    double ny = (((d[1] > 0 ? cj + 0.5 : cj - 0.5)) * kPitch - o[1]) * inv[1];
    double dx = kPitch * std::fabs(inv[0]), dy = kPitch * std::fabs(inv[1]);
    double tin = 0;
    for (int it = 0; it < 16 && tin < best; ++it) {
        Box b;
        if (cell_box(seed, ci, cj, b)) {
            double t = ray_box(o, inv, b, best);
            if (t > 0 && t < best) { best = t; hit = true; }
        }
        if (cell_car(seed, ci, cj, b)) {
            double t = ray_box(o, inv, b, best);
            if (t > 0 && t < best) { best = t; hit = true; }
        }
        if (nx < ny) { tin = nx; nx += dx; ci += sx; }
        else { tin = ny; ny += dy; cj += sy; }
    }
    return hit ? best : -1.0;
}

struct Synth {
    stl_synth_cfg_t cfg;
    std::vector<int64_t> scan_offset, kp_offset;
    std::vector<float> scan_xyz, intr, kp_xy, kp_mp, Tcw, relpose, covis_uv, he_Tc;
    std::vector<int32_t> wh;
    std::vector<uint8_t> covis_valid, he_valid;
    std::vector<double> he_Tl, Twl;
    stl_pack_t pack;
    double x_gt[7];
};

M34 lidar_pose(const stl_synth_cfg_t &c, int gi) {  // ground-truth T_wl of global keyframe gi
    double x = gi * c.kf_spacing, y = 2.0 * std::sin(x / 40.0);
    double yaw = std::atan(0.05 * std::cos(x / 40.0));
    double pitch = 0.008 * std::sin(x / 13.0), roll = 0.006 * std::cos(x / 17.0);
    return rot_zyx(yaw, pitch, roll, x, y, 1.73);
}

M34 lidar_odom(const stl_synth_cfg_t &c, int gi) {  // noisy fp64 odometry (vTwl)
    M34 T = lidar_pose(c, gi);
    Rng r(c.seed, 7000 + (uint64_t)gi, 3);
    double e[6] = {r.gauss() * 8e-4, r.gauss() * 8e-4, r.gauss() * 8e-4, r.gauss() * 0.01, r.gauss() * 0.01, r.gauss() * 0.01};
    return mul(T, se3_exp(e));
}

void slam_pose_f32(const stl_synth_cfg_t &c, const M34 &Tcl, const M34 &Tlc, const M34 &Tc0w, int gi, float out[12]) {
    // T_ci<-c0 (real) = Tcl * Twl(i)^-1 * Twl(0) * Tlc ; SLAM translation = real / s
    M34 Twc = mul(lidar_pose(c, gi), Tlc);
    M34 T = mul(inv(Twc), inv(Tc0w));  // Tc0w = T_c0<-w  => inv = T_w<-c0
    (void)Tcl;
    for (int i = 0; i < 12; ++i) out[i] = (float)((i % 4 == 3) ? T.m[i] / c.scale_gt : T.m[i]);
}

}  // namespace

extern "C" {

void stl_synth_default_cfg(stl_synth_cfg_t *c) {
    std::memset(c, 0, sizeof(*c));
    c->n_kf = 50; c->kf_begin = 0; c->n_kf_total = 50;
    c->beams = 64; c->az_steps = 1875;
    c->n_kp = 2000; c->n_covis = 3;
    c->anchored_frac = 0.6; c->mappoint_frac = 0.5; c->match_frac = 0.8; c->outlier_frac = 0.05;
    c->scale_gt = 17.3; c->kf_spacing = 1.0; c->max_range = 80.0;
    c->elev_top_deg = 2.0; c->elev_bottom_deg = -24.8;
    c->fx = 718.856f; c->fy = 718.856f; c->cx = 607.1928f; c->cy = 185.2157f;
    c->width = 1241; c->height = 376;
    c->seed = 1000;
    // KITTI-like GT extrinsic: 120 deg about (1,-1,1)/sqrt(3) plus a small generic rotation
    c->x_gt[0] = 1.2091995761561452 + 0.010; c->x_gt[1] = -1.2091995761561452 - 0.015; c->x_gt[2] = 1.2091995761561452 + 0.008;
    c->x_gt[3] = 0.21; c->x_gt[4] = 0.11; c->x_gt[5] = -0.16;
    c->x_gt[6] = c->scale_gt;
}

stl_synth_t *stl_synth_create(const stl_synth_cfg_t *cfg_in) {
    Synth *S = new Synth();
    S->cfg = *cfg_in;
    const stl_synth_cfg_t &c = S->cfg;
    const int F = c.n_kf, C = c.n_covis, K = c.n_kp;
    if (F <= 0 || C < 0 || C > STL_MAX_COVIS || K < 0 || c.beams <= 0 || c.az_steps <= 0) { delete S; return nullptr; }
    std::memcpy(S->x_gt, c.x_gt, sizeof(S->x_gt));
    S->x_gt[6] = c.scale_gt;
    const M34 Tcl = se3_exp(c.x_gt), Tlc = inv(Tcl);
    const M34 Tc0w = inv(mul(lidar_pose(c, 0), Tlc));  // T_c0<-w

    // ---- scans (ragged: rays that hit nothing return no point)
    std::vector<std::vector<float>> scans(F);
#pragma omp parallel for schedule(dynamic, 1)
    for (int f = 0; f < F; ++f) {
        const int gi = c.kf_begin + f;
        const M34 Twl = lidar_pose(c, gi);
        std::vector<float> &out = scans[f];
        out.reserve((size_t)c.beams * c.az_steps * 3);
        Rng rn(c.seed, (uint64_t)gi, 1);
        const double o[3] = {Twl.m[3], Twl.m[7], Twl.m[11]};
        for (int b = 0; b < c.beams; ++b) {
            double el = (c.elev_top_deg + (c.elev_bottom_deg - c.elev_top_deg) * (c.beams > 1 ? (double)b / (c.beams - 1) : 0.0)) *
                        0.017453292519943295;
            double ce = std::cos(el), se = std::sin(el);
            double az0 = rn.uni() * 6.283185307179586 / c.az_steps;
            for (int a = 0; a < c.az_steps; ++a) {
                double az = az0 + 6.283185307179586 * a / c.az_steps;
                double dl[3] = {ce * std::cos(az), ce * std::sin(az), se};
                double dw[3];
                for (int i = 0; i < 3; ++i) dw[i] = Twl.m[i * 4] * dl[0] + Twl.m[i * 4 + 1] * dl[1] + Twl.m[i * 4 + 2] * dl[2];
                double t = cast(c.seed, o, dw, c.max_range);
                double noise = rn.gauss() * 0.02;  // consumed for every ray: stream position independent of hits
                if (t < 2.0) continue;              // no return / inside the vehicle's blind zone
                double r = t + noise;
                out.push_back((float)(dl[0] * r));
                out.push_back((float)(dl[1] * r));
                out.push_back((float)(dl[2] * r));
            }
        }
    }
    S->scan_offset.assign(F + 1, 0);
    for (int f = 0; f < F; ++f) S->scan_offset[f + 1] = S->scan_offset[f] + (int64_t)scans[f].size() / 3;
    S->scan_xyz.resize((size_t)S->scan_offset[F] * 3);
#pragma omp parallel for schedule(static)
    for (int f = 0; f < F; ++f) {
        if (!scans[f].empty()) std::memcpy(&S->scan_xyz[(size_t)S->scan_offset[f] * 3], scans[f].data(), scans[f].size() * sizeof(float));
        std::vector<float>().swap(scans[f]);
    }

    // ---- per-keyframe camera state
    S->kp_offset.assign(F + 1, 0);
    for (int f = 0; f < F; ++f) S->kp_offset[f + 1] = S->kp_offset[f] + K;
    const size_t NK = (size_t)F * K;
    S->intr.resize((size_t)F * 4); S->wh.resize((size_t)F * 2);
    S->kp_xy.resize(NK * 2); S->kp_mp.resize(NK * 3);
    S->Tcw.resize((size_t)F * 12); S->relpose.assign((size_t)F * C * 12, 0.f);
    S->covis_valid.assign((size_t)F * C, 0); S->covis_uv.resize(NK * C * 2);
    S->he_Tc.assign((size_t)F * 12, 0.f); S->he_Tl.assign((size_t)F * 12, 0.0); S->he_valid.assign(F, 0);
    S->Twl.resize((size_t)F * 12);
    const float qnan = std::numeric_limits<float>::quiet_NaN();
    const double fx = c.fx, fy = c.fy, cx = c.cx, cy = c.cy, W = c.width, H = c.height;
    static const int covis_delta[STL_MAX_COVIS] = {1, -1, 2, -2, 3, -3, 4, -4, 5, -5};

#pragma omp parallel for schedule(dynamic, 1)
    for (int f = 0; f < F; ++f) {
        const int gi = c.kf_begin + f;
        S->intr[f * 4 + 0] = c.fx; S->intr[f * 4 + 1] = c.fy; S->intr[f * 4 + 2] = c.cx; S->intr[f * 4 + 3] = c.cy;
        S->wh[f * 2] = c.width; S->wh[f * 2 + 1] = c.height;
        float *Tcw = &S->Tcw[(size_t)f * 12];
        slam_pose_f32(c, Tcl, Tlc, Tc0w, gi, Tcw);
        float Twc_f[12];
        inv_f32(Tcw, Twc_f);
        M34 odo = lidar_odom(c, gi);
        std::memcpy(&S->Twl[(size_t)f * 12], odo.m, sizeof(odo.m));
        if (gi + 1 < c.n_kf_total) {
            S->he_valid[f] = 1;
            float Tn[12];
            slam_pose_f32(c, Tcl, Tlc, Tc0w, gi + 1, Tn);
            mul_f32(Tn, Twc_f, &S->he_Tc[(size_t)f * 12]);
            M34 Tl = mul(inv(lidar_odom(c, gi + 1)), odo);
            std::memcpy(&S->he_Tl[(size_t)f * 12], Tl.m, sizeof(Tl.m));
        }
        // real-scale relative poses camera i -> covisible camera j (ground truth, fp64) for anchor re-projection
        M34 Twci = mul(lidar_pose(c, gi), Tlc);
        M34 Tji[STL_MAX_COVIS];
        for (int s = 0; s < C; ++s) {
            int gj = gi + covis_delta[s];
            if (gj < 0 || gj >= c.n_kf_total) continue;
            S->covis_valid[(size_t)f * C + s] = 1;
            float Tj[12];
            slam_pose_f32(c, Tcl, Tlc, Tc0w, gj, Tj);
            mul_f32(Tj, Twc_f, &S->relpose[((size_t)f * C + s) * 12]);
            Tji[s] = mul(inv(mul(lidar_pose(c, gj), Tlc)), Twci);
        }
        // in-FoV scan points under the GT extrinsic
        const float *P = &S->scan_xyz[(size_t)S->scan_offset[f] * 3];
        const int64_t N = S->scan_offset[f + 1] - S->scan_offset[f];
        std::vector<uint32_t> fov;
        fov.reserve((size_t)N / 6);
        for (int64_t i = 0; i < N; ++i) {
            double p[3] = {P[i * 3], P[i * 3 + 1], P[i * 3 + 2]}, q[3];
            apply(Tcl, p, q);
            if (q[2] < 1.0) continue;
            double u = fx * q[0] / q[2] + cx, v = fy * q[1] / q[2] + cy;
            if (u >= 2 && u < W - 2 && v >= 2 && v < H - 2) fov.push_back((uint32_t)i);
        }
        Rng rk(c.seed, 2000 + (uint64_t)gi, 2);
        M34 Tc0ci = mul(Tc0w, Twci);  // camera i (real) -> camera 0 (real)
        for (int k = 0; k < K; ++k) {
            const size_t kk = (size_t)f * K + k;
            float *kp = &S->kp_xy[kk * 2], *mp = &S->kp_mp[kk * 3], *cuv = &S->covis_uv[kk * C * 2];
            mp[0] = mp[1] = mp[2] = qnan;
            for (int s = 0; s < C; ++s) cuv[s * 2] = cuv[s * 2 + 1] = qnan;
            bool anchored = !fov.empty() && rk.uni() < c.anchored_frac;
            if (anchored) {
                uint32_t i = fov[rk.below((uint32_t)fov.size())];
                double p[3] = {P[i * 3], P[i * 3 + 1], P[i * 3 + 2]}, q[3];
                apply(Tcl, p, q);
                double u = fx * q[0] / q[2] + cx + rk.gauss() * 0.5, v = fy * q[1] / q[2] + cy + rk.gauss() * 0.5;
                u = std::fmin(std::fmax(u, 0.0), W - 1.0); v = std::fmin(std::fmax(v, 0.0), H - 1.0);
                kp[0] = (float)u; kp[1] = (float)v;
                if (rk.uni() < c.mappoint_frac) {
                    double qn[3] = {q[0] + rk.gauss() * 0.02, q[1] + rk.gauss() * 0.02, q[2] + rk.gauss() * 0.02}, w[3];
                    apply(Tc0ci, qn, w);
                    for (int a = 0; a < 3; ++a) mp[a] = (float)(w[a] / c.scale_gt);
                }
                for (int s = 0; s < C; ++s) {
                    if (!S->covis_valid[(size_t)f * C + s]) continue;
                    double r1 = rk.uni(), g1 = rk.gauss(), g2 = rk.gauss();
                    if (r1 >= c.match_frac) continue;
                    double qj[3];
                    apply(Tji[s], q, qj);
                    if (qj[2] < 0.5) continue;
                    double u1 = fx * qj[0] / qj[2] + cx + g1 * 0.5, v1 = fy * qj[1] / qj[2] + cy + g2 * 0.5;
                    if (u1 < 0 || u1 >= W || v1 < 0 || v1 >= H) continue;
                    cuv[s * 2] = (float)u1; cuv[s * 2 + 1] = (float)v1;
                }
            } else {
                kp[0] = (float)(rk.uni() * (W - 1)); kp[1] = (float)(rk.uni() * (H - 1));
                for (int s = 0; s < C; ++s) {
                    double r1 = rk.uni(), r2 = rk.uni(), r3 = rk.uni();
                    if (!S->covis_valid[(size_t)f * C + s] || r1 >= c.outlier_frac) continue;
                    cuv[s * 2] = (float)(r2 * (W - 1)); cuv[s * 2 + 1] = (float)(r3 * (H - 1));
                }
                if (rk.uni() < 0.1) {  // a few map points with no scan support behind them
                    double qn[3] = {(kp[0] - cx) / fx * 20.0, (kp[1] - cy) / fy * 20.0, 20.0 + rk.uni() * 20.0}, w[3];
                    apply(Tc0ci, qn, w);
                    for (int a = 0; a < 3; ++a) mp[a] = (float)(w[a] / c.scale_gt);
                }
            }
        }
    }

    stl_pack_t &p = S->pack;
    std::memset(&p, 0, sizeof(p));
    p.n_kf = F; p.n_covis = C;
    p.scan_offset = S->scan_offset.data(); p.scan_xyz = S->scan_xyz.data();
    p.intrinsics = S->intr.data(); p.image_wh = S->wh.data();
    p.kp_offset = S->kp_offset.data(); p.kp_xy = S->kp_xy.data(); p.kp_mappoint = S->kp_mp.data();
    p.Tcw = S->Tcw.data(); p.covis_relpose = S->relpose.data(); p.covis_valid = S->covis_valid.data();
    p.covis_uv = S->covis_uv.data(); p.he_Tc = S->he_Tc.data(); p.he_Tl = S->he_Tl.data(); p.he_valid = S->he_valid.data();
    return reinterpret_cast<stl_synth_t *>(S);
}

const stl_pack_t *stl_synth_pack(const stl_synth_t *h) { return &reinterpret_cast<const Synth *>(h)->pack; }
const double *stl_synth_x_gt(const stl_synth_t *h) { return reinterpret_cast<const Synth *>(h)->x_gt; }
const double *stl_synth_Twl(const stl_synth_t *h) { return reinterpret_cast<const Synth *>(h)->Twl.data(); }
void stl_synth_destroy(stl_synth_t *h) { delete reinterpret_cast<Synth *>(h); }

// Candidates around the GT: x0 = x_gt + U(lb,ub) * spread, lb/ub from
// config/calib/00/iba_calib_global.yml:38-39.  out is [B][7].
void stl_synth_candidates(const double x_gt[7], uint64_t seed, int32_t B, double spread, double *out) {
    static const double ub[7] = {0.1, 0.1, 0.1, 0.3, 0.3, 0.3, 1.0};
    for (int b = 0; b < B; ++b) {
        Rng r(seed, 42, (uint64_t)b);
        for (int i = 0; i < 7; ++i) out[b * 7 + i] = x_gt[i] + (b == 0 ? 0.0 : (2 * r.uni() - 1) * ub[i] * spread);
    }
}

}  // extern "C"
