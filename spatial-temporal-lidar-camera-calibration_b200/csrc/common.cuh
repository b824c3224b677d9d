// common.cuh — device-side layouts and exact-arithmetic helpers shared by the kernels.
//
// HBM layout.  Per keyframe f:
//   px/py/pz[pt_off .. pt_off+n_pad)   scan points, SoA float32, sorted along a 48-bit
//                                       Morton curve, padded with NaN to a multiple of 128
//   orig[pt_off ..)                    original (reference) index of each sorted point
//   node_lo/node_hi[node_off ..)       3-level 32-ary AABB tree over blocks of 32 sorted points:
//                                       [n0 leaves][n1 level-1][n2<=32 level-2], float4 each
//   bitmap[bm_off ..)                  2 px occupancy bitmap of the keypoints dilated by
//                                       max_pixel_dist + fast-path error (1-cell apron)
//   grid_start[grid_off ..), grid_kp   8 px keypoint cell grid (counting-sorted keypoint ids)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace stl {

constexpr int kLeaf = 32;          // points per leaf = one warp-wide load
constexpr int kPadPts = 128;       // per-keyframe point padding (4 points x 32 lanes)
constexpr int kBmCell = 2;         // bitmap cell (px)
constexpr int kGridCell = 8;       // keypoint grid cell (px)
constexpr float kFastErrPx = 0.25f;  // fast-path pixel error budget baked into the bitmap dilation
constexpr int kMaxK = 32;          // k of the k-NN lives in one warp

struct DevKf {
    long long pt_off;
    long long node_off;
    long long kp_off;
    long long bm_off;
    long long grid_off;
    long long mp_off;  // first slot of this keyframe in the map-point-indexed query lists
    long long tab_off; // byte offset of this keyframe's K1 table blob in DevPack::k1tab (16-byte aligned)
    int n_mp;          // keypoints that carry a map point (upper bound of the 3-D queries)
    int tab_bytes;     // size of the blob (multiple of 16)
    int n_pts, n_pad;
    int n0, n1, n2;
    int n_kp;
    int W, H;
    int bm_wpr, bm_rows;
    int gw, gh;
    int he_valid;
    float fx, fy, cx, cy;
    float pmax;  // max |coordinate| of the scan (fast-path error bound)
    uint32_t covis_mask;  // bit j: covisible slot j of this keyframe is valid (DevPack::covis_valid)
};

// K1 table blob of one keyframe: everything the association kernel stages in shared memory, contiguous and 16-byte
// granular so that ONE bulk copy (cp.async.bulk, completion on an mbarrier) brings it in:
//   [float2 kp[n_kp]] [uint32 bitmap[bm_words]] [uint16 grid_start[ncell + 1]] [uint16 grid_kp[n_kp]] [uint32 has_mp[(n_kp + 31) / 32]]
struct K1Tab { int off_bm, off_gs, off_gk, off_mp, bytes; };
__host__ __device__ inline int k1_a16(int x) { return (x + 15) & ~15; }
__host__ __device__ inline K1Tab k1tab_layout(int n_kp, int bm_words, int ncell) {
    K1Tab t;
    t.off_bm = k1_a16(8 * n_kp);
    t.off_gs = t.off_bm + k1_a16(4 * bm_words);
    t.off_gk = t.off_gs + k1_a16(2 * (ncell + 1));
    t.off_mp = t.off_gk + k1_a16(2 * n_kp);
    t.bytes = t.off_mp + k1_a16(4 * ((n_kp + 31) / 32));
    return t;
}

struct DevCand {
    double R[9], t[3];    // Tcl  (Sim3Exp, computed on the host in libm)
    double Ri[9], ti[3];  // Tcl^-1 = [R^T | -(R^T t)]
    double s;
    float sf;             // (float)s, for the float32 map-point scaling
    float pad_;
};

// per-(candidate, keyframe) record written by K1 (fp64 so that one reduction covers all)
struct FrameRec {
    double s2d, v2d, c2d;  // 3-D/2-D term
    double she, che;       // hand-eye term
    double kept, ncorr;    // frame passed the num_min_corr gate / its correspondences
    double nq;             // 3-D queries emitted
};
// per-(candidate, keyframe, sub-block) record written by K2
struct AlignRec {
    double s3d, v3d, c3d, vpl, vpt;
};

// precomputed local plane of one scan point (plane index, optional)
struct PlaneRec {
    double nx, ny, nz, reg;
};

struct DevParams {
    double max_pixel_dist2, thr2d, thr3d, radius2, reg_thr, min_diff2;
    double max_3d_dist2, delta2d, delta3d;
    double w0, w1;
    int num_min_corr, k, min_pts, use_plane;
    int use_gpr, plane_index, variant;
    float adj_r;      // radius the leaf adjacency lists cover, rounded down (0: no lists)
    double min_diff;  // un-squared (variant 1 compares norms)
    double gpr_sigma, gpr_l, gpr_noise;
};

// ---- exact fp64 (no FMA contraction; IEEE div/sqrt) ----------------------------
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }
// (a0*b0 + a1*b1) + a2*b2
__device__ __forceinline__ double dot3e(double a0, double a1, double a2, double b0, double b1, double b2) {
    return dadd(dadd(dmul(a0, b0), dmul(a1, b1)), dmul(a2, b2));
}
// Eigen::Isometry3d * Vector3d (pointcloud.h:85) in the oracle's fixed order: t_i + ((R_i0 x + R_i1 y) + R_i2 z)
__device__ __forceinline__ void xform(const double *R, const double *t, double x, double y, double z, double &ox, double &oy,
                                      double &oz) {
    ox = dadd(t[0], dot3e(R[0], R[1], R[2], x, y, z));
    oy = dadd(t[1], dot3e(R[3], R[4], R[5], x, y, z));
    oz = dadd(t[2], dot3e(R[6], R[7], R[8], x, y, z));
}
// nanoflann L2_Simple_Adaptor::evalMetric (nanoflann.hpp:524-535): ((dx^2 + dy^2) + dz^2), diff = query - data
__device__ __forceinline__ double dist3e(double qx, double qy, double qz, double px, double py, double pz) {
    const double dx = dsub(qx, px), dy = dsub(qy, py), dz = dsub(qz, pz);
    return dadd(dadd(dmul(dx, dx), dmul(dy, dy)), dmul(dz, dz));
}
// Lower bound of dist3e over an axis-aligned box, same operation order and rounding so
// that lb <= dist3e(q, p) for every p in the box (round-to-nearest is monotonic).
__device__ __forceinline__ double box_lb(double qx, double qy, double qz, float4 lo, float4 hi) {
    const double dx = fmax(fmax(dsub((double)lo.x, qx), dsub(qx, (double)hi.x)), 0.0);
    const double dy = fmax(fmax(dsub((double)lo.y, qy), dsub(qy, (double)hi.y)), 0.0);
    const double dz = fmax(fmax(dsub((double)lo.z, qz), dsub(qz, (double)hi.z)), 0.0);
    return dadd(dadd(dmul(dx, dx), dmul(dy, dy)), dmul(dz, dz));
}

__device__ __forceinline__ float ld_nc_f(const float *p) { return __ldg(p); }
__device__ __forceinline__ float4 ld_stream_f4(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

}  // namespace stl
