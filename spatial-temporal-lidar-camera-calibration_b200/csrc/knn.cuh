// knn.cuh — warp-per-query exact nearest-neighbour search over one scan.
//
// Replaces nanoflann's KD-tree search (findNeighbors nanoflann.hpp:1588,
// searchLevel :1736, KNNResultSet :164-237) as used at iba_global.cpp:116-129,
// iba_local.cpp:283-295 and pointcloud.h:733-760.
//
// Index: the scan sorted along a Morton curve; every 32 consecutive points are a
// leaf (one coalesced warp load per SoA array) and two 32-ary levels of AABBs sit
// above the leaves.  A warp tests 32 child boxes per step (one per lane), descends
// the child with the smallest lower bound first (REDUX.MIN over float-ordered bits)
// and prunes with the exact fp64 lower bound, so the result is the exact k-NN under
// the reference's own distance arithmetic (nanoflann.hpp:524-535); ties are ordered
// by original index.  Latency-bound on L2-resident data.
#pragma once
#include <cfloat>

#include "common.cuh"
#include "geom.cuh"
#include "kernels.h"

namespace stl {

constexpr unsigned kFull = 0xffffffffu;

struct ScanView {
    const float *px, *py, *pz;
    const uint32_t *orig;
    const float4 *lo, *hi;  // [n0 leaves][n1 level-1][32 level-2]
    int n0, n1;
};

__device__ __forceinline__ ScanView make_view(const DevPack &pk, const DevKf &K) {
    ScanView v;
    v.px = pk.px + K.pt_off; v.py = pk.py + K.pt_off; v.pz = pk.pz + K.pt_off;
    v.orig = pk.orig + K.pt_off;
    v.lo = pk.node_lo + K.node_off; v.hi = pk.node_hi + K.node_off;
    v.n0 = K.n0; v.n1 = K.n1;
    return v;
}

__device__ __forceinline__ unsigned order_key(double lb) { return __float_as_uint(__double2float_rd(lb)); }

// ---- result sinks ---------------------------------------------------------------
// 1-NN: (d2, orig) lexicographic minimum.  All members are warp-uniform.
struct Sink1 {
    int n_iter = 0, n_visit = 0, n_ins = 0;
    double d = DBL_MAX;
    uint32_t oi = 0xffffffffu, pos = 0xffffffffu;
    __device__ __forceinline__ bool may_contain(double lb) const { return lb <= d; }
    __device__ __forceinline__ void visit(const ScanView &S, int leaf, double qx, double qy, double qz, int lane) {
        ++n_visit;
        const int g = leaf * kLeaf + lane;
        const double dd = dist3e(qx, qy, qz, (double)S.px[g], (double)S.py[g], (double)S.pz[g]);  // NaN for pads
        const bool q = dd <= d;
        if (!__any_sync(kFull, q)) return;
        const uint32_t o = q ? S.orig[g] : 0xffffffffu;
        unsigned hi = q ? (unsigned)__double2hiint(dd) : 0xffffffffu;
        const unsigned mhi = __reduce_min_sync(kFull, hi);
        unsigned lo = (q && hi == mhi) ? (unsigned)__double2loint(dd) : 0xffffffffu;
        const unsigned mlo = __reduce_min_sync(kFull, lo);
        const bool tie = q && hi == mhi && (unsigned)__double2loint(dd) == mlo;
        const unsigned mo = __reduce_min_sync(kFull, tie ? o : 0xffffffffu);
        const double cd = __hiloint2double((int)mhi, (int)mlo);
        if (cd < d || (cd == d && mo < oi)) {
            const int src = __ffs(__ballot_sync(kFull, tie && o == mo)) - 1;
            d = cd; oi = mo; pos = (uint32_t)(leaf * kLeaf + src);
        }
    }
};

// k-NN (k <= 32) restricted to d2 < r2: lane j holds the j-th best (d2, orig, pos).
struct SinkK {
    int n_iter = 0, n_visit = 0, n_ins = 0;
    double kd = DBL_MAX;       // per lane
    uint32_t ki = 0xffffffffu, kpos = 0xffffffffu;
    int count = 0, k;          // uniform
    double r2, wd;             // uniform: radius^2, current worst d2 when full
    uint32_t wi = 0xffffffffu; // uniform: its index
    __device__ __forceinline__ SinkK(int k_, double r2_) : k(k_), r2(r2_), wd(r2_) {}
    __device__ __forceinline__ bool accept(double d, uint32_t i) const {
        return d < r2 && (count < k || d < wd || (d == wd && i < wi));
    }
    __device__ __forceinline__ bool may_contain(double lb) const { return lb < r2 && (count < k || lb <= wd); }
    __device__ __forceinline__ void visit(const ScanView &S, int leaf, double qx, double qy, double qz, int lane) {
        const int g = leaf * kLeaf + lane;
        const double dd = dist3e(qx, qy, qz, (double)S.px[g], (double)S.py[g], (double)S.pz[g]);
        const bool pre = dd < r2 && (count < k || dd <= wd);
        ++n_visit;
        unsigned mask = __ballot_sync(kFull, pre);
        if (!mask) return;
        const uint32_t o = pre ? S.orig[g] : 0xffffffffu;
        while (mask) {
            const int src = __ffs(mask) - 1;
            mask &= mask - 1;
            const double cd = __shfl_sync(kFull, dd, src);
            const uint32_t ci = __shfl_sync(kFull, o, src);
            if (!accept(cd, ci)) continue;
            ++n_ins;
            const bool less = lane < count && (kd < cd || (kd == cd && ki < ci));
            const int at = __popc(__ballot_sync(kFull, less));
            const double sd = __shfl_up_sync(kFull, kd, 1);
            const uint32_t si = __shfl_up_sync(kFull, ki, 1), sp = __shfl_up_sync(kFull, kpos, 1);
            if (lane > at) { kd = sd; ki = si; kpos = sp; }
            else if (lane == at) { kd = cd; ki = ci; kpos = (uint32_t)(leaf * kLeaf + src); }
            if (count < k) ++count;
            if (count == k) { wd = __shfl_sync(kFull, kd, k - 1); wi = __shfl_sync(kFull, ki, k - 1); }
        }
    }
};

// ---- traversal -------------------------------------------------------------------
template <class Sink>
__device__ __forceinline__ void traverse(const ScanView &S, double qx, double qy, double qz, Sink &sink, int lane) {
    const float4 *lo2 = S.lo + S.n0 + S.n1, *hi2 = S.hi + S.n0 + S.n1;
    const float4 *lo1 = S.lo + S.n0, *hi1 = S.hi + S.n0;
    const double lb2 = box_lb(qx, qy, qz, lo2[lane], hi2[lane]);  // empty slots: +inf
    const unsigned key2 = order_key(lb2);
    unsigned done2 = 0;
    for (;;) {
        ++sink.n_iter;
        const bool c2 = !((done2 >> lane) & 1u) && sink.may_contain(lb2);
        const unsigned m2 = __reduce_min_sync(kFull, c2 ? key2 : 0xffffffffu);
        if (m2 == 0xffffffffu) break;
        const int s2 = __ffs(__ballot_sync(kFull, c2 && key2 == m2)) - 1;
        done2 |= 1u << s2;
        const int n1i = s2 * 32 + lane;
        const double lb1 = box_lb(qx, qy, qz, lo1[n1i], hi1[n1i]);
        const unsigned key1 = order_key(lb1);
        unsigned done1 = 0;
        for (;;) {
            ++sink.n_iter;
            const bool c1 = !((done1 >> lane) & 1u) && sink.may_contain(lb1);
            const unsigned m1 = __reduce_min_sync(kFull, c1 ? key1 : 0xffffffffu);
            if (m1 == 0xffffffffu) break;
            const int s1 = __ffs(__ballot_sync(kFull, c1 && key1 == m1)) - 1;
            done1 |= 1u << s1;
            const int n0i = (s2 * 32 + s1) * 32 + lane;
            const double lb0 = box_lb(qx, qy, qz, S.lo[n0i], S.hi[n0i]);
            const unsigned key0 = order_key(lb0);
            unsigned done0 = 0;
            for (;;) {
                ++sink.n_iter;
                const bool c0 = !((done0 >> lane) & 1u) && sink.may_contain(lb0);
                const unsigned m0 = __reduce_min_sync(kFull, c0 ? key0 : 0xffffffffu);
                if (m0 == 0xffffffffu) break;
                const int s0 = __ffs(__ballot_sync(kFull, c0 && key0 == m0)) - 1;
                done0 |= 1u << s0;
                sink.visit(S, (s2 * 32 + s1) * 32 + s0, qx, qy, qz, lane);
            }
        }
    }
}

// ---- local plane around a scan point -------------------------------------------------
// Given the k-NN of `c` (a scan point) held one per lane in (d2, index) order, evaluates
// the gates and the PCA plane exactly as ComputeAlignmentDist (iba_global.cpp:130-148):
// returns true and the unit normal if the neighbourhood is a valid plane.  Sums run
// serially in neighbour order (ComputeCovariance pointcloud.h:126-158), every lane
// redundantly, so there is no cross-lane reduction-order effect.
struct PlaneOut { V3 n; double reg; int m; bool gates_ok; };

__device__ __forceinline__ PlaneOut plane_from_knn(const ScanView &S, const SinkK &kn, double cx, double cy, double cz,
                                                   const DevParams &pr, int lane) {
    PlaneOut out;
    out.m = kn.count;
    out.n = {0.0, 0.0, 0.0};
    out.reg = 0;
    out.gates_ok = false;
    const int m = kn.count;
    if (m == 0) return out;
    const double last = __shfl_sync(kFull, kn.kd, m - 1);
    const bool gate = !(last < pr.min_diff2) && !(m < pr.min_pts);
    float fx = 0.f, fy = 0.f, fz = 0.f;
    if (lane < m) { fx = S.px[kn.kpos]; fy = S.py[kn.kpos]; fz = S.pz[kn.kpos]; }
    if (!gate) return out;
    out.gates_ok = true;
    double c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0, c6 = 0, c7 = 0, c8 = 0;
    for (int j = 0; j < m; ++j) {
        const double x = (double)__shfl_sync(kFull, fx, j), y = (double)__shfl_sync(kFull, fy, j), z = (double)__shfl_sync(kFull, fz, j);
        c0 += x; c1 += y; c2 += z;
        c3 += x * x; c4 += x * y; c5 += x * z;
        c6 += y * y; c7 += y * z; c8 += z * z;
    }
    const double dm = (double)m;
    c0 /= dm; c1 /= dm; c2 /= dm; c3 /= dm; c4 /= dm; c5 /= dm; c6 /= dm; c7 /= dm; c8 /= dm;
    const double cov[6] = {c3 - c0 * c0, c4 - c0 * c1, c5 - c0 * c2, c6 - c1 * c1, c7 - c1 * c2, c8 - c2 * c2};
    const V3 n = normalized(smallest_eigvec(cov));
    double reg = 0;
    for (int j = 0; j < m; ++j) {
        const double x = (double)__shfl_sync(kFull, fx, j), y = (double)__shfl_sync(kFull, fy, j), z = (double)__shfl_sync(kFull, fz, j);
        const V3 d = {x - cx, y - cy, z - cz};
        reg += fabs(dot(d, n));
    }
    out.n = n;
    out.reg = reg / (double)(m - 1);
    return out;
}

}  // namespace stl
