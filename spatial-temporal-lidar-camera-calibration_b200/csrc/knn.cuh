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
#include <cmath>

#include "common.cuh"
#include "geom.cuh"
#include "kernels.h"

namespace stl {

constexpr unsigned kFull = 0xffffffffu;

struct ScanView {
    const float *px, *py, *pz;
    const uint32_t *orig;
    const float4 *lo, *hi;  // [n0 leaves][n1 level-1][32 level-2]
    const uint16_t *adj;    // [n0][32] leaf adjacency rows (null: none)
    const float *adj_cov;   // [n0] squared distance each row covers
    unsigned long long *stats;  // optional [8] path counters (debug contexts only; null otherwise)
    int n0, n1;
};

__device__ __forceinline__ ScanView make_view(const DevPack &pk, const DevKf &K) {
    ScanView v;
    v.px = pk.px + K.pt_off; v.py = pk.py + K.pt_off; v.pz = pk.pz + K.pt_off;
    v.orig = pk.orig + K.pt_off;
    v.lo = pk.node_lo + K.node_off; v.hi = pk.node_hi + K.node_off;
    v.adj = pk.adj ? pk.adj + K.node_off * 32 : nullptr;
    v.adj_cov = pk.adj ? pk.adj_cov + K.node_off : nullptr;
    v.stats = nullptr;
    v.n0 = K.n0; v.n1 = K.n1;
    return v;
}

// float32 lower bound of the reference's fp64 squared distance to any point of a box.
// Every operation rounds toward the safe side (query rounded outward, differences / squares /
// sums rounded down) and the result is shrunk by 2^-22, far more than the <= 4*2^-53 relative
// rounding of the fp64 distance itself: lbf <= dist3e(q, p) for every p in the box, so pruning
// on it never changes the exact result.  Bounds of the sinks are kept rounded UP in float32.
struct QueryF { float xl, xh, yl, yh, zl, zh; };
__device__ __forceinline__ QueryF make_queryf(double x, double y, double z) {
    QueryF q;
    q.xl = __double2float_rd(x); q.xh = __double2float_ru(x);
    q.yl = __double2float_rd(y); q.yh = __double2float_ru(y);
    q.zl = __double2float_rd(z); q.zh = __double2float_ru(z);
    return q;
}
__device__ __forceinline__ float box_lbf(const QueryF &q, float4 lo, float4 hi) {
    const float dx = fmaxf(fmaxf(__fsub_rd(lo.x, q.xh), __fsub_rd(q.xl, hi.x)), 0.f);
    const float dy = fmaxf(fmaxf(__fsub_rd(lo.y, q.yh), __fsub_rd(q.yl, hi.y)), 0.f);
    const float dz = fmaxf(fmaxf(__fsub_rd(lo.z, q.zh), __fsub_rd(q.zl, hi.z)), 0.f);
    const float s = __fadd_rd(__fadd_rd(__fmul_rd(dx, dx), __fmul_rd(dy, dy)), __fmul_rd(dz, dz));
    return __fmul_rd(s, 0.99999976f);
}

// ---- result sinks ---------------------------------------------------------------
// 1-NN: (d2, orig) lexicographic minimum.  All members are warp-uniform.
// Besides the minimum the sink keeps g2, a float32 LOWER bound of the squared distance from the query to every scan point
// other than the current best that has been looked at or ruled out so far (visited points: their own distance, shrunk by
// 2^-22; unopened boxes: the box bound, see note_pruned / note_bound).  After an exhaustive search g2 bounds the distance
// to the second-nearest point from below; a nearby query (the LM path's twin of an evaluation query) can then be answered
// without a search whenever the gap exceeds the displacement (lm.cu, k_lm_knn_b).
struct Sink1 {
    double d = DBL_MAX;
    float df = 3.402823466e+38f;  // d rounded up to float32
    float g2 = HUGE_VALF;  // +inf
    uint32_t oi = 0xffffffffu, pos = 0xffffffffu;
    __device__ __forceinline__ bool may_contain(float lb) const { return lb <= df; }
    __device__ __forceinline__ static float below(double dd) { return __fmul_rd(__double2float_rd(dd), 0.99999976f); }
    // boxes that were tested and never opened: lane-wise bound `lb`, counted where `pruned`
    __device__ __forceinline__ void note_pruned(float lb, bool pruned) {
        const unsigned k = __reduce_min_sync(kFull, pruned ? __float_as_uint(lb) : 0x7f800000u);
        g2 = fminf(g2, __uint_as_float(k));
    }
    // a bound that holds for every point not covered otherwise (warp-uniform argument)
    __device__ __forceinline__ void note_bound(float b) { g2 = fminf(g2, b); }
    // start from a scan point expected to be close (all lanes read the same address): every box test then
    // already sees a finite bound.  The point is an ordinary candidate of the (d2, index) minimum.
    __device__ __forceinline__ void seed(const ScanView &S, uint32_t p, double qx, double qy, double qz) {
        const double dd = dist3e(qx, qy, qz, (double)S.px[p], (double)S.py[p], (double)S.pz[p]);
        if (dd == dd) { d = dd; df = __double2float_ru(dd); oi = S.orig[p]; pos = p; }
    }
    __device__ __forceinline__ void visit(const ScanView &S, int leaf, double qx, double qy, double qz, int lane) {
        const int g = leaf * kLeaf + lane;
        const double dd = dist3e(qx, qy, qz, (double)S.px[g], (double)S.py[g], (double)S.pz[g]);  // NaN for pads
        const float lowf = dd == dd ? below(dd) : __int_as_float(0x7f800000);
        const bool q = dd <= d;
        if (!__any_sync(kFull, q)) {  // nothing here beats the best: every point of the leaf except the best itself is an "other" point
            note_pruned(lowf, (uint32_t)g != pos);
            return;
        }
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
            if (pos != 0xffffffffu) g2 = fminf(g2, below(d));  // the previous best is an "other" point from now on
            note_pruned(lowf, lane != src);                     // ... and so is everything in this leaf but the winner
            d = cd; oi = mo; pos = (uint32_t)(leaf * kLeaf + src);
            df = __double2float_ru(cd);
        } else {
            note_pruned(lowf, (uint32_t)g != pos);
        }
    }
};

// k-NN (k <= 32) restricted to d2 < r2: lane j holds the j-th best (d2, orig, pos).
struct SinkK {
    double kd = DBL_MAX;       // per lane
    uint32_t ki = 0xffffffffu, kpos = 0xffffffffu;
    int count = 0, k;          // uniform
    double r2, wd;             // uniform: radius^2, current worst d2 when full
    float r2f, wdf;            // the same, rounded up to float32 (box pruning)
    uint32_t wi = 0xffffffffu; // uniform: its index
    __device__ __forceinline__ SinkK(int k_, double r2_) : k(k_), r2(r2_), wd(r2_) {
        r2f = r2_ < 3.0e38 ? __double2float_ru(r2_) : 3.402823466e+38f;
        wdf = r2f;
    }
    __device__ __forceinline__ bool accept(double d, uint32_t i) const {
        return d < r2 && (count < k || d < wd || (d == wd && i < wi));
    }
    __device__ __forceinline__ bool may_contain(float lb) const { return lb <= r2f && (count < k || lb <= wdf); }
    __device__ __forceinline__ void note_pruned(float, bool) {}
    __device__ __forceinline__ void note_bound(float) {}
    __device__ __forceinline__ void visit(const ScanView &S, int leaf, double qx, double qy, double qz, int lane) {
        const int g = leaf * kLeaf + lane;
        const double dd = dist3e(qx, qy, qz, (double)S.px[g], (double)S.py[g], (double)S.pz[g]);
        const bool pre = dd < r2 && (count < k || dd <= wd);
        unsigned mask = __ballot_sync(kFull, pre);
        if (!mask) return;
        const uint32_t o = pre ? S.orig[g] : 0xffffffffu;
        if (__popc(mask) >= (count == 0 ? 4 : 8)) {
            bulk_merge(pre ? dd : (double)INFINITY, o, (uint32_t)g, lane);
            return;
        }
        while (mask) {
            const int src = __ffs(mask) - 1;
            mask &= mask - 1;
            const double cd = __shfl_sync(kFull, dd, src);
            const uint32_t ci = __shfl_sync(kFull, o, src);
            if (!accept(cd, ci)) continue;
            const bool less = lane < count && (kd < cd || (kd == cd && ki < ci));
            const int at = __popc(__ballot_sync(kFull, less));
            const double sd = __shfl_up_sync(kFull, kd, 1);
            const uint32_t si = __shfl_up_sync(kFull, ki, 1), sp = __shfl_up_sync(kFull, kpos, 1);
            if (lane > at) { kd = sd; ki = si; kpos = sp; }
            else if (lane == at) { kd = cd; ki = ci; kpos = (uint32_t)(leaf * kLeaf + src); }
            if (count < k) ++count;
            if (count == k) { wd = __shfl_sync(kFull, kd, k - 1); wi = __shfl_sync(kFull, ki, k - 1); wdf = __double2float_ru(wd); }
        }
    }

    // compare-exchange with the lane `lane ^ j`; keep the smaller (d, i) if keep_min
    __device__ __forceinline__ static void cx(double &d, uint32_t &i, uint32_t &p, int j, bool keep_min) {
        const double od = __shfl_xor_sync(kFull, d, j);
        const uint32_t oi = __shfl_xor_sync(kFull, i, j), op = __shfl_xor_sync(kFull, p, j);
        const bool other_less = od < d || (od == d && oi < i);
        const bool other_more = od > d || (od == d && oi > i);
        if (keep_min ? other_less : other_more) { d = od; i = oi; p = op; }
    }

    // Many candidates at once (typically the first leaves): bitonic-sort the 32 candidates of the
    // leaf across the warp, then bitonic-merge them with the sorted list — ~300 warp instructions
    // whatever the number of accepted points, against ~30 per serial insertion.
    __device__ __forceinline__ void bulk_merge(double cd, uint32_t ci, uint32_t cp, int lane) {
        // (a) sort the 32 candidates on a packed 32-bit key: float32 bits of d2 (rounded down, so the
        //     order is monotone in d2) with the five low bits replaced by the lane — one SHFL and one
        //     MIN/MAX per compare-exchange instead of moving the (d2, index, position) triple.
        unsigned key = cd < (double)INFINITY ? ((__float_as_uint(__double2float_rd(cd)) & ~31u) | (unsigned)lane) : (0xffffffe0u | (unsigned)lane);
#pragma unroll
        for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
            for (int j = kk >> 1; j > 0; j >>= 1) {
                const unsigned o = __shfl_xor_sync(kFull, key, j);
                const bool keep_min = ((lane & kk) == 0) == ((lane & j) == 0);
                key = keep_min ? min(key, o) : max(key, o);
            }
        }
        {   // (b) fetch the payload of the element that belongs at this lane
            const int src = (int)(key & 31u);
            const double sd = __shfl_sync(kFull, cd, src);
            const uint32_t si = __shfl_sync(kFull, ci, src), sp = __shfl_sync(kFull, cp, src);
            cd = sd; ci = si; cp = sp;
        }
        // (c) elements whose truncated keys collide may still be out of exact (d2, index) order:
        //     odd-even transposition on the exact key until stable (almost never entered)
        const unsigned next_key = __shfl_down_sync(kFull, key, 1);  // every lane takes part in the shuffle
        if (__ballot_sync(kFull, lane < 31 && (key >> 5) == (next_key >> 5) && cd < (double)INFINITY)) {
            for (;;) {
                bool moved = false;
#pragma unroll 1
                for (int parity = 0; parity < 2; ++parity) {
                    const bool lower = (lane & 1) == parity;
                    const int partner = lower ? lane + 1 : lane - 1;
                    const bool ok = partner >= 0 && partner < 32;
                    const double od = __shfl_sync(kFull, cd, ok ? partner : lane);
                    const uint32_t oi = __shfl_sync(kFull, ci, ok ? partner : lane), op = __shfl_sync(kFull, cp, ok ? partner : lane);
                    const bool other_less = od < cd || (od == cd && oi < ci);
                    const bool other_more = od > cd || (od == cd && oi > ci);
                    if (ok && (lower ? other_less : other_more)) { cd = od; ci = oi; cp = op; moved = true; }
                }
                if (!__any_sync(kFull, moved)) break;
            }
        }
        if (count == 0) {  // first leaf: the sorted candidates ARE the list
            kd = cd; ki = ci; kpos = cp;
        } else {
            if (lane >= count) { kd = (double)INFINITY; ki = 0xffffffffu; }
            // reversed candidates against the list: element-wise minimum holds the 32 smallest, bitonic
            const double rd = __shfl_sync(kFull, cd, 31 - lane);
            const uint32_t ri = __shfl_sync(kFull, ci, 31 - lane), rp = __shfl_sync(kFull, cp, 31 - lane);
            if (rd < kd || (rd == kd && ri < ki)) { kd = rd; ki = ri; kpos = rp; }
#pragma unroll
            for (int j = 16; j > 0; j >>= 1) cx(kd, ki, kpos, j, (lane & j) == 0);
        }
        const int nvalid = __popc(__ballot_sync(kFull, kd < (double)INFINITY));
        count = nvalid < k ? nvalid : k;
        if (count == k) { wd = __shfl_sync(kFull, kd, k - 1); wi = __shfl_sync(kFull, ki, k - 1); wdf = __double2float_ru(wd); }
    }
};

// ---- traversal -------------------------------------------------------------------
template <class Sink>
__device__ __forceinline__ void traverse(const ScanView &S, double qx, double qy, double qz, Sink &sink, int lane, int first_leaf = -1) {
    // a leaf known to hold a very close point (the query itself for the k-NN searches) is visited
    // first, so that every box test already sees a tight bound; it is skipped in the descent
    if (first_leaf >= 0) sink.visit(S, first_leaf, qx, qy, qz, lane);
    const float4 *lo2 = S.lo + S.n0 + S.n1, *hi2 = S.hi + S.n0 + S.n1;
    const float4 *lo1 = S.lo + S.n0, *hi1 = S.hi + S.n0;
    const QueryF qf = make_queryf(qx, qy, qz);
    const float lb2 = box_lbf(qf, lo2[lane], hi2[lane]);  // empty slots: +inf
    const unsigned key2 = __float_as_uint(lb2);
    unsigned done2 = 0;
    for (;;) {
        const bool c2 = !((done2 >> lane) & 1u) && sink.may_contain(lb2);
        const unsigned m2 = __reduce_min_sync(kFull, c2 ? key2 : 0xffffffffu);
        if (m2 == 0xffffffffu) break;
        const int s2 = __ffs(__ballot_sync(kFull, c2 && key2 == m2)) - 1;
        done2 |= 1u << s2;
        const int n1i = s2 * 32 + lane;
        const float lb1 = box_lbf(qf, lo1[n1i], hi1[n1i]);
        const unsigned key1 = __float_as_uint(lb1);
        unsigned done1 = 0;
        for (;;) {
                const bool c1 = !((done1 >> lane) & 1u) && sink.may_contain(lb1);
            const unsigned m1 = __reduce_min_sync(kFull, c1 ? key1 : 0xffffffffu);
            if (m1 == 0xffffffffu) break;
            const int s1 = __ffs(__ballot_sync(kFull, c1 && key1 == m1)) - 1;
            done1 |= 1u << s1;
            const int n0i = (s2 * 32 + s1) * 32 + lane;
            const float lb0 = box_lbf(qf, S.lo[n0i], S.hi[n0i]);
            const unsigned key0 = __float_as_uint(lb0);
            unsigned done0 = ((s2 * 32 + s1) == (first_leaf >> 5) && first_leaf >= 0) ? (1u << (first_leaf & 31)) : 0u;
            for (;;) {
                        const bool c0 = !((done0 >> lane) & 1u) && sink.may_contain(lb0);
                const unsigned m0 = __reduce_min_sync(kFull, c0 ? key0 : 0xffffffffu);
                if (m0 == 0xffffffffu) break;
                const int s0 = __ffs(__ballot_sync(kFull, c0 && key0 == m0)) - 1;
                done0 |= 1u << s0;
                sink.visit(S, (s2 * 32 + s1) * 32 + s0, qx, qy, qz, lane);
            }
            sink.note_pruned(lb0, !((done0 >> lane) & 1u));
        }
        sink.note_pruned(lb1, !((done1 >> lane) & 1u));
    }
    sink.note_pruned(lb2, !((done2 >> lane) & 1u));
}

// Queries of a CTA are handed out through a shared-memory ticket: a warp that drew short traversals
// takes the next query instead of idling until the slowest warp of the CTA is done.
__device__ __forceinline__ int next_ticket(int *ticket, int lane) {
    int t = 0;
    if (lane == 0) t = atomicAdd(ticket, 1);
    return __shfl_sync(kFull, t, 0);
}

// ---- adjacency scan -----------------------------------------------------------------------
// The nearest leaves around leaf `home` are listed in pk.adj (build.cu, nearest first, one per lane)
// together with the squared distance `cov` the row covers: every point closer than sqrt(cov) to ANY
// point of the home box lives in a listed leaf (cov = +inf: everything within adj_r does).  A search
// whose answer provably lies inside the covered range replaces the 3-level descent by one box test
// per lane.  Returns cov, or -1 when there is no row (the caller then runs the descent).
template <class Sink>
__device__ __forceinline__ float scan_adjacent(const ScanView &S, int home, double qx, double qy, double qz, Sink &sink, int lane) {
    if (S.adj == nullptr) return -1.f;
    const float cov = S.adj_cov[home];
    if (cov < 0.f) return -1.f;
    const unsigned id = S.adj[(long long)home * 32 + lane];
    const QueryF qf = make_queryf(qx, qy, qz);
    const float lb = id != 0xffffu ? box_lbf(qf, S.lo[id], S.hi[id]) : __int_as_float(0x7f800000);
    const unsigned key = __float_as_uint(lb);
    unsigned done = 0;
    for (;;) {  // nearest box first: the bound tightens fastest
        const bool c = !((done >> lane) & 1u) && sink.may_contain(lb);
        const unsigned m = __reduce_min_sync(kFull, c ? key : 0xffffffffu);
        if (m == 0xffffffffu) break;
        const int s = __ffs(__ballot_sync(kFull, c && key == m)) - 1;
        done |= 1u << s;
        sink.visit(S, (int)__shfl_sync(kFull, id, s), qx, qy, qz, lane);
    }
    sink.note_pruned(lb, !((done >> lane) & 1u));  // listed boxes that were never opened (unlisted slots carry +inf)
    return cov;
}

// k-NN (radius-limited) around the scan point at sorted position `pos`.  The adjacency result is final
// when the row covers the whole radius, or when the list is full and its worst distance lies strictly
// inside the covered range (an unlisted point cannot even tie with it); otherwise start over with the descent.
__device__ __forceinline__ void knn_around_point(const ScanView &S, uint32_t pos, SinkK &kn, int lane) {
    const double x = (double)S.px[pos], y = (double)S.py[pos], z = (double)S.pz[pos];
    const float cov = scan_adjacent(S, (int)(pos >> 5), x, y, z, kn, lane);
    const bool done = cov > 3.0e38f || (cov >= 0.f && kn.count == kn.k && kn.wdf < cov);
    if (S.stats && lane == 0) { atomicAdd(S.stats + 0, 1ull); atomicAdd(S.stats + (done ? 1 : (cov >= 0.f ? 2 : 3)), 1ull); }
    if (done) return;
    if (cov >= 0.f) kn = SinkK(kn.k, kn.r2);
    traverse(S, x, y, z, kn, lane, (int)(pos >> 5));
}

// Distance from a query to the box of leaf `home`, rounded up (float32): what the adjacency certificate of a 1-NN adds to the
// nearest distance found (nn_near_leaf).
__device__ __forceinline__ float home_box_delta(const ScanView &S, int home, double qx, double qy, double qz) {
    const float4 lo = S.lo[home], hi = S.hi[home];
    const float xl = __double2float_rd(qx), xh = __double2float_ru(qx), yl = __double2float_rd(qy), yh = __double2float_ru(qy);
    const float zl = __double2float_rd(qz), zh = __double2float_ru(qz);
    const float gx = fmaxf(fmaxf(__fsub_ru(lo.x, xl), __fsub_ru(xh, hi.x)), 0.f);
    const float gy = fmaxf(fmaxf(__fsub_ru(lo.y, yl), __fsub_ru(yh, hi.y)), 0.f);
    const float gz = fmaxf(fmaxf(__fsub_ru(lo.z, zl), __fsub_ru(zh, hi.z)), 0.f);
    return __fsqrt_ru(__fadd_ru(__fadd_ru(__fmul_ru(gx, gx), __fmul_ru(gy, gy)), __fmul_ru(gz, gz)));
}

// 1-NN of an arbitrary query that is expected to lie near leaf `home` (the leaf of the scan point its
// keypoint was associated with).  The adjacency scan is exact when nearest distance + distance from the
// query to the home box stays inside the covered range (all bounds rounded to the safe side); otherwise
// the descent finishes the search with the bound already found (re-visiting a leaf cannot change a
// (d2, index) minimum).
// The sink arrives seeded (or empty) and `delta` = home_box_delta of the query: both are per-query scalars that a caller
// holding one query per lane computes for 32 queries at once (k_nn_knn) before the warp searches them one after another.
__device__ __forceinline__ void nn_near_leaf_seeded(const ScanView &S, float adj_r, int home, double qx, double qy, double qz, Sink1 &nn,
                                                    int lane, float delta) {
    const float cov = scan_adjacent(S, home, qx, qy, qz, nn, lane);
    if (S.stats && lane == 0) atomicAdd(S.stats + 4, 1ull);
    if (cov >= 0.f && nn.pos != 0xffffffffu) {
        const float reach = __fadd_ru(__fsqrt_ru(nn.df), delta);
        if (cov > 3.0e38f ? reach <= adj_r : reach < __fsqrt_rd(cov)) {
            if (S.stats && lane == 0) atomicAdd(S.stats + 5, 1ull);
            // leaves the row does not list lie farther than the covered range from the home box, hence from the query by
            // at least that minus the distance of the query to the home box
            const float u = __fsub_rd(cov > 3.0e38f ? adj_r : __fsqrt_rd(cov), delta);
            nn.note_bound(u > 0.f ? __fmul_rd(u, u) : 0.f);
            return;
        }
        // second chance through the leaf of the point just found (p1, at distance rho): every point as
        // close to the query as p1 lies within 2 rho of p1, hence of p1's leaf box — if that row covers
        // 2 rho, scanning it is exhaustive
        const int l1 = (int)(nn.pos >> 5);
        if (l1 != home) {
            const float cov1 = S.adj_cov[l1];
            const float two_rho = __fmul_ru(2.f, __fsqrt_ru(nn.df));
            if (cov1 >= 0.f && (cov1 > 3.0e38f ? two_rho <= adj_r : two_rho < __fsqrt_rd(cov1))) {
                const float rho_up = __fsqrt_ru(nn.df);  // distance from the query to p1, which lies in leaf l1
                scan_adjacent(S, l1, qx, qy, qz, nn, lane);
                if (S.stats && lane == 0) atomicAdd(S.stats + 5, 1ull);
                const float u = __fsub_rd(cov1 > 3.0e38f ? adj_r : __fsqrt_rd(cov1), rho_up);  // leaves not listed in l1's row
                nn.note_bound(u > 0.f ? __fmul_rd(u, u) : 0.f);
                return;
            }
        }
    }
    if (S.stats && lane == 0) atomicAdd(S.stats + (cov >= 0.f ? 6 : 7), 1ull);
    traverse(S, qx, qy, qz, nn, lane);
}

// `hint` (optional, 0xffffffff = none) is a scan point of leaf `home` believed to be near the query: it seeds the bound.
__device__ __forceinline__ void nn_near_leaf(const ScanView &S, float adj_r, int home, double qx, double qy, double qz, Sink1 &nn,
                                             int lane, uint32_t hint = 0xffffffffu) {
    if (hint != 0xffffffffu) nn.seed(S, hint, qx, qy, qz);
    nn_near_leaf_seeded(S, adj_r, home, qx, qy, qz, nn, lane, home_box_delta(S, home, qx, qy, qz));
}

// ---- local plane around a scan point (one THREAD per neighbourhood) ---------------------
// Given the neighbour list of scan point c (sorted positions in (d2, index) order, m entries,
// `last` = d2 of the m-th), evaluates the gates and the PCA plane exactly as
// ComputeAlignmentDist (iba_global.cpp:130-148) / ComputeLocalNormalSingleThre
// (pointcloud.h:651-666): running sums in neighbour order (ComputeCovariance,
// pointcloud.h:126-158), closed-form smallest eigenvector, regression error.
struct PlaneOut { V3 n; double reg; int m; bool gates_ok; };

// where the coordinates of neighbour j come from: gathered through the position list, or read from
// the transposed float4 rows the traversal kernels write (element (j, slot) at base[j * stride]:
// consecutive threads read consecutive slots, i.e. coalesced 16 B loads instead of three gathers)
struct NbGather {
    const ScanView &S;
    const uint32_t *__restrict__ nb;
    __device__ __forceinline__ void get(int j, double &x, double &y, double &z) const {
        const uint32_t p = nb[j];
        x = (double)S.px[p]; y = (double)S.py[p]; z = (double)S.pz[p];
    }
};
struct NbCoords {
    const float4 *__restrict__ base;
    long long stride;
    __device__ __forceinline__ void get(int j, double &x, double &y, double &z) const {
        const float4 v = base[(long long)j * stride];
        x = (double)v.x; y = (double)v.y; z = (double)v.z;
    }
};
// what a traversal warp stores for lane j of slot `slot`
__device__ __forceinline__ void store_nb_coords(float4 *__restrict__ nbx, long long stride, long long slot, const ScanView &S, int lane,
                                                int count, uint32_t kpos) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < count) v = make_float4(S.px[kpos], S.py[kpos], S.pz[kpos], __uint_as_float(kpos));
    nbx[(long long)lane * stride + slot] = v;
}

template <class Src>
__device__ __forceinline__ PlaneOut plane_fit(const Src &src, int m, double last, double cx, double cy, double cz, const DevParams &pr,
                                              const bool stable = false) {
    PlaneOut out;
    out.m = m;
    out.n = {0.0, 0.0, 0.0};
    out.reg = 0;
    out.gates_ok = false;
    if (stable) {  // iba_global_stable.cpp:154
        if (m < 3) return out;
    } else if (m == 0 || (last < pr.min_diff2) || (m < pr.min_pts)) {
        return out;
    }
    out.gates_ok = true;
    double c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0, c6 = 0, c7 = 0, c8 = 0;
    for (int j = 0; j < m; ++j) {
        double x, y, z;
        src.get(j, x, y, z);
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
        double x, y, z;
        src.get(j, x, y, z);
        const V3 d = {x - cx, y - cy, z - cz};
        reg += fabs(dot(d, n));
    }
    out.n = n;
    out.reg = reg / (double)(m - 1);
    // iba_global_stable.cpp:167-171: after the regression gate, a neighbourhood whose farthest point
    // (norm = sqrt of the k-NN distance) is closer than min_diff_dist is no plane either
    if (stable && sqrt(last) < pr.min_diff) out.reg = __longlong_as_double(0x7ff0000000000000ll);
    return out;
}


__device__ __forceinline__ PlaneOut plane_thread(const ScanView &S, const uint32_t *__restrict__ nb, int m, double last, double cx,
                                                 double cy, double cz, const DevParams &pr, const bool stable = false) {
    return plane_fit(NbGather{S, nb}, m, last, cx, cy, cz, pr, stable);
}

// plane of scan point `pos` from the precomputed index (same PlaneOut as plane_thread)
__device__ __forceinline__ PlaneOut plane_lookup(const DevPack &pk, const DevKf &K, uint32_t pos) {
    const PlaneRec r = pk.pl_rec[K.pt_off + pos];
    const int mm = pk.pl_m[K.pt_off + pos];
    PlaneOut out;
    out.gates_ok = mm >= 0;
    out.m = mm >= 0 ? mm : -(mm + 1);
    out.n = {r.nx, r.ny, r.nz};
    out.reg = r.reg;
    return out;
}

}  // namespace stl
