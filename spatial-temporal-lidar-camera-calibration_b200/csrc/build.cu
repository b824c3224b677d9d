// build.cu — K0: one-off per-scan 3-D index build (replaces the KDTree3D-per-scan
// construction of BALoss's constructor, src/examples/iba_global.cpp:362-367).
//
// Candidate-independent.  For a chunk of keyframes whose raw float32 points are on
// the device:  bbox -> 48-bit Morton key (16 bit/axis, isotropic cell) tagged with
// the keyframe -> one stable radix sort (cub) -> scatter into the padded SoA layout
// -> AABBs of every block of 32 points, then two 32-ary levels above them.
// HBM-bound streaming; algorithmic bytes ~ 12 B read + 16 B written per point.
#include <cub/cub.cuh>

#include "kernels.h"

namespace stl {
namespace {

__device__ __forceinline__ unsigned long long spread16(unsigned v) {  // abcd -> a00b00c00d
    unsigned long long x = v & 0xffffull;
    x = (x | (x << 16)) & 0x0000ff0000ffull;
    x = (x | (x << 8)) & 0x00f00f00f00full;
    x = (x | (x << 4)) & 0x0c30c30c30c3ull;
    x = (x | (x << 2)) & 0x249249249249ull;
    return x;
}

// one block per keyframe: bbox + pmax
__global__ void k_bbox(const float *__restrict__ raw, const long long *__restrict__ raw_off, float *__restrict__ bbox, DevKf *kf,
                       int kf_begin) {
    const int f = blockIdx.x;
    const long long b = raw_off[f], e = raw_off[f + 1];
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (long long i = b + threadIdx.x; i < e; i += blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = raw[i * 3 + a];
            lo[a] = fminf(lo[a], v);
            hi[a] = fmaxf(hi[a], v);
        }
    }
    __shared__ float slo[3][32], shi[3][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        for (int o = 16; o; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
        if (lane == 0) { slo[a][w] = lo[a]; shi[a][w] = hi[a]; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = blockDim.x >> 5;
        float pm = 0.f;
        for (int a = 0; a < 3; ++a) {
            float l = slo[a][0], h = shi[a][0];
            for (int i = 1; i < nw; ++i) { l = fminf(l, slo[a][i]); h = fmaxf(h, shi[a][i]); }
            bbox[f * 6 + a] = l;
            bbox[f * 6 + 3 + a] = h;
            pm = fmaxf(pm, fmaxf(fabsf(l), fabsf(h)));
        }
        if (!(pm >= 1.f)) pm = 1.f;
        kf[kf_begin + f].pmax = pm;
    }
}

// grid (tiles, nkf)
__global__ void k_morton(const float *__restrict__ raw, const long long *__restrict__ raw_off, const float *__restrict__ bbox,
                         unsigned long long *__restrict__ keys, uint32_t *__restrict__ vals) {
    const int f = blockIdx.y;
    const long long b = raw_off[f], e = raw_off[f + 1];
    const float lx = bbox[f * 6], ly = bbox[f * 6 + 1], lz = bbox[f * 6 + 2];
    float ext = fmaxf(fmaxf(bbox[f * 6 + 3] - lx, bbox[f * 6 + 4] - ly), bbox[f * 6 + 5] - lz);
    if (!(ext > 0.f)) ext = 1.f;
    const float scale = 65535.0f / ext;
    for (long long i = b + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < e; i += (long long)gridDim.x * blockDim.x) {
        const float x = raw[i * 3], y = raw[i * 3 + 1], z = raw[i * 3 + 2];
        const unsigned qx = (unsigned)fminf(fmaxf((x - lx) * scale, 0.f), 65535.f);
        const unsigned qy = (unsigned)fminf(fmaxf((y - ly) * scale, 0.f), 65535.f);
        const unsigned qz = (unsigned)fminf(fmaxf((z - lz) * scale, 0.f), 65535.f);
        const unsigned long long m = spread16(qx) | (spread16(qy) << 1) | (spread16(qz) << 2);
        keys[i] = ((unsigned long long)f << 48) | m;
        vals[i] = (uint32_t)(i - b);
    }
}

// grid (tiles, nkf): sorted rank j of keyframe f -> padded SoA slot
__global__ void k_scatter(const float *__restrict__ raw, const long long *__restrict__ raw_off, const uint32_t *__restrict__ vals,
                          const DevKf *__restrict__ kf, int kf_begin, float *__restrict__ px, float *__restrict__ py,
                          float *__restrict__ pz, uint32_t *__restrict__ orig) {
    const int f = blockIdx.y;
    const DevKf K = kf[kf_begin + f];
    const long long b = raw_off[f];
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < K.n_pad; j += gridDim.x * blockDim.x) {
        const long long dst = K.pt_off + j;
        if (j < K.n_pts) {
            const uint32_t v = vals[b + j];
            const float *p = raw + (b + v) * 3;
            px[dst] = p[0]; py[dst] = p[1]; pz[dst] = p[2];
            orig[dst] = v;
        } else {
            const float qn = __int_as_float(0x7fc00000);
            px[dst] = qn; py[dst] = qn; pz[dst] = qn;
            orig[dst] = 0xffffffffu;
        }
    }
}


// ---------------------------------------------------------------------------------------------
// KD refinement of the Morton order.  Blocks of 32 Morton-consecutive points make poor leaves
// (the curve jumps; LiDAR returns lie on 2-D surfaces), so every aligned group of 8192 sorted
// points (= 256 leaves = 8 of the 32-ary level-1 nodes) is re-ordered in shared memory by eight
// rounds of "sort the segment along the longest axis of its bounding box, split in the middle"
// — a balanced KD-tree whose cells are exactly the implicit 32-point leaves and 1024-point
// level-1 nodes.  Halves the boxes an exact query has to open (profiles/r01_leaf_quality.txt).
constexpr int kGroup = 8192;
constexpr int kRefThreads = 1024;
constexpr int kGroupLevels = 8;  // 8192 -> 32

__device__ __forceinline__ int f2ord(float f) { const int b = __float_as_int(f); return b >= 0 ? b : b ^ 0x7fffffff; }

struct RefineSmem {
    float x[kGroup], y[kGroup], z[kGroup];
    uint32_t o[kGroup];
    float key[kGroup];
    unsigned short perm[kGroup];
    int bb[128][6];
    unsigned char axis[128];
};

__device__ __forceinline__ void cmp_swap(RefineSmem &S, int i, int l) {
    const float ki = S.key[i], kl = S.key[l];
    const unsigned short pi = S.perm[i], pl = S.perm[l];
    if (ki > kl || (ki == kl && pi > pl)) { S.key[i] = kl; S.key[l] = ki; S.perm[i] = pl; S.perm[l] = pi; }
}

// grid (groups, nkf)
__global__ void __launch_bounds__(kRefThreads, 1)
k_kd_refine(const DevKf *__restrict__ kf, int kf_begin, float *__restrict__ px, float *__restrict__ py, float *__restrict__ pz,
            uint32_t *__restrict__ orig) {
    extern __shared__ __align__(16) unsigned char refine_raw[];
    RefineSmem &S = *reinterpret_cast<RefineSmem *>(refine_raw);
    const DevKf K = kf[kf_begin + blockIdx.y];
    const int g0 = blockIdx.x * kGroup;
    if (g0 >= K.n_pad) return;
    const int M = min(kGroup, K.n_pad - g0);
    const long long base = K.pt_off + g0;
    const int tid = threadIdx.x;
    const float qn = __int_as_float(0x7fc00000);
    for (int i = tid; i < kGroup; i += kRefThreads) {
        const bool in = i < M;
        S.x[i] = in ? px[base + i] : qn;
        S.y[i] = in ? py[base + i] : qn;
        S.z[i] = in ? pz[base + i] : qn;
        S.o[i] = in ? orig[base + i] : 0xffffffffu;
        S.perm[i] = (unsigned short)i;
    }
    __syncthreads();
    for (int lev = 0; lev < kGroupLevels; ++lev) {
        const int seg_size = kGroup >> lev, nseg = 1 << lev, seg_shift = 13 - lev;
        for (int i = tid; i < nseg * 6; i += kRefThreads) S.bb[i / 6][i % 6] = (i % 6) < 3 ? 0x7fffffff : (int)0x80000000;
        __syncthreads();
        for (int i = tid; i < kGroup; i += kRefThreads) {
            const int p = S.perm[i];
            const float x = S.x[p];
            if (x == x) {
                const int sg = i >> seg_shift;
                atomicMin(&S.bb[sg][0], f2ord(x)); atomicMax(&S.bb[sg][3], f2ord(x));
                atomicMin(&S.bb[sg][1], f2ord(S.y[p])); atomicMax(&S.bb[sg][4], f2ord(S.y[p]));
                atomicMin(&S.bb[sg][2], f2ord(S.z[p])); atomicMax(&S.bb[sg][5], f2ord(S.z[p]));
            }
        }
        __syncthreads();
        if (tid < nseg) {
            int ax = 0;
            if (S.bb[tid][0] <= S.bb[tid][3]) {  // at least one real point
                float e[3];
                for (int a = 0; a < 3; ++a) {
                    const int lo = S.bb[tid][a], hi = S.bb[tid][3 + a];
                    const float fl = __int_as_float(lo >= 0 ? lo : lo ^ 0x7fffffff), fh = __int_as_float(hi >= 0 ? hi : hi ^ 0x7fffffff);
                    e[a] = fh - fl;
                }
                if (e[1] > e[ax]) ax = 1;
                if (e[2] > e[ax]) ax = 2;
            }
            S.axis[tid] = (unsigned char)ax;
        }
        __syncthreads();
        for (int i = tid; i < kGroup; i += kRefThreads) {
            const int p = S.perm[i], ax = S.axis[i >> seg_shift];
            const float v = ax == 0 ? S.x[p] : (ax == 1 ? S.y[p] : S.z[p]);
            S.key[i] = (v == v) ? v : INFINITY;
        }
        __syncthreads();
        // all-ascending bitonic network, independent inside every aligned block of seg_size
        for (int k = 2; k <= seg_size; k <<= 1) {
            const int hk = k >> 1;
            for (int p = tid; p < kGroup / 2; p += kRefThreads) {  // flip step: i <-> i ^ (k - 1)
                const int i = (p / hk) * k + (p % hk);
                cmp_swap(S, i, i ^ (k - 1));
            }
            __syncthreads();
            for (int j = hk >> 1; j > 0; j >>= 1) {  // disperse steps: i <-> i + j
                for (int p = tid; p < kGroup / 2; p += kRefThreads) {
                    const int i = (p / j) * 2 * j + (p % j);
                    cmp_swap(S, i, i + j);
                }
                __syncthreads();
            }
        }
    }
    for (int i = tid; i < M; i += kRefThreads) {
        const int p = S.perm[i];
        px[base + i] = S.x[p]; py[base + i] = S.y[p]; pz[base + i] = S.z[p];
        orig[base + i] = S.o[p];
    }
}

__device__ __forceinline__ void warp_box(float4 &lo, float4 &hi) {
    for (int o = 16; o; o >>= 1) {
        lo.x = fminf(lo.x, __shfl_xor_sync(0xffffffffu, lo.x, o));
        lo.y = fminf(lo.y, __shfl_xor_sync(0xffffffffu, lo.y, o));
        lo.z = fminf(lo.z, __shfl_xor_sync(0xffffffffu, lo.z, o));
        hi.x = fmaxf(hi.x, __shfl_xor_sync(0xffffffffu, hi.x, o));
        hi.y = fmaxf(hi.y, __shfl_xor_sync(0xffffffffu, hi.y, o));
        hi.z = fmaxf(hi.z, __shfl_xor_sync(0xffffffffu, hi.z, o));
    }
}

// one warp per leaf slot (n0 per keyframe); grid (ceil(max_n0/warps), nkf)
__global__ void k_leaf_aabb(const DevKf *__restrict__ kf, int kf_begin, const float *__restrict__ px, const float *__restrict__ py,
                            const float *__restrict__ pz, float4 *__restrict__ node_lo, float4 *__restrict__ node_hi) {
    const DevKf K = kf[kf_begin + blockIdx.y];
    const int lane = threadIdx.x & 31;
    const int leaf = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (leaf >= K.n0) return;
    float4 lo = make_float4(INFINITY, INFINITY, INFINITY, 0.f), hi = make_float4(-INFINITY, -INFINITY, -INFINITY, 0.f);
    const int j = leaf * kLeaf + lane;
    if (j < K.n_pts) {  // fminf/fmaxf ignore the NaN pads anyway; this also skips slots past n_pad
        const long long s = K.pt_off + j;
        lo.x = hi.x = px[s]; lo.y = hi.y = py[s]; lo.z = hi.z = pz[s];
    }
    warp_box(lo, hi);
    if (lane == 0) { node_lo[K.node_off + leaf] = lo; node_hi[K.node_off + leaf] = hi; }
}

// one warp per inner node; level 1: children = leaves, level 2: children = level-1 nodes
__global__ void k_inner_aabb(const DevKf *__restrict__ kf, int kf_begin, int level, float4 *__restrict__ node_lo,
                             float4 *__restrict__ node_hi) {
    const DevKf K = kf[kf_begin + blockIdx.y];
    const int lane = threadIdx.x & 31;
    const int node = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int n_here = level == 1 ? K.n1 : 32;  // level-2 slots are always padded to 32
    if (node >= n_here) return;
    const long long child_base = K.node_off + (level == 1 ? 0 : K.n0);
    const int n_child = level == 1 ? K.n0 : K.n1;
    const long long out_base = K.node_off + (level == 1 ? K.n0 : K.n0 + K.n1);
    float4 lo = make_float4(INFINITY, INFINITY, INFINITY, 0.f), hi = make_float4(-INFINITY, -INFINITY, -INFINITY, 0.f);
    const int c = node * 32 + lane;
    if (c < n_child) { lo = node_lo[child_base + c]; hi = node_hi[child_base + c]; }
    warp_box(lo, hi);
    if (lane == 0) { node_lo[out_base + node] = lo; node_hi[out_base + node] = hi; }
}

// ---- leaf adjacency ---------------------------------------------------------------------------
// For every leaf box A: the (at most 32) nearest leaves whose box lies within r of A (conservative
// float32 lower bound of the box-to-box distance), nearest first — one per lane of the warp that later
// scans them — plus the COVERAGE of the row: the squared box distance of the nearest leaf that did not
// fit (+inf when every leaf within r is listed).  A search around a point of A whose answer lies closer
// than the coverage only has to look at the listed leaves, so the 3-level descent is replaced by one box
// test per lane.  Rows of empty leaves (or with > 256 leaves in range) stay empty: coverage -1.
constexpr int kAdjCand = 256;

__device__ __forceinline__ float boxbox_lb(float4 alo, float4 ahi, float4 blo, float4 bhi) {
    const float dx = fmaxf(fmaxf(__fsub_rd(blo.x, ahi.x), __fsub_rd(alo.x, bhi.x)), 0.f);
    const float dy = fmaxf(fmaxf(__fsub_rd(blo.y, ahi.y), __fsub_rd(alo.y, bhi.y)), 0.f);
    const float dz = fmaxf(fmaxf(__fsub_rd(blo.z, ahi.z), __fsub_rd(alo.z, bhi.z)), 0.f);
    const float s = __fadd_rd(__fadd_rd(__fmul_rd(dx, dx), __fmul_rd(dy, dy)), __fmul_rd(dz, dz));
    return __fmul_rd(s, 0.99999976f);
}

__device__ __forceinline__ unsigned long long warp_min64(unsigned long long v) {
    for (int o = 16; o; o >>= 1) {
        const unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w < v ? w : v;
    }
    return v;
}

__global__ void __launch_bounds__(256)
k_leaf_adj(const DevKf *__restrict__ kf, int kf_begin, const float4 *__restrict__ node_lo, const float4 *__restrict__ node_hi,
           uint16_t *__restrict__ adj, float *__restrict__ adj_cov, const float r2, const float min_cov) {
    __shared__ unsigned long long cand[8][kAdjCand];
    constexpr unsigned long long kNone = 0xffffffffffffffffull;
    const DevKf K = kf[kf_begin + blockIdx.y];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int leaf = blockIdx.x * 8 + warp;
    if (leaf >= K.n0) return;
    const float4 *lo0 = node_lo + K.node_off, *hi0 = node_hi + K.node_off;
    const float4 *lo1 = lo0 + K.n0, *hi1 = hi0 + K.n0, *lo2 = lo1 + K.n1, *hi2 = hi1 + K.n1;
    const float4 alo = lo0[leaf], ahi = hi0[leaf];
    int cnt = 0;
    bool over = !(alo.x <= ahi.x) || K.n0 > 65535;
    if (!over) {
        unsigned m2 = __ballot_sync(0xffffffffu, boxbox_lb(alo, ahi, lo2[lane], hi2[lane]) <= r2);
        while (m2 && !over) {
            const int s2 = __ffs(m2) - 1;
            m2 &= m2 - 1;
            unsigned m1 = __ballot_sync(0xffffffffu, boxbox_lb(alo, ahi, lo1[s2 * 32 + lane], hi1[s2 * 32 + lane]) <= r2);
            while (m1 && !over) {
                const int s1 = __ffs(m1) - 1;
                m1 &= m1 - 1;
                const int n0i = (s2 * 32 + s1) * 32 + lane;
                const float lb = boxbox_lb(alo, ahi, lo0[n0i], hi0[n0i]);
                const unsigned m0 = __ballot_sync(0xffffffffu, lb <= r2);
                const int c = __popc(m0);
                if (cnt + c > kAdjCand) { over = true; break; }
                if ((m0 >> lane) & 1u) cand[warp][cnt + __popc(m0 & ((1u << lane) - 1))] = ((unsigned long long)__float_as_uint(lb) << 32) | (unsigned)n0i;
                cnt += c;
            }
        }
    }
    __syncwarp();
    // the 32 smallest (distance, leaf) keys by repeated extraction; lane t keeps the t-th
    unsigned long long mine = kNone;
    float cov = -1.f;
    if (!over) {
        for (int t = 0; t <= 32; ++t) {
            unsigned long long best = kNone;
            for (int i = lane; i < cnt; i += 32) best = cand[warp][i] < best ? cand[warp][i] : best;
            best = warp_min64(best);
            if (t == 32) {  // the nearest leaf left out bounds what the row covers
                cov = best == kNone ? __int_as_float(0x7f800000) : __uint_as_float((unsigned)(best >> 32));
                break;
            }
            if (best == kNone) { cov = __int_as_float(0x7f800000); break; }
            if (lane == t) mine = best;
            for (int i = lane; i < cnt; i += 32)
                if (cand[warp][i] == best) cand[warp][i] = kNone;
            __syncwarp();
        }
    }
    // a truncated row that covers less than a third of the radius would mostly be scanned in vain
    if (cov >= 0.f && cov < min_cov) { cov = -1.f; mine = kNone; }
    adj[(K.node_off + leaf) * 32 + lane] = mine == kNone ? (uint16_t)0xffffu : (uint16_t)(mine & 0xffffu);
    if (lane == 0) adj_cov[K.node_off + leaf] = cov;
}

}  // namespace

cudaError_t build_scan_index(const float *d_raw, const long long *h_raw_off, int nkf, int kf_begin, const DevKf *h_kf, DevPack &pack,
                             float adj_r2, cudaStream_t st, BuildScratch &scr, float *kernel_ms) {
    const long long n = h_raw_off[nkf] - h_raw_off[0];
    cudaError_t err = cudaSuccess;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;  // device time of the build kernels alone (no allocation, no H2D copy)
    long long *d_off = nullptr;
    float *d_bbox = nullptr;
    unsigned long long *d_keys = nullptr, *d_keys2 = nullptr;
    uint32_t *d_vals = nullptr, *d_vals2 = nullptr;
    void *d_tmp = nullptr;
    size_t tmp_bytes = 0;
    int max_pad = 0, max_n0 = 0, max_n1 = 0;
    for (int f = 0; f < nkf; ++f) {
        max_pad = max_pad > h_kf[kf_begin + f].n_pad ? max_pad : h_kf[kf_begin + f].n_pad;
        max_n0 = max_n0 > h_kf[kf_begin + f].n0 ? max_n0 : h_kf[kf_begin + f].n0;
        max_n1 = max_n1 > h_kf[kf_begin + f].n1 ? max_n1 : h_kf[kf_begin + f].n1;
    }
    // chunk-relative offsets
    long long *h_rel = (long long *)malloc(sizeof(long long) * (nkf + 1));
    for (int f = 0; f <= nkf; ++f) h_rel[f] = h_raw_off[f] - h_raw_off[0];
#define STL_TRY(x) do { err = (x); if (err != cudaSuccess) goto done; } while (0)
    STL_TRY(scr.need(0, sizeof(long long) * (nkf + 1))); d_off = (long long *)scr.buf[0];
    STL_TRY(cudaMemcpyAsync(d_off, h_rel, sizeof(long long) * (nkf + 1), cudaMemcpyHostToDevice, st));
    STL_TRY(scr.need(1, sizeof(float) * 6 * nkf)); d_bbox = (float *)scr.buf[1];
    if (n > 0) {
        STL_TRY(scr.need(2, sizeof(unsigned long long) * n)); d_keys = (unsigned long long *)scr.buf[2];
        STL_TRY(scr.need(3, sizeof(unsigned long long) * n)); d_keys2 = (unsigned long long *)scr.buf[3];
        STL_TRY(scr.need(4, sizeof(uint32_t) * n)); d_vals = (uint32_t *)scr.buf[4];
        STL_TRY(scr.need(5, sizeof(uint32_t) * n)); d_vals2 = (uint32_t *)scr.buf[5];
    }
    if (n > 0) {  // the sort scratch is sized before the timed part
        int kf_bits0 = 1;
        while ((1 << kf_bits0) < nkf) ++kf_bits0;
        STL_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys, d_keys2, d_vals, d_vals2, n, 0, 48 + kf_bits0, st));
        STL_TRY(scr.need(6, tmp_bytes)); d_tmp = scr.buf[6];
    }
    if (kernel_ms) {
        STL_TRY(cudaEventCreate(&ev0)); STL_TRY(cudaEventCreate(&ev1));
        STL_TRY(cudaEventRecord(ev0, st));
    }
    k_bbox<<<nkf, 512, 0, st>>>(d_raw, d_off, d_bbox, pack.kf, kf_begin);
    if (n > 0) {
        const int tiles = (max_pad + 256 * 8 - 1) / (256 * 8);
        k_morton<<<dim3(tiles, nkf), 256, 0, st>>>(d_raw, d_off, d_bbox, d_keys, d_vals);
        int kf_bits = 1;
        while ((1 << kf_bits) < nkf) ++kf_bits;
        STL_TRY(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_keys, d_keys2, d_vals, d_vals2, n, 0, 48 + kf_bits, st));
    }
    {
        const int tiles = (max_pad + 256 * 8 - 1) / (256 * 8);
        k_scatter<<<dim3(tiles > 0 ? tiles : 1, nkf), 256, 0, st>>>(d_raw, d_off, d_vals2, pack.kf, kf_begin, pack.px, pack.py, pack.pz,
                                                                    pack.orig);
        if (max_pad > 0) {
            static bool configured = false;
            if (!configured) {
                STL_TRY(cudaFuncSetAttribute(k_kd_refine, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RefineSmem)));
                configured = true;
            }
            k_kd_refine<<<dim3((max_pad + kGroup - 1) / kGroup, nkf), kRefThreads, sizeof(RefineSmem), st>>>(pack.kf, kf_begin, pack.px, pack.py,
                                                                                                          pack.pz, pack.orig);
        }
        k_leaf_aabb<<<dim3((max_n0 + 7) / 8, nkf), 256, 0, st>>>(pack.kf, kf_begin, pack.px, pack.py, pack.pz, pack.node_lo, pack.node_hi);
        k_inner_aabb<<<dim3((max_n1 + 7) / 8, nkf), 256, 0, st>>>(pack.kf, kf_begin, 1, pack.node_lo, pack.node_hi);
        k_inner_aabb<<<dim3(4, nkf), 256, 0, st>>>(pack.kf, kf_begin, 2, pack.node_lo, pack.node_hi);
        if (pack.adj && adj_r2 > 0.f)
            k_leaf_adj<<<dim3((max_n0 + 7) / 8, nkf), 256, 0, st>>>(pack.kf, kf_begin, pack.node_lo, pack.node_hi, pack.adj, pack.adj_cov, adj_r2,
                                                                    adj_r2 * (getenv("STL_ADJ_TRUNC") ? (float)atof(getenv("STL_ADJ_TRUNC")) : 1.f / 9.f));
    }
    STL_TRY(cudaGetLastError());
    if (kernel_ms) STL_TRY(cudaEventRecord(ev1, st));
    STL_TRY(cudaStreamSynchronize(st));
    if (kernel_ms) {
        float ms = 0.f;
        STL_TRY(cudaEventElapsedTime(&ms, ev0, ev1));
        *kernel_ms += ms;
    }
#undef STL_TRY
done:
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    free(h_rel);
    return err;
}

}  // namespace stl
