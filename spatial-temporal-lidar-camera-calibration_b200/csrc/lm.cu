// lm.cu — the Levenberg-Marquardt side of the boundary.
//
// K4a  k_lm_associate: BuildProblem (src/examples/iba_local.cpp:145-323) — for every 2-D
//      correspondence of the association pass (K1 at the current estimate):
//        ComputeLocalNeighbor            include/pointcloud.h:733-760   (iba_local.cpp:207)
//        plane fit + regression gate     iba_local.cpp:218-231
//        covisible observations          iba_local.cpp:245-259
//        map point -> LiDAR, 1-NN gate   iba_local.cpp:239-240,283-290
//        ComputeLocalNormalSingleThre    include/pointcloud.h:699-717,651-666 (iba_local.cpp:295)
//      and freezes the residual blocks on the device.
// K4b  k_linearize: evaluates the frozen blocks at B parameter vectors —
//        IBA_PlaneFactor::operator()     include/IBACalib2.hpp:152-184 (g2o twin IBACalib.hpp:103-140)
//        Point2Point_Factor              include/IBACalib2.hpp:570-584
//        Point2Plane_Factor              include/IBACalib2.hpp:611-625
//      with forward-mode duals (dual.cuh), ceres::HuberLoss + Corrector semantics
//      (iba_local.cpp:263,291; rho'' <= 0 => residual and Jacobian scaled by sqrt(rho')),
//      and reduces cost, J^T r and the 7x7 J^T J in fp64 (warp shuffles, fixed order).
#include <cub/cub.cuh>

#include "../../include/stlcalib.h"
#include "dual.cuh"
#include "knn.cuh"
#include "lm.h"

namespace stl {
namespace {

constexpr int kWarps = 8;
constexpr int kAssocSub = 8;
constexpr int kLinVals = 40;  // cost, g[7], H upper 28, n2d, npt, npl, nres
constexpr int kLinThreads = 128;

// ------------------------------------------------------------------ K4a
__global__ void __launch_bounds__(kWarps * 32)
k_lm_associate(const DevPack pk, const DevWork wk, const DevParams pr, LmState lm) {
    const int j = blockIdx.x % kAssocSub, f = blockIdx.x / kAssocSub;
    const int nc = wk.n_corr[f];
    if (nc < pr.num_min_corr) return;  // iba_local.cpp:192
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ double plane_sm[kWarps][kPlaneSmemDoubles];
    const DevKf K = pk.kf[f];
    const DevCand &c0 = wk.cand[0];
    const ScanView S = make_view(pk, K);
    const float *Tcw = pk.Tcw + (long long)f * 12;
    const int C = pk.n_covis;
    for (int i = j * kWarps + warp; i < nc; i += kAssocSub * kWarps) {
        const long long slot = K.kp_off + i;
        const uint32_t kp = wk.corr_kp[slot], sp = wk.corr_sp[slot];
        const double cx = (double)S.px[sp], cy = (double)S.py[sp], cz = (double)S.pz[sp];
        // The reference runs ComputeLocalNeighbor for every correspondence and only then drops the
        // ones without a map point (iba_local.cpp:207-213) or without a covisible observation
        // (:259); both tests are independent of the neighbour search, so they come first here.
        const float *mp = pk.kp_mp + (K.kp_off + kp) * 3;
        if (isnan(mp[0])) continue;  // iba_local.cpp:213
        int ncov = 0;
        for (int s = 0; s < C; ++s)
            if (pk.covis_valid[f * C + s] && !isnan(pk.covis_uv[(K.kp_off + kp) * C + s].x)) ++ncov;
        if (ncov == 0) continue;  // iba_local.cpp:259
        // ComputeLocalNeighbor around the scan point
        SinkK kn(pr.k, pr.radius2);
        traverse(S, cx, cy, cz, kn, lane);
        const PlaneOut po = plane_from_knn(S, kn, cx, cy, cz, pr, lane, plane_sm[warp]);
        if (!po.gates_ok) continue;  // m < min_pts || d2[m-1] < min_diff^2
        const bool valid_plane = po.reg < pr.reg_thr;  // strict '<' (iba_local.cpp:231)
        if (lane == 0) {
            lm.slot_kf[slot] = f;
            lm.slot_kp[slot] = kp;
            if (valid_plane) {
                double *g = lm.geo2d + slot * 6;
                g[0] = cx; g[1] = cy; g[2] = cz; g[3] = po.n.x; g[4] = po.n.y; g[5] = po.n.z;
                lm.flag2d[slot] = 1;
            }
        }
        // MapPoint = Tcw * Pw in fp64, no scale (iba_local.cpp:239-240)
        const double a = (double)mp[0], b = (double)mp[1], cc = (double)mp[2];
        const double Mx = dadd(dot3e((double)Tcw[0], (double)Tcw[1], (double)Tcw[2], a, b, cc), (double)Tcw[3]);
        const double My = dadd(dot3e((double)Tcw[4], (double)Tcw[5], (double)Tcw[6], a, b, cc), (double)Tcw[7]);
        const double Mz = dadd(dot3e((double)Tcw[8], (double)Tcw[9], (double)Tcw[10], a, b, cc), (double)Tcw[11]);
        double qx, qy, qz;
        xform(c0.Ri, c0.ti, dmul(Mx, c0.s), dmul(My, c0.s), dmul(Mz, c0.s), qx, qy, qz);  // initSE3.inverse() * (MapPoint * init_scale)
        Sink1 nn;
        traverse(S, qx, qy, qz, nn, lane);
        if (nn.d > pr.max_3d_dist2) continue;  // iba_local.cpp:289
        const double nx = (double)S.px[nn.pos], ny = (double)S.py[nn.pos], nz = (double)S.pz[nn.pos];
        PlaneOut p2 = po;  // the map point's neighbour is very often the associated scan point itself
        if (nn.pos != sp) {
            SinkK kn2(pr.k, pr.radius2);
            traverse(S, nx, ny, nz, kn2, lane);
            p2 = plane_from_knn(S, kn2, nx, ny, nz, pr, lane, plane_sm[warp]);
        }
        const bool state = p2.gates_ok && p2.reg < pr.reg_thr;
        if (lane == 0) {
            double *g = lm.geo3d + slot * 9;
            g[0] = Mx; g[1] = My; g[2] = Mz; g[3] = nx; g[4] = ny; g[5] = nz;
            g[6] = p2.gates_ok ? p2.n.x : 0.0; g[7] = p2.gates_ok ? p2.n.y : 0.0; g[8] = p2.gates_ok ? p2.n.z : 1.0;
            lm.type3d[slot] = state ? 2 : 1;  // Point2Plane_Factor : Point2Point_Factor (iba_local.cpp:300-309)
            lm.flag3d[slot] = 1;
        }
    }
}

__global__ void k_count_types(const uint8_t *__restrict__ type3d, const int *__restrict__ idx3d, int n3d, int *out) {
    int pt = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n3d; i += gridDim.x * blockDim.x) pt += type3d[idx3d[i]] == 1;
    for (int o = 16; o; o >>= 1) pt += __shfl_down_sync(0xffffffffu, pt, o);
    if ((threadIdx.x & 31) == 0 && pt) atomicAdd(out, pt);
}

// ------------------------------------------------------------------ K4b
struct Acc {
    double v[kLinVals];
};

__device__ __forceinline__ void huber(double sq, double delta, double &rho0, double &sr) {
    if (sq > delta * delta) {  // ceres::HuberLoss::Evaluate
        const double r = sqrt(sq);
        rho0 = 2.0 * delta * r - delta * delta;
        sr = sqrt(fmax(DBL_MIN, delta / r));
    } else {
        rho0 = sq;
        sr = 1.0;
    }
}

__device__ __forceinline__ void accumulate(Acc &A, const D7 &e, double sr) {
    const double r = sr * e.a;
    double J[7];
#pragma unroll
    for (int a = 0; a < 7; ++a) J[a] = sr * e.v[a];
    int h = 8;
#pragma unroll
    for (int a = 0; a < 7; ++a) {
        A.v[1 + a] += J[a] * r;
#pragma unroll
        for (int b = a; b < 7; ++b) A.v[h++] += J[a] * J[b];
    }
}

__device__ __forceinline__ void mv3(const D7 *M, const D7 *p, D7 *o) {
#pragma unroll
    for (int i = 0; i < 3; ++i) o[i] = (M[i * 3] * p[0] + M[i * 3 + 1] * p[1]) + M[i * 3 + 2] * p[2];
}

// grid (chunks, B)
__global__ void __launch_bounds__(kLinThreads)
k_linearize(const DevPack pk, const DevParams pr, const LmState lm, const LmCand *__restrict__ cands, double *__restrict__ partial) {
    __shared__ LmCand c;
    {
        const double *src = reinterpret_cast<const double *>(cands + blockIdx.y);
        double *dst = reinterpret_cast<double *>(&c);
        for (int i = threadIdx.x; i < (int)(sizeof(LmCand) / 8); i += kLinThreads) dst[i] = src[i];
    }
    __syncthreads();
    Acc A;
#pragma unroll
    for (int i = 0; i < kLinVals; ++i) A.v[i] = 0.0;
    const int C = pk.n_covis;
    const int stride = gridDim.x * kLinThreads, t0 = blockIdx.x * kLinThreads + threadIdx.x;

    // ---- 3-D/2-D blocks: IBA_PlaneFactor (IBACalib2.hpp:152-184)
    for (int it = t0; it < lm.n2d; it += stride) {
        const int slot = lm.idx2d[it];
        const int f = lm.slot_kf[slot];
        const uint32_t kp = lm.slot_kp[slot];
        const DevKf &K = pk.kf[f];
        const double fx = K.fx, fy = K.fy, cx = K.cx, cy = K.cy;
        const float2 kxy = pk.kp_xy[K.kp_off + kp];
        const double *g = lm.geo2d + (long long)slot * 6;
        const D7 p0[3] = {d7_const(g[0]), d7_const(g[1]), d7_const(g[2])}, n0[3] = {d7_const(g[3]), d7_const(g[4]), d7_const(g[5])};
        D7 p0c[3], n0c[3];
        mv3(c.R, p0, p0c);
#pragma unroll
        for (int i = 0; i < 3; ++i) p0c[i] = p0c[i] + c.t[i];
        mv3(c.R, n0, n0c);
        const double Cxz = ((double)kxy.x - cx) / fx, Cyz = ((double)kxy.y - cy) / fy;
        const D7 num = (n0c[0] * p0c[0] + n0c[1] * p0c[1]) + n0c[2] * p0c[2];
        const D7 den = (n0c[0] * Cxz + n0c[1] * Cyz) + n0c[2];
        const D7 Z0 = num / den;
        const D7 P0[3] = {Z0 * Cxz, Z0 * Cyz, Z0};
        // pass 1: squared norm of the block (values only) for the robust kernel
        double sq = 0.0;
        int nres = 0;
        for (int s = 0; s < C; ++s) {
            if (!pk.covis_valid[f * C + s]) continue;
            const float2 uv = pk.covis_uv[(K.kp_off + kp) * C + s];
            if (isnan(uv.x)) continue;
            const float *rp = pk.relpose + ((long long)f * C + s) * 12;
            double P1[3];
#pragma unroll
            for (int i = 0; i < 3; ++i)
                P1[i] = (((double)rp[i * 4] * P0[0].a + (double)rp[i * 4 + 1] * P0[1].a) + (double)rp[i * 4 + 2] * P0[2].a) + (double)rp[i * 4 + 3] * c.s.a;
            const double eu = (fx * P1[0] / P1[2] + cx) - (double)uv.x, ev = (fy * P1[1] / P1[2] + cy) - (double)uv.y;
            sq += eu * eu;
            sq += ev * ev;
            nres += 2;
        }
        double rho0, sr;
        huber(sq, pr.delta2d, rho0, sr);
        A.v[0] += 0.5 * rho0;
        A.v[36] += 1.0;
        A.v[39] += (double)nres;
        // pass 2: duals
        for (int s = 0; s < C; ++s) {
            if (!pk.covis_valid[f * C + s]) continue;
            const float2 uv = pk.covis_uv[(K.kp_off + kp) * C + s];
            if (isnan(uv.x)) continue;
            const float *rp = pk.relpose + ((long long)f * C + s) * 12;
            D7 P1[3];
#pragma unroll
            for (int i = 0; i < 3; ++i)
                P1[i] = ((P0[0] * (double)rp[i * 4] + P0[1] * (double)rp[i * 4 + 1]) + P0[2] * (double)rp[i * 4 + 2]) + c.s * (double)rp[i * 4 + 3];
            const D7 eu = ((P1[0] * fx) / P1[2] + cx) - (double)uv.x;
            const D7 ev = ((P1[1] * fy) / P1[2] + cy) - (double)uv.y;
            accumulate(A, eu, sr);
            accumulate(A, ev, sr);
        }
    }

    // ---- 3-D/3-D blocks: Point2Point_Factor / Point2Plane_Factor (IBACalib2.hpp:570-584,611-625)
    for (int it = t0; it < lm.n3d; it += stride) {
        const int slot = lm.idx3d[it];
        const double *g = lm.geo3d + (long long)slot * 9;
        const D7 Ms[3] = {c.s * g[0], c.s * g[1], c.s * g[2]};  // MapPoint * s
        D7 M[3];
        mv3(c.Rlc, Ms, M);
#pragma unroll
        for (int i = 0; i < 3; ++i) M[i] = M[i] + c.tlc[i];
        const D7 d[3] = {M[0] - g[3], M[1] - g[4], M[2] - g[5]};
        double rho0, sr;
        if (lm.type3d[slot] == 1) {
            huber((d[0].a * d[0].a + d[1].a * d[1].a) + d[2].a * d[2].a, pr.delta3d, rho0, sr);
            accumulate(A, d[0], sr);
            accumulate(A, d[1], sr);
            accumulate(A, d[2], sr);
            A.v[37] += 1.0;
            A.v[39] += 3.0;
        } else {
            const D7 e = (d[0] * g[6] + d[1] * g[7]) + d[2] * g[8];
            huber(e.a * e.a, pr.delta3d, rho0, sr);
            accumulate(A, e, sr);
            A.v[38] += 1.0;
            A.v[39] += 1.0;
        }
        A.v[0] += 0.5 * rho0;
    }

    // ---- CTA reduction (fixed order)
    __shared__ double red[kLinThreads / 32][kLinVals];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < kLinVals; ++i) {
        double x = A.v[i];
        for (int o = 16; o; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if (lane == 0) red[warp][i] = x;
    }
    __syncthreads();
    if (threadIdx.x < kLinVals) {
        double x = 0.0;
        for (int w = 0; w < kLinThreads / 32; ++w) x += red[w][threadIdx.x];
        partial[((long long)blockIdx.y * gridDim.x + blockIdx.x) * kLinVals + threadIdx.x] = x;
    }
}

// one CTA per candidate: sums the per-CTA partials in order and expands H to the full symmetric 7x7
__global__ void k_lin_finish(const double *__restrict__ partial, int nchunks, double *__restrict__ out) {
    __shared__ double tot[kLinVals];
    const int b = blockIdx.x;
    if (threadIdx.x < kLinVals) {
        double x = 0.0;
        for (int i = 0; i < nchunks; ++i) x += partial[((long long)b * nchunks + i) * kLinVals + threadIdx.x];
        tot[threadIdx.x] = x;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double *o = out + (long long)b * STL_LIN_NSUMS;
        o[0] = tot[0];
        for (int a = 0; a < 7; ++a) o[1 + a] = tot[1 + a];
        int h = 8;
        for (int a = 0; a < 7; ++a)
            for (int c = a; c < 7; ++c) { o[8 + a * 7 + c] = tot[h]; o[8 + c * 7 + a] = tot[h]; ++h; }
        o[57] = tot[36]; o[58] = tot[37]; o[59] = tot[38]; o[60] = tot[39];
    }
}

template <class T> void dfree(T *&p) { if (p) { cudaFree(p); p = nullptr; } }

}  // namespace

void lm_free(LmState &lm) {
    dfree(lm.slot_kf); dfree(lm.slot_kp); dfree(lm.flag2d); dfree(lm.type3d); dfree(lm.flag3d); dfree(lm.geo2d); dfree(lm.geo3d);
    dfree(lm.idx2d); dfree(lm.idx3d); dfree(lm.d_counts); dfree(lm.d_tmp); dfree(lm.partial); dfree(lm.d_cand);
    if (lm.h_cand) cudaFreeHost(lm.h_cand);
    if (lm.h2d_done) cudaEventDestroy(lm.h2d_done);
    lm = LmState();
}

cudaError_t lm_associate(const DevPack &pk, const DevWork &wk, const DevParams &pr, LmState &lm, cudaStream_t st) {
    cudaError_t e;
#define TRY(x) do { e = (x); if (e != cudaSuccess) return e; } while (0)
    const long long ns = pk.n_kp_total > 0 ? pk.n_kp_total : 1;
    if (lm.n_slots != ns) {
        lm_free(lm);
        lm.n_slots = ns;
        TRY(cudaMalloc(&lm.slot_kf, 4 * ns)); TRY(cudaMalloc(&lm.slot_kp, 4 * ns));
        TRY(cudaMalloc(&lm.flag2d, ns)); TRY(cudaMalloc(&lm.type3d, ns)); TRY(cudaMalloc(&lm.flag3d, ns));
        TRY(cudaMalloc(&lm.geo2d, 48 * ns)); TRY(cudaMalloc(&lm.geo3d, 72 * ns));
        TRY(cudaMalloc(&lm.idx2d, 4 * ns)); TRY(cudaMalloc(&lm.idx3d, 4 * ns));
        TRY(cudaMalloc(&lm.d_counts, 16));
        size_t tb = 0;
        cub::CountingInputIterator<int> it(0);
        TRY(cub::DeviceSelect::Flagged(nullptr, tb, it, lm.flag2d, lm.idx2d, lm.d_counts, (int)ns, st));
        lm.tmp_bytes = tb;
        TRY(cudaMalloc(&lm.d_tmp, tb));
    }
    lm.ready = false;
    TRY(cudaMemsetAsync(lm.flag2d, 0, ns, st));
    TRY(cudaMemsetAsync(lm.type3d, 0, ns, st));
    TRY(cudaMemsetAsync(lm.flag3d, 0, ns, st));
    TRY(cudaMemsetAsync(lm.d_counts, 0, 16, st));
    k_lm_associate<<<(unsigned)(pk.n_kf * kAssocSub), kWarps * 32, 0, st>>>(pk, wk, pr, lm);
    TRY(cudaGetLastError());
    cub::CountingInputIterator<int> it(0);
    size_t tb = lm.tmp_bytes;
    TRY(cub::DeviceSelect::Flagged(lm.d_tmp, tb, it, lm.flag2d, lm.idx2d, lm.d_counts, (int)ns, st));
    tb = lm.tmp_bytes;
    TRY(cub::DeviceSelect::Flagged(lm.d_tmp, tb, it, lm.flag3d, lm.idx3d, lm.d_counts + 1, (int)ns, st));
    int h[2] = {0, 0};
    TRY(cudaMemcpyAsync(h, lm.d_counts, 8, cudaMemcpyDeviceToHost, st));
    TRY(cudaStreamSynchronize(st));
    lm.n2d = h[0]; lm.n3d = h[1];
    int npt = 0;
    if (lm.n3d > 0) {
        k_count_types<<<64, 256, 0, st>>>(lm.type3d, lm.idx3d, lm.n3d, lm.d_counts + 2);
        TRY(cudaGetLastError());
        TRY(cudaMemcpyAsync(&npt, lm.d_counts + 2, 4, cudaMemcpyDeviceToHost, st));
        TRY(cudaStreamSynchronize(st));
    }
    lm.n_blocks[0] = lm.n2d; lm.n_blocks[1] = npt; lm.n_blocks[2] = lm.n3d - npt;
    lm.ready = true;
    return cudaSuccess;
}

cudaError_t lm_linearize(const DevPack &pk, const DevParams &pr, LmState &lm, const double *x, int B, double *d_out, cudaStream_t st) {
    cudaError_t e;
    if (B > lm.cand_cap) {
        dfree(lm.d_cand);
        if (lm.h_cand) cudaFreeHost(lm.h_cand);
        lm.h_cand = nullptr;
        TRY(cudaMalloc(&lm.d_cand, sizeof(LmCand) * B));
        TRY(cudaMallocHost(&lm.h_cand, sizeof(LmCand) * B));
        lm.cand_cap = B;
    }
    LmCand *hc = reinterpret_cast<LmCand *>(lm.h_cand);
    if (!lm.h2d_done) TRY(cudaEventCreateWithFlags(&lm.h2d_done, cudaEventDisableTiming));
    TRY(cudaEventSynchronize(lm.h2d_done));
    for (int b = 0; b < B; ++b) make_lm_candidate(x + (size_t)b * 7, hc + b);
    TRY(cudaMemcpyAsync(lm.d_cand, hc, sizeof(LmCand) * B, cudaMemcpyHostToDevice, st));
    TRY(cudaEventRecord(lm.h2d_done, st));
    const int work = lm.n2d > lm.n3d ? lm.n2d : lm.n3d;
    int chunks = (work + kLinThreads * 2 - 1) / (kLinThreads * 2);
    if (chunks < 1) chunks = 1;
    if (chunks > 148 * 8) chunks = 148 * 8;
    const long long need = (long long)B * chunks * kLinVals;
    if (need > lm.partial_cap) {
        dfree(lm.partial);
        TRY(cudaMalloc(&lm.partial, 8 * need));
        lm.partial_cap = need;
    }
    k_linearize<<<dim3(chunks, B), kLinThreads, 0, st>>>(pk, pr, lm, reinterpret_cast<const LmCand *>(lm.d_cand), lm.partial);
    TRY(cudaGetLastError());
    k_lin_finish<<<B, 64, 0, st>>>(lm.partial, chunks, d_out);
    TRY(cudaGetLastError());
#undef TRY
    return cudaSuccess;
}

}  // namespace stl
