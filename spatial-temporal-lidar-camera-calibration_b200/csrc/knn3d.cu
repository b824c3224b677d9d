// knn3d.cu — K2: 3-D/3-D alignment term, one warp per (candidate, keyframe, map point).
//
// Replaces, per 2-D correspondence that carries a map point:
//   map point -> LiDAR frame        src/examples/iba_global.cpp:231-234
//   ComputeAlignmentDist            src/examples/iba_global.cpp:111-156
//     (1-NN :116-121, k-NN of the neighbour :125-129, radius truncation :130-133,
//      gates :136-139, ComputeCovariance + FastEigen3x3_EV :140-143, regression
//      gate and point-to-plane / point-to-point distance :144-154)
//   thresholded accumulation        src/examples/iba_global.cpp:239-251
// plus the stand-alone k-NN entry used by the parity tests.
#include "kernels.h"
#include "knn.cuh"

namespace stl {
namespace {

constexpr int kWarps = 8;

__global__ void __launch_bounds__(kWarps * 32, 3)
k_align3d(const DevPack pk, const DevWork wk, const DevParams pr, const int B, const int debug) {
    const int sub = wk.sub;
    const int j = blockIdx.x % sub;
    const int bf = blockIdx.x / sub;
    const int f = bf / B, b = bf - f * B;
    const long long rec = (long long)b * pk.n_kf + f;
    const int nq = wk.n_q[rec];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ double red[5][kWarps];
    __shared__ double plane_sm[kWarps][kPlaneSmemDoubles];
    double s3d = 0, v3d = 0, c3d = 0, vpl = 0, vpt = 0;
    if (nq > 0) {
        const DevKf K = pk.kf[f];
        const DevCand &c = wk.cand[b];
        const ScanView S = make_view(pk, K);
        const long long base = (long long)b * pk.n_kp_total + K.kp_off;
        const float *Tcw = pk.Tcw + (long long)f * 12;
        double Rcw[9], tcw[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
#pragma unroll
            for (int a = 0; a < 3; ++a) Rcw[i * 3 + a] = (double)Tcw[i * 4 + a];
            tcw[i] = dmul((double)Tcw[i * 4 + 3], c.s);  // TcwRS.topRightCorner *= scale (iba_global.cpp:208)
        }
        for (int qi = j * kWarps + warp; qi < nq; qi += sub * kWarps) {
            const uint32_t ci = wk.q_corr[base + qi];
            const uint32_t kp = wk.corr_kp[base + ci];
            const float *mp = pk.kp_mp + (K.kp_off + kp) * 3;
            // GetWorldPos()*scale evaluated in float32, then widened (iba_global.cpp:232)
            const double wx = (double)__fmul_rn(mp[0], c.sf), wy = (double)__fmul_rn(mp[1], c.sf), wz = (double)__fmul_rn(mp[2], c.sf);
            const double cxm = dadd(dot3e(Rcw[0], Rcw[1], Rcw[2], wx, wy, wz), tcw[0]);
            const double cym = dadd(dot3e(Rcw[3], Rcw[4], Rcw[5], wx, wy, wz), tcw[1]);
            const double czm = dadd(dot3e(Rcw[6], Rcw[7], Rcw[8], wx, wy, wz), tcw[2]);
            double qx, qy, qz;
            xform(c.Ri, c.ti, cxm, cym, czm, qx, qy, qz);  // Tcl.inverse() * P

            Sink1 nn;
            traverse(S, qx, qy, qz, nn, lane);
            const double nx = (double)S.px[nn.pos], ny = (double)S.py[nn.pos], nz = (double)S.pz[nn.pos];
            const double dx = dsub(nx, qx), dy = dsub(ny, qy), dz = dsub(nz, qz);
            double dist = sqrt(dot3e(dx, dy, dz, dx, dy, dz));  // pt2pt
            int is_plane = 0, m = 0;
            int stat_k[3] = {0, 0, 0};
            if (pr.use_plane) {
                SinkK kn(pr.k, pr.radius2);
                traverse(S, nx, ny, nz, kn, lane);
                stat_k[0] = kn.n_iter; stat_k[1] = kn.n_visit; stat_k[2] = kn.n_ins;
                const PlaneOut po = plane_from_knn(S, kn, nx, ny, nz, pr, lane, plane_sm[warp]);
                m = po.m;
                if (po.gates_ok && !(po.reg > pr.reg_thr)) {
                    is_plane = 1;
                    dist = fabs(dot3e(dx, dy, dz, po.n.x, po.n.y, po.n.z));
                }
                if (debug) {
                    if (lane < kMaxK) wk.dbg_knn[(K.kp_off + qi) * kMaxK + lane] = lane < m ? kn.ki : 0xffffffffu;
                }
            }
            if (dist < pr.thr3d) {
                s3d += dist; v3d += 1.0;
                if (is_plane) vpl += 1.0; else vpt += 1.0;
            }
            c3d += 1.0;
            if (debug && lane == 0) {
                atomicAdd(&wk.dbg_stats[0], 1ull);
                atomicAdd(&wk.dbg_stats[1], (unsigned long long)nn.n_iter);
                atomicAdd(&wk.dbg_stats[2], (unsigned long long)nn.n_visit);
                atomicAdd(&wk.dbg_stats[3], (unsigned long long)stat_k[0]);
                atomicAdd(&wk.dbg_stats[4], (unsigned long long)stat_k[1]);
                atomicAdd(&wk.dbg_stats[5], (unsigned long long)stat_k[2]);
                atomicAdd(&wk.dbg_stats[6], (unsigned long long)m);
                wk.dbg_nn[K.kp_off + qi] = nn.oi;
                wk.dbg_m[K.kp_off + qi] = m;
                wk.dbg_plane[K.kp_off + qi] = is_plane;
                wk.dbg_dist[K.kp_off + qi] = dist;
            }
        }
    }
    if (lane == 0) { red[0][warp] = s3d; red[1][warp] = v3d; red[2][warp] = c3d; red[3][warp] = vpl; red[4][warp] = vpt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        AlignRec r = {0, 0, 0, 0, 0};
        for (int w = 0; w < kWarps; ++w) { r.s3d += red[0][w]; r.v3d += red[1][w]; r.c3d += red[2][w]; r.vpl += red[3][w]; r.vpt += red[4][w]; }
        wk.align[rec * sub + j] = r;
    }
}

// stand-alone exact k-NN: one warp per query
__global__ void __launch_bounds__(kWarps * 32)
k_knn3d(const DevPack pk, const int kf, const double *__restrict__ q, const int nq, const int k, const double radius2,
        uint32_t *__restrict__ out_idx, double *__restrict__ out_d2, int *__restrict__ out_cnt) {
    const int lane = threadIdx.x & 31;
    const int qi = blockIdx.x * kWarps + (threadIdx.x >> 5);
    if (qi >= nq) return;
    const DevKf K = pk.kf[kf];
    const ScanView S = make_view(pk, K);
    SinkK kn(k, radius2 > 0 ? radius2 : DBL_MAX);
    traverse(S, q[qi * 3], q[qi * 3 + 1], q[qi * 3 + 2], kn, lane);
    if (lane < k) {
        out_idx[(long long)qi * k + lane] = lane < kn.count ? kn.ki : 0xffffffffu;
        out_d2[(long long)qi * k + lane] = lane < kn.count ? kn.kd : INFINITY;
    }
    if (lane == 0) out_cnt[qi] = kn.count;
}

}  // namespace

cudaError_t launch_align3d(const DevPack &pk, const DevWork &wk, const DevParams &pr, int B, int debug, cudaStream_t st) {
    if (B <= 0 || pk.n_kf <= 0) return cudaSuccess;
    k_align3d<<<(unsigned)(pk.n_kf * B * wk.sub), kWarps * 32, 0, st>>>(pk, wk, pr, B, debug);
    return cudaGetLastError();
}

cudaError_t launch_knn3d(const DevPack &pk, int kf, const double *d_q, int nq, int k, double radius2, uint32_t *d_idx, double *d_d2,
                         int *d_cnt, cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    k_knn3d<<<(nq + kWarps - 1) / kWarps, kWarps * 32, 0, st>>>(pk, kf, d_q, nq, k, radius2, d_idx, d_d2, d_cnt);
    return cudaGetLastError();
}

}  // namespace stl
