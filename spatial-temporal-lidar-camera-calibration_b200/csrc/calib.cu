// calib.cu — the two problems that precede the cost evaluation on the same 7-parameter block:
//
//   k_he_linearize     EdgeHE + EdgeRegulation of the hand-eye initialisation
//                      include/NLHECalib.hpp:27-86 (error / Jacobian exactly as coded there: the rotation and the
//                      translation residual are ADDED into one 3-vector, the Jacobian is [hat(R r_b) | Ra - I | ta]),
//                      :88-116 (regularisation edge), g2o robust-kernel semantics (:141-143)
//   k_calib_linearize  calibEdge of Optimizer::OptimizeExtrinsicGlobal / Local
//                      src/orb_slam/src/Optimizer.cc:65-205 (three Rodrigues transforms on g2o's autodiff numbers,
//                      re-projection error), :1399-1744 (information = invSigma2 I, Huber delta = sqrt(5.991), levels)
//
// Both reduce, per candidate, sum rho(chi2), g = J^T rho' Omega e and H = J^T rho' Omega J (g2o's constructQuadraticForm:
// b = -g) in fp64 with a fixed summation order, and optionally return the chi2 of every edge (the reference re-classifies
// outliers from them, Optimizer.cc:1505-1530).  One thread per edge, duals with 7 partials for the autodiff edge.
#include <vector>

#include "../../include/stlcalib.h"
#include "calibmath.hpp"
#include "kernels.h"

namespace stl {
namespace {

constexpr int kVals = 38;  // cost, g[7], H upper 28, active edges, residuals
constexpr int kT = 128;

struct Acc38 { double v[kVals]; };

__device__ __forceinline__ void g2o_huber(double e2, double delta, double &rho0, double &rho1) {
    if (delta > 0.0 && e2 > delta * delta) {  // g2o::RobustKernelHuber::robustify
        const double sq = sqrt(e2);
        rho0 = 2.0 * sq * delta - delta * delta;
        rho1 = delta / sq;
    } else {
        rho0 = e2;
        rho1 = 1.0;
    }
}

// adds one residual row: weight w = rho' * information
__device__ __forceinline__ void add_row(Acc38 &A, double r, const double *J, double w) {
    int h = 8;
#pragma unroll
    for (int a = 0; a < 7; ++a) {
        A.v[1 + a] += (J[a] * w) * r;
#pragma unroll
        for (int b = a; b < 7; ++b) A.v[h++] += (J[a] * w) * J[b];
    }
}

__device__ __forceinline__ void cta_reduce_store(const Acc38 &A, double *__restrict__ partial) {
    __shared__ double red[kT / 32][kVals];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < kVals; ++i) {
        double x = A.v[i];
        for (int o = 16; o; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if (lane == 0) red[warp][i] = x;
    }
    __syncthreads();
    if (threadIdx.x < kVals) {
        double x = 0.0;
        for (int w = 0; w < kT / 32; ++w) x += red[w][threadIdx.x];
        partial[threadIdx.x] = x;
    }
}

// grid (chunks, B)
__global__ void __launch_bounds__(kT)
k_he_linearize(const HeEdge *__restrict__ edges, int n, const DevCand *__restrict__ cands, const double *__restrict__ xs, double huber_delta,
               double regulation, double *__restrict__ partial, double *__restrict__ chi2_out) {
    const DevCand &c = cands[blockIdx.y];
    Acc38 A;
#pragma unroll
    for (int i = 0; i < kVals; ++i) A.v[i] = 0.0;
    for (int i = blockIdx.x * kT + threadIdx.x; i < n; i += gridDim.x * kT) {
        const HeEdge E = edges[i];
        // errRotVec = Rab * r_b - r_a;  errTran = (Ra - I) tab + ta s - Rab tb     (NLHECalib.hpp:41-48)
        double Rrb[3], Rtb[3], e[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            Rrb[r] = (c.R[r * 3] * E.rb[0] + c.R[r * 3 + 1] * E.rb[1]) + c.R[r * 3 + 2] * E.rb[2];
            Rtb[r] = (c.R[r * 3] * E.tb[0] + c.R[r * 3 + 1] * E.tb[1]) + c.R[r * 3 + 2] * E.tb[2];
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const double A0 = E.Ra[r * 3] - (r == 0 ? 1.0 : 0.0), A1 = E.Ra[r * 3 + 1] - (r == 1 ? 1.0 : 0.0), A2 = E.Ra[r * 3 + 2] - (r == 2 ? 1.0 : 0.0);
            const double tran = (((A0 * c.t[0] + A1 * c.t[1]) + A2 * c.t[2]) + E.ta[r] * c.s) - Rtb[r];  // A * [tab; s] - b
            e[r] = (Rrb[r] - E.ra[r]) + tran;                                                                // weight 1 (edge ctor, :136)
        }
        const double e2 = E.info * ((e[0] * e[0] + e[1] * e[1]) + e[2] * e[2]);
        double rho0, rho1;
        g2o_huber(e2, huber_delta, rho0, rho1);
        if (chi2_out) chi2_out[(long long)blockIdx.y * n + i] = e2;
        A.v[0] += rho0;
        A.v[36] += 1.0;
        A.v[37] += 3.0;
        // _jacobianOplusXi = [ -skew(R r_b) | Ra - I | ta ] with Eigen::skew as defined at NLHECalib.hpp:20-24 (= -hat), i.e. hat(R r_b)
        const double J[3][7] = {
            {0.0, -Rrb[2], Rrb[1], E.Ra[0] - 1.0, E.Ra[1], E.Ra[2], E.ta[0]},
            {Rrb[2], 0.0, -Rrb[0], E.Ra[3], E.Ra[4] - 1.0, E.Ra[5], E.ta[1]},
            {-Rrb[1], Rrb[0], 0.0, E.Ra[6], E.Ra[7], E.Ra[8] - 1.0, E.ta[2]}};
#pragma unroll
        for (int r = 0; r < 3; ++r) add_row(A, e[r], J[r], rho1 * E.info);
    }
    if (regulation > 0.0 && blockIdx.x == 0 && threadIdx.x == 0) {  // EdgeRegulation (:88-116): error = params[3..5], information = regulation * I
        const double *x = xs + blockIdx.y * 7;
        for (int r = 0; r < 3; ++r) {
            double J[7] = {0, 0, 0, 0, 0, 0, 0};
            J[3 + r] = 1.0;
            add_row(A, x[3 + r], J, regulation);
        }
        A.v[0] += regulation * ((x[3] * x[3] + x[4] * x[4]) + x[5] * x[5]);
        A.v[37] += 3.0;
    }
    cta_reduce_store(A, partial + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * kVals);
}

// Rodrigues on a dual vector with a dual axis (Optimizer.cc:118-136,163-181): X cth + (a x X) sth + a (a . X) (1 - cth)
__device__ __forceinline__ void rodrigues_dd(const D7 *a, const D7 &cth, const D7 &sth, const D7 &omc, const D7 *X, D7 *o) {
    D7 aX[3];
    aX[0] = a[1] * X[2] - a[2] * X[1];
    aX[1] = a[2] * X[0] - a[0] * X[2];
    aX[2] = a[0] * X[1] - a[1] * X[0];
    const D7 d = (a[0] * X[0] + a[1] * X[1]) + a[2] * X[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) o[i] = (X[i] * cth + aX[i] * sth) + (a[i] * d) * omc;
}

// grid (chunks, B)
__global__ void __launch_bounds__(kT)
k_calib_linearize(const CalibKf *__restrict__ kfs, const int *__restrict__ edge_kf, const double *__restrict__ Xw, const double *__restrict__ obs,
                  const float *__restrict__ inv_sigma2, const unsigned char *__restrict__ level, long long n, const CalibCand *__restrict__ cands,
                  double huber_delta, double *__restrict__ partial, double *__restrict__ chi2_out) {
    __shared__ CalibCand c;
    {
        const double *src = reinterpret_cast<const double *>(cands + blockIdx.y);
        double *dst = reinterpret_cast<double *>(&c);
        for (int i = threadIdx.x; i < (int)(sizeof(CalibCand) / 8); i += kT) dst[i] = src[i];
    }
    __syncthreads();
    Acc38 A;
#pragma unroll
    for (int i = 0; i < kVals; ++i) A.v[i] = 0.0;
    for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < n; i += (long long)gridDim.x * kT) {
        const CalibKf K = kfs[edge_kf[i]];
        const D7 Xc0[3] = {c.scale * Xw[i * 3], c.scale * Xw[i * 3 + 1], c.scale * Xw[i * 3 + 2]};  // scale * Xw
        D7 Xl0[3], Xli[3], Xci[3];
        if (!c.small1) {
            rodrigues_dd(c.v, c.cth, c.sth, c.omc, Xc0, Xl0);
        } else {  // Xc0 + w x Xc0
            Xl0[0] = Xc0[0] + (c.w1[1] * Xc0[2] - c.w1[2] * Xc0[1]);
            Xl0[1] = Xc0[1] + (c.w1[2] * Xc0[0] - c.w1[0] * Xc0[2]);
            Xl0[2] = Xc0[2] + (c.w1[0] * Xc0[1] - c.w1[1] * Xc0[0]);
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) Xl0[a] = Xl0[a] + c.p[a];
        if (!K.small2) {  // constant axis: the same formula with plain doubles on the axis side
            D7 aX[3];
            aX[0] = Xl0[2] * K.a2[1] - Xl0[1] * K.a2[2];
            aX[1] = Xl0[0] * K.a2[2] - Xl0[2] * K.a2[0];
            aX[2] = Xl0[1] * K.a2[0] - Xl0[0] * K.a2[1];
            const D7 d = (Xl0[0] * K.a2[0] + Xl0[1] * K.a2[1]) + Xl0[2] * K.a2[2];
#pragma unroll
            for (int a = 0; a < 3; ++a) Xli[a] = (Xl0[a] * K.cth2 + aX[a] * K.sth2) + (d * K.a2[a]) * K.omc2;
        } else {
            Xli[0] = Xl0[0] + (Xl0[2] * K.w2[1] - Xl0[1] * K.w2[2]);
            Xli[1] = Xl0[1] + (Xl0[0] * K.w2[2] - Xl0[2] * K.w2[0]);
            Xli[2] = Xl0[2] + (Xl0[1] * K.w2[0] - Xl0[0] * K.w2[1]);
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) Xli[a] = Xli[a] + K.t2[a];
        if (!c.small3) {
            rodrigues_dd(c.a3, c.cth3, c.sth3, c.omc3, Xli, Xci);
        } else {
            Xci[0] = Xli[0] + (c.w3[1] * Xli[2] - c.w3[2] * Xli[1]);
            Xci[1] = Xli[1] + (c.w3[2] * Xli[0] - c.w3[0] * Xli[2]);
            Xci[2] = Xli[2] + (c.w3[0] * Xli[1] - c.w3[1] * Xli[0]);
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) Xci[a] = Xci[a] + c.t3[a];
        const D7 pre0 = (Xci[0] * K.fx) / Xci[2] + K.cx, pre1 = (Xci[1] * K.fy) / Xci[2] + K.cy;
        const D7 e0 = -(pre0 - obs[i * 2]), e1 = -(pre1 - obs[i * 2 + 1]);  // measurement - pre
        const double info = (double)inv_sigma2[i];
        const double e2 = info * (e0.a * e0.a + e1.a * e1.a);
        if (chi2_out) chi2_out[(long long)blockIdx.y * n + i] = e2;
        if (level && level[i]) continue;  // setLevel(1): left out of the optimisation, chi2 still reported (Optimizer.cc:1509-1512)
        double rho0, rho1;
        g2o_huber(e2, huber_delta, rho0, rho1);
        A.v[0] += rho0;
        A.v[36] += 1.0;
        A.v[37] += 2.0;
        add_row(A, e0.a, e0.v, rho1 * info);
        add_row(A, e1.a, e1.v, rho1 * info);
    }
    cta_reduce_store(A, partial + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * kVals);
}

__global__ void __launch_bounds__(256)
k_finish38(const double *__restrict__ partial, int nchunks, double *__restrict__ out) {
    __shared__ double tot[kVals];
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int v = warp; v < kVals; v += 8) {
        double x = 0.0;
        for (int i = lane; i < nchunks; i += 32) x += partial[((long long)b * nchunks + i) * kVals + v];
        for (int o = 16; o; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if (lane == 0) tot[v] = x;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double *o = out + (long long)b * STL_LIN_NSUMS;
        for (int i = 0; i < STL_LIN_NSUMS; ++i) o[i] = 0.0;
        o[0] = tot[0];
        for (int a = 0; a < 7; ++a) o[1 + a] = tot[1 + a];
        int h = 8;
        for (int a = 0; a < 7; ++a)
            for (int c = a; c < 7; ++c) { o[8 + a * 7 + c] = tot[h]; o[8 + c * 7 + a] = tot[h]; ++h; }
        o[57] = tot[36];  // active edges
        o[60] = tot[37];  // scalar residuals
    }
}

template <class T> cudaError_t upload(T *&d, const T *h, size_t n, cudaStream_t st) {
    cudaError_t e = cudaMalloc(&d, sizeof(T) * (n ? n : 1));
    if (e == cudaSuccess && n) e = cudaMemcpyAsync(d, h, sizeof(T) * n, cudaMemcpyHostToDevice, st);
    return e;
}

}  // namespace

cudaError_t he_linearize(const stl_he_edges_t &ed, const double *x, int B, double *h_out /*[B][62]*/, double *h_chi2, cudaStream_t st) {
    const int n = ed.n;
    std::vector<HeEdge> he((size_t)n);
    for (int i = 0; i < n; ++i) make_he_edge(ed.Ta + (size_t)i * 12, ed.Tb + (size_t)i * 12, ed.info ? ed.info[i] : 1.0, &he[i]);
    std::vector<DevCand> hc((size_t)B);
    for (int b = 0; b < B; ++b) make_candidate(x + (size_t)b * 7, &hc[b]);
    HeEdge *d_e = nullptr; DevCand *d_c = nullptr; double *d_x = nullptr, *d_p = nullptr, *d_o = nullptr, *d_chi = nullptr;
    const int chunks = n > 0 ? (n + kT - 1) / kT < 592 ? (n + kT - 1) / kT : 592 : 1;
    cudaError_t e = upload(d_e, he.data(), (size_t)n, st);
    if (e == cudaSuccess) e = upload(d_c, hc.data(), (size_t)B, st);
    if (e == cudaSuccess) e = upload(d_x, x, (size_t)B * 7, st);
    if (e == cudaSuccess) e = cudaMalloc(&d_p, sizeof(double) * kVals * chunks * B);
    if (e == cudaSuccess) e = cudaMalloc(&d_o, sizeof(double) * STL_LIN_NSUMS * B);
    if (e == cudaSuccess && h_chi2) e = cudaMalloc(&d_chi, sizeof(double) * (size_t)(n ? n : 1) * B);
    if (e == cudaSuccess) {
        k_he_linearize<<<dim3(chunks, B), kT, 0, st>>>(d_e, n, d_c, d_x, ed.huber_delta, ed.regulation, d_p, d_chi);
        k_finish38<<<B, 256, 0, st>>>(d_p, chunks, d_o);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_out, d_o, sizeof(double) * STL_LIN_NSUMS * B, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && h_chi2 && n) e = cudaMemcpyAsync(h_chi2, d_chi, sizeof(double) * (size_t)n * B, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_e); cudaFree(d_c); cudaFree(d_x); cudaFree(d_p); cudaFree(d_o); cudaFree(d_chi);
    return e;
}

cudaError_t calib_linearize(const stl_calib_edges_t &ed, const double *x, int B, double *h_out, double *h_chi2, cudaStream_t st) {
    const long long n = ed.n_edges;
    std::vector<CalibKf> hk((size_t)ed.n_kf);
    std::vector<int> ekf((size_t)(n ? n : 1));
    for (int f = 0; f < ed.n_kf; ++f) {
        make_calib_kf(ed.Tlw_quat + (size_t)f * 6, ed.intrinsics + (size_t)f * 4, &hk[f]);
        for (long long i = ed.edge_offset[f]; i < ed.edge_offset[f + 1]; ++i) ekf[(size_t)i] = f;
    }
    std::vector<CalibCand> hc((size_t)B);
    for (int b = 0; b < B; ++b) make_calib_candidate(x + (size_t)b * 7, &hc[b]);
    CalibKf *d_k = nullptr; int *d_ekf = nullptr; double *d_X = nullptr, *d_obs = nullptr, *d_p = nullptr, *d_o = nullptr, *d_chi = nullptr;
    float *d_is = nullptr; unsigned char *d_lv = nullptr; CalibCand *d_c = nullptr;
    const long long want = (n + kT - 1) / kT;
    const int chunks = (int)(want < 1 ? 1 : (want > 1184 ? 1184 : want));
    cudaError_t e = upload(d_k, hk.data(), hk.size(), st);
    if (e == cudaSuccess) e = upload(d_ekf, ekf.data(), (size_t)n, st);
    if (e == cudaSuccess) e = upload(d_X, ed.Xw, (size_t)n * 3, st);
    if (e == cudaSuccess) e = upload(d_obs, ed.obs, (size_t)n * 2, st);
    if (e == cudaSuccess) e = upload(d_is, ed.inv_sigma2, (size_t)n, st);
    if (e == cudaSuccess && ed.level) e = upload(d_lv, ed.level, (size_t)n, st);
    if (e == cudaSuccess) e = upload(d_c, hc.data(), (size_t)B, st);
    if (e == cudaSuccess) e = cudaMalloc(&d_p, sizeof(double) * kVals * chunks * B);
    if (e == cudaSuccess) e = cudaMalloc(&d_o, sizeof(double) * STL_LIN_NSUMS * B);
    if (e == cudaSuccess && h_chi2) e = cudaMalloc(&d_chi, sizeof(double) * (size_t)(n ? n : 1) * B);
    if (e == cudaSuccess) {
        k_calib_linearize<<<dim3(chunks, B), kT, 0, st>>>(d_k, d_ekf, d_X, d_obs, d_is, d_lv, n, d_c, ed.huber_delta, d_p, d_chi);
        k_finish38<<<B, 256, 0, st>>>(d_p, chunks, d_o);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_out, d_o, sizeof(double) * STL_LIN_NSUMS * B, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && h_chi2 && n) e = cudaMemcpyAsync(h_chi2, d_chi, sizeof(double) * (size_t)n * B, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_k); cudaFree(d_ekf); cudaFree(d_X); cudaFree(d_obs); cudaFree(d_is); cudaFree(d_lv); cudaFree(d_c); cudaFree(d_p); cudaFree(d_o); cudaFree(d_chi);
    return e;
}

}  // namespace stl
