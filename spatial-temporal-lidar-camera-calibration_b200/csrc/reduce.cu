// reduce.cu — K3: per-candidate reduction of the per-keyframe partial records into the
// accumulators of BAError (src/examples/iba_global.cpp:175-186,239-251,318-326).
// fp64, fixed order (strided partials per thread, then a fixed shuffle tree): the
// result is bit-reproducible run to run.  One CTA per candidate.
#include "../../include/stlcalib.h"
#include "kernels.h"
#include "p2p.cuh"

namespace stl {
namespace {

constexpr int kT = 256;

__global__ void __launch_bounds__(kT) k_reduce(const DevPack pk, const DevWork wk, const DevParams pr, double *__restrict__ out, const int out_stride, const P2pView P) {
    const int b = blockIdx.x;
    const int F = pk.n_kf, sub = wk.sub;
    double v[STL_EVAL_NSUMS];
#pragma unroll
    for (int i = 0; i < STL_EVAL_NSUMS; ++i) v[i] = 0;
    for (int f = threadIdx.x; f < F; f += kT) {
        const FrameRec r = wk.frame[(long long)b * F + f];
        v[0] += r.s2d; v[2] += r.she; v[3] += r.che; v[4] += r.c2d; v[5] += r.v2d; v[10] += r.kept; v[11] += r.ncorr;
        if (pr.w1 <= 1e-10) {  // iba_global.cpp:214-219: no 3-D term, one "valid" count per kept frame
            v[6] += r.kept; v[7] += r.kept;
        } else {
            for (int j = 0; j < sub; ++j) {
                const AlignRec a = wk.align[((long long)b * F + f) * sub + j];
                v[1] += a.s3d; v[6] += a.c3d; v[7] += a.v3d; v[8] += a.vpl; v[9] += a.vpt;
            }
        }
    }
    __shared__ double red[STL_EVAL_NSUMS][kT / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < STL_EVAL_NSUMS; ++i) {
        double x = v[i];
        for (int o = 16; o; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if (lane == 0) red[i][warp] = x;
    }
    __syncthreads();
    if (threadIdx.x < STL_EVAL_NSUMS) {
        double x = 0;
        for (int w = 0; w < kT / 32; ++w) x += red[threadIdx.x][w];
        out[(long long)b * out_stride + threadIdx.x] = x;
    }
    if (P.n > 1) p2p_allreduce_record(P, b, out + (long long)b * out_stride + P.off);  // keyframes sharded over GPUs: sum the shards here
}

}  // namespace

cudaError_t launch_reduce(const DevPack &pk, const DevWork &wk, const DevParams &pr, int B, double *d_out, cudaStream_t st, int out_stride,
                          const P2pView *p2p) {
    if (B <= 0) return cudaSuccess;
    k_reduce<<<B, kT, 0, st>>>(pk, wk, pr, d_out, out_stride > 0 ? out_stride : STL_EVAL_NSUMS, p2p ? *p2p : P2pView());
    return cudaGetLastError();
}

}  // namespace stl
