// p2p.cuh — the all-reduce of the per-candidate record, done by the LAST kernel of a call over NVLink peer memory.
//
// The exchange of the path is tiny (12, 62 or 74 doubles per candidate) and sits at the end of a chain of short kernels:
// on a keyframe shard of an 8-GPU run a separate collective launch costs as much as a compute stage.  So the kernel that
// produces a candidate's record (k_reduce / k_lin_finish, one CTA per candidate) also exchanges it: the CTA stores the
// record into a slot of EVERY rank's receive buffer (cudaIpc-mapped peer memory, plain stores over NVLink / NVSwitch),
// fences, raises a per-(rank, candidate) flag on every rank, waits for the flags of all ranks in its own buffer and adds
// the records in RANK ORDER — the same order on every rank, so all ranks hold bit-identical totals.  Buffers are
// double-buffered by the parity of a call sequence number (a rank can be at most one exchange ahead of the slowest one).
// ncclAllReduce (capi.cu) remains the path for anything that does not fit (B > kP2pMaxB, peer mapping unavailable,
// STL_NO_P2P=1); a wait that exceeds ~2 s raises an error flag instead of hanging the GPU.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace stl {

constexpr int kP2pMaxRanks = 8;
constexpr int kP2pMaxB = 256;   // candidates per exchange (one CTA each)
constexpr int kP2pWidth = 80;   // doubles per slot (>= STL_STEP_NSUMS)

struct P2pView {
    int n = 0;              // ranks (0: no exchange in this kernel)
    int rank = 0;
    unsigned seq = 0;       // number of this exchange, the same on every rank
    int off = 0, width = 0; // the record of candidate b starts `off` doubles before/after the kernel's own output row
    double *data[kP2pMaxRanks] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};    // [2][n][kP2pMaxB][kP2pWidth] on rank r
    unsigned *flag[kP2pMaxRanks] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // [2][n][kP2pMaxB] on rank r
    int *err = nullptr;     // local: set when a wait timed out
};

#ifdef __CUDACC__
// Called by every thread of the CTA that owns candidate b; rec = this candidate's record in local global memory
// (already written by this CTA or by an earlier kernel).  On return rec holds the sum over the ranks.
__device__ __forceinline__ void p2p_allreduce_record(const P2pView &P, int b, double *rec) {
    if (P.n <= 1) return;
    __threadfence_block();
    __syncthreads();  // the record is complete
    const int par = (int)(P.seq & 1u), w = P.width;
    const long long slot = ((long long)(par * P.n + P.rank) * kP2pMaxB + b);
    for (int i = threadIdx.x; i < w * P.n; i += blockDim.x) {
        const int r = i / w, k = i - r * w;
        P.data[r][slot * kP2pWidth + k] = rec[k];
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < P.n) {
        *reinterpret_cast<volatile unsigned *>(P.flag[threadIdx.x] + slot) = P.seq;  // raise my flag on rank threadIdx.x
        volatile unsigned *f = P.flag[P.rank] + ((long long)(par * P.n + (int)threadIdx.x) * kP2pMaxB + b);
        const long long t0 = clock64();
        while (*f != P.seq) {
            if (clock64() - t0 > 4000000000ll) { atomicExch(P.err, 1); break; }
        }
    }
    __threadfence_system();
    __syncthreads();
    for (int k = threadIdx.x; k < w; k += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < P.n; ++r)
            s += *reinterpret_cast<volatile double *>(P.data[P.rank] + ((long long)(par * P.n + r) * kP2pMaxB + b) * kP2pWidth + k);
        rec[k] = s;
    }
}
#endif

}  // namespace stl
