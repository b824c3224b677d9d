// lm.h — LM path (Ceres / g2o side of the boundary): frozen residual blocks + linearisation.
#pragma once
#include "kernels.h"

namespace stl {

// Residual blocks frozen by stl_associate (BuildProblem, src/examples/iba_local.cpp:145-323).
// One slot per map-point-carrying correspondence of the association pass (slot = mp_off[kf] + qi); dense index
// lists select the slots that carry a 3-D/2-D block (IBA_PlaneFactor) and a 3-D/3-D block
// (Point2Point_Factor / Point2Plane_Factor).
struct LmState {
    bool ready = false;
    long long n_blocks[4] = {0, 0, 0, 0};  // plane (3-D/2-D), point-to-point, point-to-plane, GPR (3-D/2-D)
    long long n_slots = 0;
    int sub = 4;                  // CTAs per keyframe of the association kernels (= DevWork::sub)
    int *slot_kf = nullptr;       // [n_slots]
    uint32_t *slot_kp = nullptr;  // [n_slots]
    uint8_t *flags = nullptr;     // one allocation: flag2d | type3d | flag3d | flagG | d_counts (cleared by one memset)
    size_t flags_bytes = 0;
    uint8_t *flag2d = nullptr;    // [n_slots] 1 = plane block
    uint8_t *type3d = nullptr;    // [n_slots] 0 none, 1 point-to-point, 2 point-to-plane
    uint8_t *flag3d = nullptr;    // [n_slots] type3d != 0
    double *geo2d = nullptr;      // [n_slots][6]  p0 (scan point, LiDAR frame), n0 (normal)
    double *geo3d = nullptr;      // [n_slots][9]  map point (camera frame, unscaled), query point, normal
    // association scratch, one slot per map-point-carrying correspondence (mp_off[f] + qi)
    uint8_t *stage = nullptr;     // 1 = plane at the scan point passed the gates, 3-D search pending
    double *plane_a = nullptr;    // [n_mp][4] normal + regression error at the scan point
    uint32_t *nnb_pos = nullptr;  // [n_mp] 1-NN of the map point (0xffffffff = beyond max_3d_dist)
    uint32_t *nbb = nullptr;      // [n_mp][32] neighbour list of that point
    float4 *nbbx = nullptr;       // [32][n_mp] coordinates of those neighbours, transposed
    long long nbbx_stride = 0;
    int *nbb_m = nullptr;         // [n_mp] (-2 = same point as the associated one)
    double *nbb_last = nullptr;
    int *idx2d = nullptr, *idx3d = nullptr;  // [n_slots] dense slot lists (idx3d: on demand, lm_compact3d)
    bool idx3d_valid = false;
    // IBA_GPRFactor blocks (use_gpr): flag + dense list by correspondence slot, neighbour list by map-point slot
    uint8_t *flagG = nullptr;
    int *idxG = nullptr, *slot_mp = nullptr;
    uint32_t *gpr_nb = nullptr;   // [n_mp][32]
    int *gpr_m = nullptr;         // [n_mp]
    double *gpr_hyper = nullptr;  // [n_mp][2] per-block (sigma, l) fitted at association time (null: the per-problem values)
    int nG = 0;
    int *d_counts = nullptr;      // [4] device: plane blocks, 3-D blocks, point-to-point among them, GPR blocks
    int *h_counts = nullptr;      // pinned copy, valid behind counts_done
    cudaEvent_t counts_done = nullptr;
    bool counts_valid = false;    // n2d / n3d / nG / n_blocks below mirror the device counts
    bool use_gpr = false;
    long long max_blocks = 0;     // upper bound of any block count (map-point-carrying keypoints)
    int n2d = 0, n3d = 0;
    void *d_tmp = nullptr;        // cub scratch
    size_t tmp_bytes = 0;
    double *partial = nullptr;    // [B][grid][kLinVals]
    long long partial_cap = 0;
    void *d_cand = nullptr;       // [B] LmCand
    void *h_cand = nullptr;       // pinned
    int cand_cap = 0;
    cudaEvent_t h2d_done = nullptr;  // guards the pinned staging buffer against reuse while a copy is in flight
};

void lm_free(LmState &lm);
// wk must hold the K1 result (correspondences) of the association extrinsic at candidate slot 0.
// nn_hint / nn_g2 (optional): per map-point slot, the 1-NN position an evaluation at the SAME extrinsic found and its
// lower bound of the squared distance to any other scan point (Sink1::g2): seeds, or settles, the 3-D search
// part: 0 = everything; 1 = up to the plane at the associated scan point (needs K1's result only); 2 = the rest (needs nn_hint /
// nn_g2, i.e. K2a, when given) — the overlapped step enqueues the two parts around a wait for K2a
cudaError_t lm_associate(const DevPack &pk, const DevWork &wk, const DevParams &pr, LmState &lm, cudaStream_t st,
                         const uint32_t *nn_hint = nullptr, const float *nn_g2 = nullptr, int part = 0, bool nn_folded = false,
                         bool defer_counts = false);
// defer_counts: the block-count read-back is left to lm_copy_counts (after whatever the caller enqueues next)
cudaError_t lm_copy_counts(LmState &lm, cudaStream_t st);
cudaError_t lm_stage_candidates(LmState &lm, const double *x, int B, cudaStream_t st);
// nn_folded: LmState::nnb_pos / nbb_m were written by K2a (launch_align3d with lm_pos / lm_m): k_lm_knn_b is not launched
cudaError_t lm_reserve(const DevPack &pk, const DevParams &pr, LmState &lm, cudaStream_t st);
// x: HOST [B][7]; d_out: DEVICE [B][STL_LIN_NSUMS]
// optional per-block output of a linearisation (device pointers; B must be 1)
struct BlockOut {
    int32_t *type = nullptr, *kf = nullptr, *kp = nullptr, *nres = nullptr;  // [n_blocks]
    double *res = nullptr;  // [n_blocks][rmax]
    double *jac = nullptr;  // [n_blocks][rmax][7]
    int rmax = 0;
};

// out_stride: doubles between the records of consecutive candidates in d_out (0 = STL_LIN_NSUMS)
cudaError_t lm_linearize(const DevPack &pk, const DevParams &pr, LmState &lm, const double *x, int B, double *d_out, cudaStream_t st,
                         const BlockOut *blocks = nullptr, int out_stride = 0, const P2pView *p2p = nullptr, cudaEvent_t before_finish = nullptr,
                         bool cand_staged = false);
cudaError_t lm_compact3d(const DevPack &pk, LmState &lm, cudaStream_t st);
// cand_staged: the caller has already run lm_stage_candidates(lm, x, B, st) for exactly these x on this stream
// before_finish (optional): event the finishing kernel waits for (the rest of the record it completes / exchanges)
// waits for the last association and mirrors its block counts into lm.n2d / n3d / nG / n_blocks
cudaError_t lm_block_counts(LmState &lm);
// GPR::fit for every GPR block of the last association (IBA_GPRFactor's constructor, IBACalib2.hpp:441-461): the
// training pixels / depths are formed on the device with the association extrinsic (wk.cand[0]), the two-parameter
// fit runs on the host (OpenMP over blocks), the result lands in lm.gpr_hyper.  Synchronous.
cudaError_t lm_fit_gpr_hyper(const DevPack &pk, const DevWork &wk, const DevParams &pr, LmState &lm, cudaStream_t st, int flavour);
// copies (sigma, l) of the GPR blocks, block order, to host memory out[nG][2]
cudaError_t lm_get_gpr_hyper(const DevParams &pr, LmState &lm, double *out, cudaStream_t st);

}  // namespace stl
