// lm.h — LM path (Ceres / g2o side of the boundary): frozen residual blocks + linearisation.
#pragma once
#include "kernels.h"

namespace stl {

// Residual block frozen by stl_associate (BuildProblem, iba_local.cpp:263-309). SoA on the device.
struct LmState {
    bool ready = false;
    long long n_blocks[3] = {0, 0, 0};  // plane (3-D/2-D), point-to-point, point-to-plane
    long long cap = 0;
    int *count = nullptr;       // device counters [3] + total
    int *b_type = nullptr;      // [cap] 0/1/2
    int *b_kf = nullptr;        // [cap]
    uint32_t *b_kp = nullptr;   // [cap]
    double *b_geo = nullptr;    // [cap][9]: plane: p0[3], n0[3], -  |  3-D: map_pt[3], query_pt[3], normal[3]
    double *partial = nullptr;  // linearisation partials
    long long partial_cap = 0;
    double *d_x = nullptr;      // [B][LmCand] device candidates
    long long x_cap = 0;
    void *h_x = nullptr;        // pinned staging
};

void lm_free(LmState &lm);
cudaError_t lm_associate(const DevPack &pk, const DevWork &wk, const DevParams &pr, LmState &lm, cudaStream_t st);
cudaError_t lm_linearize(const DevPack &pk, const DevParams &pr, LmState &lm, const double *x, int B, double *d_out, cudaStream_t st);

}  // namespace stl
