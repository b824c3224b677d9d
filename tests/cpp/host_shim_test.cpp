// Exercises include/stlcalib_host.hpp exactly the way a reference-side caller would
// (BAError tuple, BALoss::eval_x BBO, LMProblem), on a synthetic pack.  Prints one line per
// candidate; tests/test_gpu_cpp_shim.py compares it with the Python/ctypes path.
// Without a GPU it must fail loudly (exit code 3), never fall back.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "stlcalib_host.hpp"
#include "stlsynth.h"

int main(int argc, char **argv) {
    const int nkf = argc > 1 ? std::atoi(argv[1]) : 3;
    stl_synth_cfg_t cfg;
    stl_synth_default_cfg(&cfg);
    cfg.n_kf = nkf; cfg.n_kf_total = nkf; cfg.beams = 32; cfg.az_steps = 900; cfg.n_kp = 500; cfg.seed = 21;
    stl_synth_t *S = stl_synth_create(&cfg);
    if (!S) return 2;
    const int B = 3;
    std::vector<double> X(B * 7);
    stl_synth_candidates(stl_synth_x_gt(S), 42, B, 0.4, X.data());
    stl_params_t p;
    stl_default_params(&p);
    try {
        stl::Context ctx(p, 0);
        ctx.upload(*stl_synth_pack(S));
        stl::BALoss loss(ctx);
        for (int b = 0; b < B; ++b) {
            auto [f1, f2, C, valid, cnt] = stl::BAError(&X[b * 7], ctx);
            double bbo[4];
            bool count = false;
            const bool ok = loss.eval_x(&X[b * 7], bbo, count);
            std::printf("BA %d %.17g %.17g %.17g %d %d | BBO %d %d %s\n", b, f1, f2, C, valid, cnt, (int)ok, (int)count,
                        stl::BALoss::bbo_string(bbo).c_str());
        }
        stl::LMProblem prob(ctx);
        auto nb = prob.build(&X[0]);
        stl_lin_sums_t L = prob.evaluate(&X[7]);
        std::printf("LM %lld %lld %lld %.17g %.17g %.17g\n", (long long)nb[0], (long long)nb[1], (long long)nb[2], L.cost, L.g[0], L.H[0]);
        // per-block view (what a ceres::CostFunction / g2o edge returns): block count, first block, sum of squares
        const stl::Context::Blocks B = prob.blocks(&X[7], 6);
        double ss = 0;
        for (size_t i = 0; i < B.size(); ++i)
            for (int r = 0; r < B.n_res[i]; ++r) ss += B.residuals[i * B.rmax + r] * B.residuals[i * B.rmax + r];
        std::printf("BLK %zu %d %d %d %.17g %.17g\n", B.size(), B.type[0], B.n_res[0], B.kp[0], B.residuals[0], ss);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "host shim: %s\n", e.what());
        stl_synth_destroy(S);
        return 3;
    }
    stl_synth_destroy(S);
    return 0;
}
