// Compiles the three optimiser adapters (include/adapters/*.hpp) against stand-in third-party headers
// (tests/cpp/stubs: NOMAD / Ceres / g2o are not installable here) and drives them the way iba_global / iba_local /
// a g2o optimiser would: NOMAD eval_x + eval_block, Ceres EvaluationCallback + per-block CostFunction + HuberLoss,
// g2o unary edges on a Sim3 vertex.  Prints what tests/test_cpp_shim.py compares with the ctypes path.
// Without a GPU it fails loudly (exit code 3).
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include "adapters/stl_ceres.hpp"
#include "adapters/stl_g2o.hpp"
#include "adapters/stl_nomad.hpp"
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
        // ---- NOMAD (iba_global.cpp:583-587: evaluator handed over as unique_ptr<Evaluator>)
        std::unique_ptr<NOMAD::Evaluator> ev(new stl::NomadBALoss(std::make_shared<NOMAD::EvalParameters>(), &ctx));
        NOMAD::Block block;
        for (int b = 0; b < B; ++b) {
            auto pt = std::make_shared<NOMAD::EvalPoint>(7);
            for (int i = 0; i < 7; ++i) (*pt)[i] = NOMAD::Double(X[b * 7 + i]);
            block.push_back(pt);
        }
        bool count = false;
        NOMAD::EvalPoint x0 = *block[1];
        const bool ok = ev->eval_x(x0, NOMAD::Double(1e20), count);
        std::printf("NOMAD_X %d %d %s\n", (int)ok, (int)count, x0.getBBO().c_str());
        std::vector<bool> counts;
        const std::vector<bool> oks = ev->eval_block(block, NOMAD::Double(1e20), counts);
        for (int b = 0; b < B; ++b) std::printf("NOMAD_B %d %d %d %s\n", b, (int)oks[b], (int)counts[b], block[b]->getBBO().c_str());
        // ---- Ceres (iba_local.cpp:263-309,434-446)
        double params[7];
        for (int i = 0; i < 7; ++i) params[i] = X[i];
        stl::StlEvaluationCallback cb(&ctx, params, 6);
        const size_t nblocks = cb.Build(params);
        std::vector<std::unique_ptr<ceres::CostFunction>> costs;
        std::vector<std::unique_ptr<ceres::LossFunction>> losses;
        for (size_t i = 0; i < nblocks; ++i) {
            costs.emplace_back(new stl::StlBlockCost(&cb, i));
            losses.emplace_back(new ceres::HuberLoss(cb.huber_delta(i)));
        }
        for (int i = 0; i < 7; ++i) params[i] = X[7 + i];  // the solver moved the point
        cb.PrepareForEvaluation(true, true);
        double cost = 0, g0 = 0, H00 = 0;
        for (size_t i = 0; i < nblocks; ++i) {  // what ceres::Problem::Evaluate assembles (residual_block.cc + corrector.cc, rho'' <= 0)
            double r[20], J[20 * 7];
            double *jac[1] = {J};
            const double *pp[1] = {params};
            costs[i]->Evaluate(pp, r, jac);
            const int nr = costs[i]->num_residuals();
            double sq = 0;
            for (int k = 0; k < nr; ++k) sq += r[k] * r[k];
            double rho[3];
            losses[i]->Evaluate(sq, rho);
            cost += 0.5 * rho[0];
            const double sr = std::sqrt(rho[1]);
            for (int k = 0; k < nr; ++k) { g0 += (sr * J[k * 7]) * (sr * r[k]); H00 += (sr * J[k * 7]) * (sr * J[k * 7]); }
        }
        std::printf("CERES %zu %.17g %.17g %.17g\n", nblocks, cost, g0, H00);
        // ---- g2o (IBACalib.hpp:74-155)
        stl::StlBlockStore store(&ctx);
        stl::StlVertexSim3 v;
        g2o::Vector7 est;
        for (int i = 0; i < 7; ++i) est[i] = X[i];
        v.setEstimate(est);
        const size_t ne = store.build(est.data());
        std::vector<std::unique_ptr<stl::StlPlaneEdge>> edges;
        for (size_t i = 0; i < ne; ++i) { edges.emplace_back(new stl::StlPlaneEdge(&store, i)); edges.back()->setVertex(0, &v); }
        double upd[7];
        for (int i = 0; i < 7; ++i) upd[i] = X[7 + i] - X[i];
        v.oplus(upd);  // VertexSim3::oplusImpl is plain addition
        double chi2 = 0, j00 = 0;
        for (auto &e : edges) { e->computeError(); e->linearizeOplus(); chi2 += e->chi2(); j00 += e->jacobianOplusXi()(0, 0); }
        std::printf("G2O %zu %.17g %.17g\n", ne, chi2, j00);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "adapters: %s\n", e.what());
        stl_synth_destroy(S);
        return 3;
    }
    stl_synth_destroy(S);
    return 0;
}
