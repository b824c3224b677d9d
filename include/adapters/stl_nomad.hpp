// stl_nomad.hpp — the NOMAD 4 evaluator of iba_global, backed by the CUDA path.
//
// Drop-in for `class BALoss : public NOMAD::Evaluator` (src/examples/iba_global.cpp:346-405): same base class, same
// constructor convention (evalParams + EvalType::BB), same eval_x contract — inputs x[i].todouble(), i < 7 (:383-384),
// output x.setBBO("f C1 C2 C3") (:389-393), countEval = true, return true — and handed over the same way
// (`setEvaluator(std::move(unique_ptr))`, :583-587).  eval_block is NOMAD 4's batched virtual: a whole MADS poll goes to
// the GPU in ONE stl_eval_batch call, which is where the 256-candidate throughput comes from.
// Compiles against the real <Nomad/nomad.hpp>; the repository's tests compile it against tests/cpp/stubs (NOMAD is not
// installable here).
#pragma once
#include <memory>
#include <string>
#include <vector>

#include <Nomad/nomad.hpp>

#include "../stlcalib_host.hpp"

namespace stl {

class NomadBALoss : public NOMAD::Evaluator {
  public:
    NomadBALoss(const std::shared_ptr<NOMAD::EvalParameters> &evalParams, Context *ctx)
        : NOMAD::Evaluator(evalParams, NOMAD::EvalType::BB), ctx_(ctx) {}
    ~NomadBALoss() override = default;

    bool eval_x(NOMAD::EvalPoint &x, const NOMAD::Double & /*hMax*/, bool &countEval) const override {
        double xvec[7], bbo[4];
        for (int i = 0; i < 7; ++i) xvec[i] = x[i].todouble();  // iba_global.cpp:383-384
        BALoss(*ctx_).eval_x(xvec, bbo, countEval);
        x.setBBO(BALoss::bbo_string(bbo));                      // iba_global.cpp:389-393
        return true;
    }

    // one device call for the whole block (NOMAD::Evaluator::eval_block, the default implementation loops over eval_x)
    std::vector<bool> eval_block(NOMAD::Block &block, const NOMAD::Double & /*hMax*/, std::vector<bool> &countEval) const override {
        const int B = (int)block.size();
        std::vector<double> X((size_t)B * 7), bbo((size_t)B * 4);
        for (int b = 0; b < B; ++b)
            for (int i = 0; i < 7; ++i) X[(size_t)b * 7 + i] = (*block[b])[i].todouble();
        BALoss(*ctx_).eval_block(X.data(), B, bbo.data(), countEval);
        for (int b = 0; b < B; ++b) block[b]->setBBO(BALoss::bbo_string(&bbo[(size_t)b * 4]));
        return std::vector<bool>((size_t)B, true);
    }

  private:
    Context *ctx_;  // non-owning, like every pointer member of the reference's BALoss (iba_global.cpp:399-404)
};

}  // namespace stl
