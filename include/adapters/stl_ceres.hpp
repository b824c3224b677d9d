// stl_ceres.hpp — the Ceres problem of iba_local, backed by the CUDA path.
//
// iba_local builds one residual block per frozen correspondence — IBA_PlaneFactor / Point2Point_Factor /
// Point2Plane_Factor (/ IBA_GPRFactor) wrapped in ceres::HuberLoss (src/examples/iba_local.cpp:263-309) — over ONE
// 7-double parameter block, and solves with options.num_threads = hardware_concurrency() (:439).  Here the blocks are
// frozen on the GPU by stl_associate (BuildProblem) and ALL of them are evaluated by one stl_eval_blocks call per
// iterate: `StlEvaluationCallback::PrepareForEvaluation` runs it when Ceres announces a new point, and every
// `StlBlockCost::Evaluate` only copies its rows out (row-major Jacobians, as Ceres expects).  Ceres keeps applying its
// own loss function and its own trust-region logic: nothing above the cost functions changes.
//   problem options: options.evaluation_callback = &callback   (ceres::Problem::Options)
//   per block:       problem.AddResidualBlock(new StlBlockCost(&callback, i), new ceres::HuberLoss(delta_i), params)
#pragma once
#include <cstring>
#include <vector>

#include <ceres/ceres.h>

#include "../stlcalib_host.hpp"

namespace stl {

class StlEvaluationCallback : public ceres::EvaluationCallback {
  public:
    // params: the 7-double block Ceres optimises (the pointer handed to AddResidualBlock)
    StlEvaluationCallback(Context *ctx, const double *params, int rmax) : ctx_(ctx), params_(params), rmax_(rmax) {}

    // BuildProblem at x0 (iba_local.cpp:443): returns the number of residual blocks to add
    size_t Build(const double x0[7]) {
        n_ = ctx_->associate(x0);
        total_ = n_[0] + n_[1] + n_[2] + n_[3];
        valid_ = false;
        Refresh();
        return blocks_.size();
    }
    void PrepareForEvaluation(bool /*evaluate_jacobians*/, bool new_evaluation_point) override {
        if (new_evaluation_point || !valid_) Refresh();
    }
    const Context::Blocks &blocks() const { return blocks_; }
    int rmax() const { return rmax_; }
    // Huber delta of block i: robust_kernel_delta for the 3-D/2-D factors, robust_kernel_3ddelta for the 3-D/3-D ones
    double huber_delta(size_t i) const {
        const int t = blocks_.type[i];
        return (t == 0 || t == 3) ? ctx_->params().robust_kernel_delta : ctx_->params().robust_kernel_3ddelta;
    }

  private:
    void Refresh() {
        blocks_ = ctx_->eval_blocks(params_, total_, rmax_);
        valid_ = true;
    }
    Context *ctx_;
    const double *params_;
    int rmax_;
    std::array<int64_t, 4> n_{};
    int64_t total_ = 0;
    bool valid_ = false;
    Context::Blocks blocks_;
};

// One residual block of the frozen problem: what IBA_PlaneFactor::Create / Point2Point_Factor::Create /
// Point2Plane_Factor::Create return (IBACalib2.hpp:202-212,592,635), served from the batch evaluation.
class StlBlockCost : public ceres::CostFunction {
  public:
    StlBlockCost(const StlEvaluationCallback *cb, size_t index) : cb_(cb), i_(index) {
        set_num_residuals(cb->blocks().n_res[index]);
        mutable_parameter_block_sizes()->push_back(7);
    }
    bool Evaluate(double const *const * /*parameters*/, double *residuals, double **jacobians) const override {
        const Context::Blocks &b = cb_->blocks();
        const int nr = b.n_res[i_], rmax = b.rmax;
        std::memcpy(residuals, &b.residuals[i_ * rmax], sizeof(double) * nr);
        if (jacobians && jacobians[0]) std::memcpy(jacobians[0], &b.jacobians[i_ * rmax * 7], sizeof(double) * nr * 7);
        return true;
    }

  private:
    const StlEvaluationCallback *cb_;
    size_t i_;
};

}  // namespace stl
