// TEST STUB of <ceres/ceres.h> (Ceres is not installable here): the slice of the interface that
// include/adapters/stl_ceres.hpp touches, shaped after Ceres' public headers (cost_function.h, evaluation_callback.h,
// loss_function.h) as the reference uses them (src/examples/iba_local.cpp:263-309,434-446).  Not a solver.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>
namespace ceres {
class CostFunction {
  public:
    virtual ~CostFunction() = default;
    virtual bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const = 0;
    const std::vector<int32_t> &parameter_block_sizes() const { return sizes_; }
    int num_residuals() const { return nres_; }
  protected:
    std::vector<int32_t> *mutable_parameter_block_sizes() { return &sizes_; }
    void set_num_residuals(int n) { nres_ = n; }
  private:
    std::vector<int32_t> sizes_;
    int nres_ = 0;
};
class EvaluationCallback {
  public:
    virtual ~EvaluationCallback() = default;
    virtual void PrepareForEvaluation(bool evaluate_jacobians, bool new_evaluation_point) = 0;
};
class LossFunction {
  public:
    virtual ~LossFunction() = default;
    virtual void Evaluate(double sq_norm, double out[3]) const = 0;
};
class HuberLoss : public LossFunction {  // loss_function.cc
  public:
    explicit HuberLoss(double a) : a_(a), b_(a * a) {}
    void Evaluate(double s, double rho[3]) const override {
        if (s > b_) {
            const double r = std::sqrt(s);
            rho[0] = 2.0 * a_ * r - b_;
            rho[1] = std::max(std::numeric_limits<double>::min(), a_ / r);
            rho[2] = -rho[1] / (2.0 * s);
        } else {
            rho[0] = s; rho[1] = 1.0; rho[2] = 0.0;
        }
    }
  private:
    double a_, b_;
};
}  // namespace ceres
