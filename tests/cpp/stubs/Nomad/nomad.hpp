// TEST STUB of <Nomad/nomad.hpp> (NOMAD 4 is not installable here): only the slice of the interface that
// include/adapters/stl_nomad.hpp touches, shaped after NOMAD 4's public headers as the reference uses them
// (src/examples/iba_global.cpp:13-14,346-405,551-599).  Not a NOMAD implementation.
#pragma once
#include <memory>
#include <string>
#include <vector>
namespace NOMAD {
class Double {
  public:
    Double() = default;
    Double(double v) : v_(v) {}  // NOLINT
    double todouble() const { return v_; }
  private:
    double v_ = 0.0;
};
enum class EvalType { BB, SURROGATE, MODEL };
class EvalParameters {};
class EvalPoint {
  public:
    explicit EvalPoint(size_t n = 0) : x_(n) {}
    Double &operator[](size_t i) { return x_[i]; }
    const Double &operator[](size_t i) const { return x_[i]; }
    size_t size() const { return x_.size(); }
    void setBBO(const std::string &bbo) { bbo_ = bbo; }
    const std::string &getBBO() const { return bbo_; }
  private:
    std::vector<Double> x_;
    std::string bbo_;
};
typedef std::vector<std::shared_ptr<EvalPoint>> Block;
class Evaluator {
  public:
    Evaluator(const std::shared_ptr<EvalParameters> &p, EvalType t) : params_(p), type_(t) {}
    virtual ~Evaluator() = default;
    virtual bool eval_x(EvalPoint &x, const Double &hMax, bool &countEval) const = 0;
    virtual std::vector<bool> eval_block(Block &block, const Double &hMax, std::vector<bool> &countEval) const {
        std::vector<bool> ok(block.size());
        countEval.assign(block.size(), false);
        for (size_t i = 0; i < block.size(); ++i) { bool c = false; ok[i] = eval_x(*block[i], hMax, c); countEval[i] = c; }
        return ok;
    }
  protected:
    std::shared_ptr<EvalParameters> params_;
    EvalType type_;
};
}  // namespace NOMAD
