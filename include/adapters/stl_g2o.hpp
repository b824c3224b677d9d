// stl_g2o.hpp — the g2o edge of IBACalib.hpp, backed by the CUDA path.
//
// `IBAPlaneEdge : g2o::BaseUnaryEdge<20, g2o::VectorN<20>, VertexSim3>` (include/IBACalib.hpp:74-155) holds up to ten
// covisible observations, zero-pads its 20-vector (:133-137) and lets g2o differentiate it (G2O_MAKE_AUTO_AD_FUNCTIONS,
// :154); the vertex is VertexSim3 : BaseVertex<7, Vector7> with oplusImpl `_estimate += update` (g2o_tools.h:13-24).
// `StlPlaneEdge` is the same edge type with computeError() / linearizeOplus() served from the batch evaluation of all
// frozen blocks (stl_eval_blocks with rmax = 20 is exactly this layout).  `StlBlockStore::refresh` is called once per
// iterate — from g2o's preIteration action, or lazily by the first edge that sees a changed estimate.
#pragma once
#include <cstring>
#include <vector>

#include <g2o/core/base_unary_edge.h>
#include <g2o/core/base_vertex.h>

#include "../stlcalib_host.hpp"

namespace stl {

// VertexSim3 (g2o_tools.h:13-30): raw 7-vector, additive update
class StlVertexSim3 : public g2o::BaseVertex<7, g2o::Vector7> {
  public:
    void setToOriginImpl() override { _estimate.setZero(); }
    void oplusImpl(const double *update) override {
        for (int i = 0; i < 7; ++i) _estimate[i] += update[i];  // _estimate += update (g2o_tools.h:21-24)
    }
    bool read(std::istream &) override { return false; }
    bool write(std::ostream &) const override { return false; }
};

class StlBlockStore {
  public:
    explicit StlBlockStore(Context *ctx) : ctx_(ctx) {}
    size_t build(const double x0[7]) {
        const auto n = ctx_->associate(x0);
        total_ = n[0] + n[1] + n[2] + n[3];
        have_ = false;
        refresh(x0);
        return blocks_.size();
    }
    void refresh(const double x[7]) {
        if (have_ && std::memcmp(x, x_, sizeof(x_)) == 0) return;
        blocks_ = ctx_->eval_blocks(x, total_, 20);  // 20 = 2 x 10 covisible observations, zero padded (IBACalib.hpp:74)
        std::memcpy(x_, x, sizeof(x_));
        have_ = true;
    }
    const Context::Blocks &blocks() const { return blocks_; }

  private:
    Context *ctx_;
    int64_t total_ = 0;
    bool have_ = false;
    double x_[7] = {0, 0, 0, 0, 0, 0, 0};
    Context::Blocks blocks_;
};

class StlPlaneEdge : public g2o::BaseUnaryEdge<20, g2o::VectorN<20>, StlVertexSim3> {
  public:
    StlPlaneEdge(StlBlockStore *store, size_t index) : store_(store), i_(index) {}
    void computeError() override {
        const StlVertexSim3 *v = static_cast<const StlVertexSim3 *>(_vertices[0]);
        store_->refresh(v->estimate().data());
        const Context::Blocks &b = store_->blocks();
        for (int r = 0; r < 20; ++r) _error[r] = b.residuals[i_ * 20 + r];
    }
    void linearizeOplus() override {
        const StlVertexSim3 *v = static_cast<const StlVertexSim3 *>(_vertices[0]);
        store_->refresh(v->estimate().data());
        const Context::Blocks &b = store_->blocks();
        for (int r = 0; r < 20; ++r)
            for (int c = 0; c < 7; ++c) _jacobianOplusXi(r, c) = b.jacobians[(i_ * 20 + r) * 7 + c];
    }
    bool read(std::istream &) override { return false; }   // IBACalib.hpp:150-151
    bool write(std::ostream &) const override { return false; }

  private:
    StlBlockStore *store_;
    size_t i_;
};

}  // namespace stl
