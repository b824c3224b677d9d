// TEST STUB, see base_vertex.h
#pragma once
#include "base_vertex.h"
namespace g2o {
template <int D, typename E, typename VertexXi> class BaseUnaryEdge {
  public:
    static const int Dimension = D;
    virtual ~BaseUnaryEdge() = default;
    void setVertex(size_t i, VertexBase *v) { if (_vertices.size() <= i) _vertices.resize(i + 1); _vertices[i] = v; }
    virtual void computeError() = 0;
    virtual void linearizeOplus() = 0;
    virtual bool read(std::istream &) = 0;
    virtual bool write(std::ostream &) const = 0;
    const E &error() const { return _error; }
    const MatrixRC<D, VertexXi::Dimension> &jacobianOplusXi() const { return _jacobianOplusXi; }
    double chi2() const { double s = 0; for (int i = 0; i < D; ++i) s += _error[i] * _error[i]; return s; }  // identity information
  protected:
    std::vector<VertexBase *> _vertices;
    E _error;
    MatrixRC<D, VertexXi::Dimension> _jacobianOplusXi;
};
}  // namespace g2o
