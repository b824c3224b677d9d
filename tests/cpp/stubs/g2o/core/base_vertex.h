// TEST STUB of g2o's BaseVertex / BaseUnaryEdge (g2o is not installable here): the members that
// include/adapters/stl_g2o.hpp touches, shaped after g2o 20230223 as the reference uses it (include/IBACalib.hpp:74-155,
// include/g2o_tools.h:13-30).  Fixed-size storage stands in for Eigen.
#pragma once
#include <array>
#include <istream>
#include <ostream>
#include <vector>
namespace g2o {
template <int N> struct VectorN {
    std::array<double, N> v{};
    double &operator[](int i) { return v[i]; }
    const double &operator[](int i) const { return v[i]; }
    double *data() { return v.data(); }
    const double *data() const { return v.data(); }
    void setZero() { v.fill(0.0); }
};
typedef VectorN<7> Vector7;
template <int R, int C> struct MatrixRC {
    std::array<double, R * C> m{};
    double &operator()(int r, int c) { return m[r * C + c]; }
    const double &operator()(int r, int c) const { return m[r * C + c]; }
};
class VertexBase {
  public:
    virtual ~VertexBase() = default;
};
template <int D, typename T> class BaseVertex : public VertexBase {
  public:
    static const int Dimension = D;
    const T &estimate() const { return _estimate; }
    void setEstimate(const T &e) { _estimate = e; }
    virtual void setToOriginImpl() = 0;
    virtual void oplusImpl(const double *update) = 0;
    virtual bool read(std::istream &) = 0;
    virtual bool write(std::ostream &) const = 0;
    void oplus(const double *u) { oplusImpl(u); }
  protected:
    T _estimate;
};
}  // namespace g2o
