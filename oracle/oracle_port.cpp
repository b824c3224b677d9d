// oracle_port.cpp — CPU ORACLE (test infrastructure, NOT product code).
// Self-contained build: KNN through the in-repo restatement of nanoflann.
#include "oracle_core.hpp"
#include "oracle_kdtree.hpp"

namespace {
template <int DIM>
struct PortTree {
    orc::KDTreePort<DIM> t;
    PortTree(const double *pts, size_t n, int leaf) : t(pts, n, leaf) {}
    size_t knn(const double *q, size_t k, uint32_t *idx, double *d2) const {
        orc::KnnSet rs(k, idx, d2);
        t.find(rs, q);
        return rs.count;
    }
};
}  // namespace
#define ORC_TREE2 PortTree<2>
#define ORC_TREE3 PortTree<3>
#define ORC_BACKEND_NAME "port"
#include "oracle_capi.inc"
