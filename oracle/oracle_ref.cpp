// oracle_ref.cpp — CPU ORACLE (test infrastructure, NOT product code).
// Same oracle, but every neighbour search goes through the REFERENCE's own
// vendored nanoflann v1.5.0 and its KDTreeVectorOfVectorsAdaptor, compiled from
// the sources where they lie (-I/root/reference/include); the binary goes to
// oracle/_ref/ (git-ignored).  Types as in iba_global.cpp:18-23.
#include <array>
#include <vector>

#include "KDTreeVectorOfVectorsAdaptor.h"
#include "nanoflann.hpp"
#include "oracle_core.hpp"

namespace {
template <int DIM>
struct RefTree {
    typedef std::vector<std::array<double, DIM>> Vec;
    typedef nanoflann::KDTreeVectorOfVectorsAdaptor<Vec, double, DIM, nanoflann::metric_L2_Simple, std::uint32_t> KD;
    Vec data;
    std::unique_ptr<KD> kd;
    RefTree(const double *pts, size_t n, int leaf) : data(n) {
        for (size_t i = 0; i < n; ++i)
            for (int d = 0; d < DIM; ++d) data[i][d] = pts[i * DIM + d];
        if (n > 0) kd.reset(new KD(DIM, data, leaf));  // the adaptor asserts on an empty set (KDTreeVectorOfVectorsAdaptor.h:86)
    }
    size_t knn(const double *q, size_t k, uint32_t *idx, double *d2) const {
        if (!kd) return 0;
        nanoflann::KNNResultSet<double, std::uint32_t> rs(k);
        rs.init(idx, d2);
        kd->index->findNeighbors(rs, q, nanoflann::SearchParameters());
        return rs.size();
    }
};
}  // namespace
#define ORC_TREE2 RefTree<2>
#define ORC_TREE3 RefTree<3>
#define ORC_BACKEND_NAME "nanoflann-1.5.0(reference)"
#include "oracle_capi.inc"
