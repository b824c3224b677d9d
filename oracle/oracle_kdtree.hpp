// oracle_kdtree.hpp — CPU ORACLE (test infrastructure, NOT product code).
//
// Restatement of the KD-tree the reference uses for every neighbour search:
// nanoflann v1.5.0 (vendored at include/nanoflann.hpp, nanoflann.hpp:63)
// through KDTreeVectorOfVectorsAdaptor (include/KDTreeVectorOfVectorsAdaptor.h:81)
// with metric_L2_Simple, IndexType = uint32_t (iba_global.cpp:18-23).
//
// Restated (not copied) from the published algorithm so that the oracle is
// self-contained and can be rebuilt where /root/reference does not exist:
//   * build: recursive midpoint split on the dimension of largest spread,
//     clamped to the data extent, three-way partition, balanced fallback
//     (divideTree nanoflann.hpp:1039-1096, middleSplit_ :1209-1259,
//      planeSplit :1270-1312, computeBoundingBox :1696-1727);
//   * search: depth-first, nearer child first, per-dimension lower-bound
//     bookkeeping, prune when mindist*(1+eps) > worst (findNeighbors :1588-1610,
//     computeInitialDistances :1314-1336, searchLevel :1736-1811);
//   * result set: sorted insertion with strict '>' (ties keep visit order,
//     a candidate equal to the current worst is rejected) (KNNResultSet
//     :164-237, :1751).
// oracle/_ref builds the same oracle against the REAL vendored header; the
// test-suite checks both give identical (index, distance) lists.
#pragma once
#include <cstdint>
#include <limits>
#include <utility>
#include <vector>

namespace orc {

// KNNResultSet (nanoflann.hpp:164-237)
struct KnnSet {
    uint32_t *indices;
    double *dists;
    size_t capacity, count;
    KnnSet(size_t cap, uint32_t *idx, double *d) : indices(idx), dists(d), capacity(cap), count(0) {
        if (capacity) dists[capacity - 1] = std::numeric_limits<double>::max();
    }
    inline double worst() const { return dists[capacity - 1]; }
    inline void add(double dist, uint32_t index) {
        size_t i;
        for (i = count; i > 0; --i) {
            if (dists[i - 1] > dist) {
                if (i < capacity) { dists[i] = dists[i - 1]; indices[i] = indices[i - 1]; }
            } else {
                break;
            }
        }
        if (i < capacity) { dists[i] = dist; indices[i] = index; }
        if (count < capacity) count++;
    }
};

template <int DIM>
class KDTreePort {
  public:
    // pts: [n][DIM] contiguous fp64 (the adaptor reads m_data[idx][dim])
    KDTreePort(const double *pts, size_t n, int leaf_max_size) : pts_(pts), n_(n), leaf_(leaf_max_size) {
        acc_.resize(n);
        for (size_t i = 0; i < n; ++i) acc_[i] = (uint32_t)i;
        if (n == 0) return;
        for (int d = 0; d < DIM; ++d) root_lo_[d] = root_hi_[d] = at(acc_[0], d);
        for (size_t k = 1; k < n; ++k)
            for (int d = 0; d < DIM; ++d) {
                const double v = at(acc_[k], d);
                if (v < root_lo_[d]) root_lo_[d] = v;
                if (v > root_hi_[d]) root_hi_[d] = v;
            }
        nodes_.reserve(2 * n / (size_t)(leaf_ > 0 ? leaf_ : 1) + 16);
        double lo[DIM], hi[DIM];
        for (int d = 0; d < DIM; ++d) { lo[d] = root_lo_[d]; hi[d] = root_hi_[d]; }
        root_ = divide(0, n, lo, hi);
        for (int d = 0; d < DIM; ++d) { root_lo_[d] = lo[d]; root_hi_[d] = hi[d]; }  // divideTree updates the root bbox in place
    }

    size_t size() const { return n_; }

    // findNeighbors (nanoflann.hpp:1588-1610) with SearchParameters() (eps = 0)
    void find(KnnSet &rs, const double *q) const {
        if (n_ == 0) return;
        const float epsError = 1.0f;
        double dists[DIM];
        double dist = 0;
        for (int d = 0; d < DIM; ++d) {
            dists[d] = 0;
            if (q[d] < root_lo_[d]) { dists[d] = (q[d] - root_lo_[d]) * (q[d] - root_lo_[d]); dist += dists[d]; }
            if (q[d] > root_hi_[d]) { dists[d] = (q[d] - root_hi_[d]) * (q[d] - root_hi_[d]); dist += dists[d]; }
        }
        search(rs, q, root_, dist, dists, epsError);
    }

  private:
    struct Node {
        int32_t child1, child2;  // -1: leaf
        uint32_t left, right;    // leaf range in acc_
        int divfeat;
        double divlow, divhigh;
    };
    const double *pts_;
    size_t n_;
    int leaf_;
    std::vector<uint32_t> acc_;
    std::vector<Node> nodes_;
    int32_t root_ = -1;
    double root_lo_[DIM], root_hi_[DIM];

    inline double at(uint32_t idx, int d) const { return pts_[(size_t)idx * DIM + d]; }

    void minmax(size_t ind, size_t count, int d, double &mn, double &mx) const {
        mn = mx = at(acc_[ind], d);
        for (size_t i = 1; i < count; ++i) {
            const double v = at(acc_[ind + i], d);
            if (v < mn) mn = v;
            if (v > mx) mx = v;
        }
    }

    // planeSplit (nanoflann.hpp:1270-1312)
    void plane_split(size_t ind, size_t count, int cutfeat, double cutval, size_t &lim1, size_t &lim2) {
        size_t left = 0, right = count - 1;
        for (;;) {
            while (left <= right && at(acc_[ind + left], cutfeat) < cutval) ++left;
            while (right && left <= right && at(acc_[ind + right], cutfeat) >= cutval) --right;
            if (left > right || !right) break;
            std::swap(acc_[ind + left], acc_[ind + right]);
            ++left; --right;
        }
        lim1 = left;
        right = count - 1;
        for (;;) {
            while (left <= right && at(acc_[ind + left], cutfeat) <= cutval) ++left;
            while (right && left <= right && at(acc_[ind + right], cutfeat) > cutval) --right;
            if (left > right || !right) break;
            std::swap(acc_[ind + left], acc_[ind + right]);
            ++left; --right;
        }
        lim2 = left;
    }

    // middleSplit_ (nanoflann.hpp:1209-1259)
    void middle_split(size_t ind, size_t count, size_t &index, int &cutfeat, double &cutval, const double *lo, const double *hi) {
        const double EPS = 0.00001;
        double max_span = hi[0] - lo[0];
        for (int d = 1; d < DIM; ++d) { const double span = hi[d] - lo[d]; if (span > max_span) max_span = span; }
        double max_spread = -1;
        cutfeat = 0;
        for (int d = 0; d < DIM; ++d) {
            const double span = hi[d] - lo[d];
            if (span > (1 - EPS) * max_span) {
                double mn, mx;
                minmax(ind, count, d, mn, mx);
                const double spread = mx - mn;
                if (spread > max_spread) { cutfeat = d; max_spread = spread; }
            }
        }
        const double split_val = (lo[cutfeat] + hi[cutfeat]) / 2;
        double mn, mx;
        minmax(ind, count, cutfeat, mn, mx);
        if (split_val < mn) cutval = mn;
        else if (split_val > mx) cutval = mx;
        else cutval = split_val;
        size_t lim1, lim2;
        plane_split(ind, count, cutfeat, cutval, lim1, lim2);
        if (lim1 > count / 2) index = lim1;
        else if (lim2 < count / 2) index = lim2;
        else index = count / 2;
    }

    // divideTree (nanoflann.hpp:1039-1096); lo/hi is the in/out bounding box
    int32_t divide(size_t left, size_t right, double *lo, double *hi) {
        const int32_t me = (int32_t)nodes_.size();
        nodes_.push_back(Node());
        if ((right - left) <= (size_t)leaf_) {
            Node &nd = nodes_[me];
            nd.child1 = nd.child2 = -1;
            nd.left = (uint32_t)left; nd.right = (uint32_t)right;
            for (int d = 0; d < DIM; ++d) lo[d] = hi[d] = at(acc_[left], d);
            for (size_t k = left + 1; k < right; ++k)
                for (int d = 0; d < DIM; ++d) {
                    const double v = at(acc_[k], d);
                    if (lo[d] > v) lo[d] = v;
                    if (hi[d] < v) hi[d] = v;
                }
        } else {
            size_t idx; int cutfeat; double cutval;
            middle_split(left, right - left, idx, cutfeat, cutval, lo, hi);
            double llo[DIM], lhi[DIM], rlo[DIM], rhi[DIM];
            for (int d = 0; d < DIM; ++d) { llo[d] = rlo[d] = lo[d]; lhi[d] = rhi[d] = hi[d]; }
            lhi[cutfeat] = cutval;
            const int32_t c1 = divide(left, left + idx, llo, lhi);
            rlo[cutfeat] = cutval;
            const int32_t c2 = divide(left + idx, right, rlo, rhi);
            Node &nd = nodes_[me];
            nd.child1 = c1; nd.child2 = c2;
            nd.divfeat = cutfeat;
            nd.divlow = lhi[cutfeat];
            nd.divhigh = rlo[cutfeat];
            for (int d = 0; d < DIM; ++d) { lo[d] = llo[d] < rlo[d] ? llo[d] : rlo[d]; hi[d] = lhi[d] > rhi[d] ? lhi[d] : rhi[d]; }
        }
        return me;
    }

    // searchLevel (nanoflann.hpp:1736-1811)
    void search(KnnSet &rs, const double *q, int32_t node, double mindist, double *dists, const float epsError) const {
        const Node &nd = nodes_[node];
        if (nd.child1 < 0 && nd.child2 < 0) {
            const double worst_dist = rs.worst();  // read once per leaf, like the reference
            for (uint32_t i = nd.left; i < nd.right; ++i) {
                const uint32_t a = acc_[i];
                double dist = 0;  // L2_Simple_Adaptor::evalMetric (nanoflann.hpp:524-535)
                for (int d = 0; d < DIM; ++d) { const double diff = q[d] - at(a, d); dist += diff * diff; }
                if (dist < worst_dist) rs.add(dist, a);
            }
            return;
        }
        const int idx = nd.divfeat;
        const double val = q[idx];
        const double diff1 = val - nd.divlow, diff2 = val - nd.divhigh;
        int32_t best, other;
        double cut_dist;
        if ((diff1 + diff2) < 0) { best = nd.child1; other = nd.child2; cut_dist = (val - nd.divhigh) * (val - nd.divhigh); }
        else { best = nd.child2; other = nd.child1; cut_dist = (val - nd.divlow) * (val - nd.divlow); }
        search(rs, q, best, mindist, dists, epsError);
        const double dst = dists[idx];
        mindist = mindist + cut_dist - dst;
        dists[idx] = cut_dist;
        if (mindist * epsError <= rs.worst()) search(rs, q, other, mindist, dists, epsError);
        dists[idx] = dst;
    }
};

}  // namespace orc
