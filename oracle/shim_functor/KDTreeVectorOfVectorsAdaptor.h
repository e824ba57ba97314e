// TEST INFRASTRUCTURE.  Stand-in for nanoflann's KDTreeVectorOfVectorsAdaptor (external, absent), restated as exhaustive
// searches with nanoflann's documented results [external]:
//   query(q, k, idx, d2): the k nearest points in ascending squared L2 distance (d2 = sum over dims of (q_d - p_d)^2 in
//                         dimension order); fewer than k points leave the tail of the outputs untouched.  Ties: lower index
//                         first (nanoflann's own tie order depends on the tree shape — PARITY UNPINNED at that boundary).
//   index->radiusSearch(q, r2, out, params): all points with d2 < r2 (strict, RadiusResultSet::addPoint), here in index order
//                         (the reference asks for unsorted results and only uses them as a set).
#ifndef ECB_ORACLE_KDTREE_ADAPTOR_SHIM
#define ECB_ORACLE_KDTREE_ADAPTOR_SHIM
#include <algorithm>
#include <cstddef>
#include <utility>
#include <vector>

#include <nanoflann.hpp>

template <class VectorOfVectorsType, typename num_t = double, int DIM = -1, class Distance = nanoflann::metric_L2, typename IndexType = size_t>
struct KDTreeVectorOfVectorsAdaptor {
    typedef KDTreeVectorOfVectorsAdaptor self_t;
    struct Index {
        const VectorOfVectorsType *pts;
        int dim;
        num_t dist(const num_t *q, size_t i) const {
            num_t r = num_t();
            for (int d = 0; d < dim; ++d) {
                const num_t diff = q[d] - (*pts)[i][d];
                r += diff * diff;
            }
            return r;
        }
        size_t radiusSearch(const num_t *q, const num_t &radius, std::vector<std::pair<IndexType, num_t>> &out,
                            const nanoflann::SearchParams &) const {
            out.clear();
            for (size_t i = 0; i < pts->size(); ++i) {
                const num_t d = dist(q, i);
                if (d < radius) out.emplace_back((IndexType) i, d);
            }
            return out.size();
        }
    };
    Index *index;
    KDTreeVectorOfVectorsAdaptor(const size_t dimensionality, const VectorOfVectorsType &mat, const int = 10)
        : index(new Index{&mat, (int) dimensionality}) {}
    KDTreeVectorOfVectorsAdaptor(const KDTreeVectorOfVectorsAdaptor &) = delete;
    ~KDTreeVectorOfVectorsAdaptor() { delete index; }
    void query(const num_t *query_point, const size_t num_closest, IndexType *out_indices, num_t *out_distances_sq,
               const int = 10) const {
        std::vector<std::pair<num_t, IndexType>> all;
        for (size_t i = 0; i < index->pts->size(); ++i) all.emplace_back(index->dist(query_point, i), (IndexType) i);
        const size_t k = std::min(num_closest, all.size());
        std::partial_sort(all.begin(), all.begin() + k, all.end());
        for (size_t j = 0; j < k; ++j) {
            out_indices[j] = all[j].second;
            out_distances_sq[j] = all[j].first;
        }
    }
};
#endif
