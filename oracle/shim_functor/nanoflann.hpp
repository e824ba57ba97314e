// TEST INFRASTRUCTURE.  Stand-in for <nanoflann.hpp> (external, absent): the names the reference mentions.  The searches the
// hot path uses are restated in KDTreeVectorOfVectorsAdaptor.h of this directory; KDTreeSingleIndexAdaptor only appears in the
// reference's disabled reduceMap experiment (EventCalibSpline.cpp:345-549, `reduceMap: 0`) and is a shell that is never run.
#ifndef ECB_ORACLE_NANOFLANN_SHIM
#define ECB_ORACLE_NANOFLANN_SHIM
#include <cstddef>
#include <utility>
#include <vector>
namespace nanoflann {
struct metric_L2 {};
struct metric_L2_Simple {};
struct SearchParams {
    SearchParams(int = 32, float = 0, bool = true) {}
};
struct KDTreeSingleIndexAdaptorParams {
    explicit KDTreeSingleIndexAdaptorParams(size_t = 10) {}
};
template <class T, class DataSource, typename DistanceType = T>
struct SO3_Adaptor;
template <class Distance, class DatasetAdaptor, int DIM = -1, typename IndexType = size_t>
class KDTreeSingleIndexAdaptor {
public:
    KDTreeSingleIndexAdaptor(int, const DatasetAdaptor &, const KDTreeSingleIndexAdaptorParams & = KDTreeSingleIndexAdaptorParams()) {}
    void buildIndex() {}
    size_t radiusSearch(const double *, const double &, std::vector<std::pair<IndexType, double>> &out, const SearchParams &) const {
        out.clear();
        return 0;
    }
};
}  // namespace nanoflann
#endif
