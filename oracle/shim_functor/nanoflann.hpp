// TEST INFRASTRUCTURE.  Stand-in for <nanoflann.hpp> (external, absent): the names the reference mentions.  The searches
// themselves are restated in KDTreeVectorOfVectorsAdaptor.h of this directory.
#ifndef ECB_ORACLE_NANOFLANN_SHIM
#define ECB_ORACLE_NANOFLANN_SHIM
namespace nanoflann {
struct metric_L2 {};
struct metric_L2_Simple {};
struct SearchParams {
    SearchParams(int = 32, float = 0, bool = true) {}
};
template <class Distance, class DatasetAdaptor, int DIM = -1, typename IndexType = unsigned long>
class KDTreeSingleIndexAdaptor;
template <class T, class DataSource, typename DistanceType = T>
struct SO3_Adaptor;
}  // namespace nanoflann
#endif
