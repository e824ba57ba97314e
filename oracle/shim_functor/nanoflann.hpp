// TEST INFRASTRUCTURE.  Stand-in for <nanoflann.hpp> (external, absent): the reference's utility.hpp names two nanoflann
// templates in an alias (SO3_KDTree) that the code compiled in place through this directory never instantiates.
#ifndef ECB_ORACLE_NANOFLANN_SHIM
#define ECB_ORACLE_NANOFLANN_SHIM
namespace nanoflann {
template <class Distance, class DatasetAdaptor, int DIM = -1, typename IndexType = unsigned long>
class KDTreeSingleIndexAdaptor;
template <class T, class DataSource, typename DistanceType = T>
struct SO3_Adaptor;
}  // namespace nanoflann
#endif
