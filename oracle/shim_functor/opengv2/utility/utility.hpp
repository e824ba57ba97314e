// TEST INFRASTRUCTURE.  Stand-in for core/utility/include/opengv2/utility/utility.hpp: the one alias the functor header uses.
#ifndef ECB_ORACLE_UTILITY_SHIM
#define ECB_ORACLE_UTILITY_SHIM
#include <Eigen/Eigen>
#include <vector>
namespace opengv2 {
template <class M> using vectorofEigenMatrix = std::vector<M, Eigen::aligned_allocator<M>>;
}
#endif
