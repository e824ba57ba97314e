// TEST INFRASTRUCTURE.  Stand-in for core/spline/include/opengv2/spline/BsplineSO3.hpp and the Sophus names the functor
// header mentions.  Sophus is external and absent; the SO(3) functor (CalibReprojectionError_SO3) is NOT instantiated through
// this shim — it would only test the shim's own exp / log — so SO3<T> carries just what the non-template code needs to parse.
#ifndef ECB_ORACLE_BSPLINESO3_SHIM
#define ECB_ORACLE_BSPLINESO3_SHIM
#include <Eigen/Eigen>
#include <vector>
namespace Sophus {
template <class T> using Vector3 = Eigen::Matrix<T, 3, 1>;
template <class T> using Vector4 = Eigen::Matrix<T, 4, 1>;
template <class T>
struct SO3 {
    static constexpr int num_parameters = 4;
    Eigen::Quaternion<T> q;
    const Eigen::Quaternion<T> &unit_quaternion() const { return q; }
};
typedef SO3<double> SO3d;
}  // namespace Sophus
namespace opengv2 {
class BsplineSO3 {
public:
    void evaluate(double, int, Sophus::SO3d &, std::vector<Eigen::Vector3d> &) const {}
};
}  // namespace opengv2
#endif
