// TEST INFRASTRUCTURE.  Stand-in for core/spline/include/opengv2/spline/BsplineSO3.hpp and the Sophus names the reference
// mentions.  Sophus is external and absent; the SO(3) path (useSO3: 1) is NOT exercised through this shim — it would only test
// the shim's own exp / log — so SO3<T> and BsplineSO3 carry just what the non-template code needs to compile.
#ifndef ECB_ORACLE_BSPLINESO3_SHIM
#define ECB_ORACLE_BSPLINESO3_SHIM
#include <Eigen/Eigen>
#include <ceres/rotation.h>
#include <vector>
namespace Sophus {
template <class T> using Vector3 = Eigen::Matrix<T, 3, 1>;
template <class T> using Vector4 = Eigen::Matrix<T, 4, 1>;
template <class T> struct Constants {
    static T epsilon() { return T(1e-10); }
};
template <class T>
struct SO3 {
    static constexpr int num_parameters = 4;
    Eigen::Quaternion<T> q;
    SO3() : q(T(1.0), T(0.0), T(0.0), T(0.0)) {}
    explicit SO3(const Eigen::Quaternion<T> &q_) : q(q_) {}
    const Eigen::Quaternion<T> &unit_quaternion() const { return q; }
    T *data() { return q.coeffs().data(); }
};
typedef SO3<double> SO3d;
}  // namespace Sophus
namespace opengv2 {
class LocalParameterizationSO3 : public ceres::LocalParameterization {};
class BsplineSO3 {
public:
    typedef std::vector<Sophus::SO3d, Eigen::aligned_allocator<Sophus::SO3d>> VV;
    BsplineSO3(int, const VV &Q, int, const std::vector<double> &) : cp_(Q) {}
    VV &controlPoints() { return cp_; }
    size_t findSpan(double) const { return 3; }
    void derBasisFuns(double, size_t, int, std::vector<std::vector<double>> &b) const { b.assign(1, std::vector<double>(3, 0.0)); }
    void evaluate(double, int, Sophus::SO3d &, std::vector<Eigen::Vector3d> &) const {}

private:
    VV cp_;
};
}  // namespace opengv2
#endif
