// TEST INFRASTRUCTURE.  Stand-in for core/sensor/include/opengv2/sensor/PinholeCamera.hpp: K and distCoeffs are only handed to
// the projectPoints hook.
#ifndef ECB_ORACLE_PINHOLECAMERA_SHIM
#define ECB_ORACLE_PINHOLECAMERA_SHIM
#include <opengv2/sensor/CameraBase.hpp>
namespace opengv2 {
class PinholeCamera : public CameraBase {
public:
    explicit PinholeCamera(const Eigen::Vector2d &size) : CameraBase(size) {}
    const Eigen::Matrix3d &K() const { return K_; }
    const Eigen::VectorXd &distCoeffs() const { return dist_; }

private:
    Eigen::Matrix3d K_;
    Eigen::VectorXd dist_;
};
}  // namespace opengv2
#endif
