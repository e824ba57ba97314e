// TEST INFRASTRUCTURE.  Stand-in for core/sensor/include/opengv2/sensor/CameraBase.hpp: size() and undistortPoint() are all
// EventFrame.cpp asks of the camera (debug image only).
#ifndef ECB_ORACLE_CAMERABASE_SHIM
#define ECB_ORACLE_CAMERABASE_SHIM
#include <Eigen/Eigen>
#include <memory>
namespace opengv2 {
class CameraBase {
public:
    typedef std::shared_ptr<CameraBase> Ptr;
    explicit CameraBase(const Eigen::Vector2d &size = Eigen::Vector2d(0, 0)) : size_(size) {}
    virtual ~CameraBase() {}
    const Eigen::Vector2d &size() const { return size_; }
    virtual Eigen::Vector2d undistortPoint(const Eigen::Vector2d &p) const { return p; }

protected:
    Eigen::Vector2d size_;
};
}  // namespace opengv2
#endif
