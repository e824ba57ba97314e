// TEST INFRASTRUCTURE.  Stand-in for core/frame/include/opengv2/frame/CameraFrame.hpp: EventFrame only forwards (image, sensor).
#ifndef ECB_ORACLE_CAMERAFRAME_SHIM
#define ECB_ORACLE_CAMERAFRAME_SHIM
#include <opencv2/opencv.hpp>
#include <opengv2/sensor/CameraBase.hpp>
namespace opengv2 {
class CameraFrame {
public:
    CameraFrame(const cv::Mat &image, CameraBase::Ptr sensor) : image_(image), sensor_(sensor) {}
    virtual ~CameraFrame() {}

protected:
    cv::Mat image_;
    CameraBase::Ptr sensor_;
};
}  // namespace opengv2
#endif
