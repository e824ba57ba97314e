// TEST INFRASTRUCTURE.  Stand-in for core/frame/include/opengv2/frame/CameraFrame.hpp: EventFrame only forwards (image, sensor).
#ifndef ECB_ORACLE_CAMERAFRAME_SHIM
#define ECB_ORACLE_CAMERAFRAME_SHIM
#include <opencv2/opencv.hpp>
#include <opengv2/feature/FeatureBase.hpp>
#include <opengv2/sensor/CameraBase.hpp>
#include <vector>
namespace opengv2 {
class CameraFrame {
public:
    CameraFrame(const cv::Mat &image, CameraBase::Ptr sensor) : image_(image), sensor_(sensor) {}
    virtual ~CameraFrame() {}
    std::vector<FeatureBase::Ptr> &features() { return features_; }
    const cv::Mat &image() const { return image_; }
    CameraBase::Ptr sensor() const { return sensor_; }

protected:
    cv::Mat image_;
    CameraBase::Ptr sensor_;
    std::vector<FeatureBase::Ptr> features_;
};
}  // namespace opengv2
#endif
