// TEST INFRASTRUCTURE.  Stand-in for core/frame/include/opengv2/frame/Bodyframe.hpp: time stamp, body pose, one frame, the
// optimisation scratch `optT` and the (identity) sensor extrinsics — the accessors EventCalibSpline.cpp uses.
#ifndef ECB_ORACLE_BODYFRAME_SHIM
#define ECB_ORACLE_BODYFRAME_SHIM
#include <Eigen/Eigen>
#include <memory>
#include <vector>
#include <opengv2/frame/CameraFrame.hpp>
namespace opengv2 {
class Bodyframe {
public:
    typedef std::shared_ptr<Bodyframe> Ptr;
    Bodyframe(std::shared_ptr<CameraFrame> frame, double timeStamp, const Eigen::Vector3d &twb, const Eigen::Quaterniond &Qwb)
        : frame_(frame), ts_(timeStamp), twb_(twb), Qwb_(Qwb) {}
    double timeStamp() const noexcept { return ts_; }
    const Eigen::Vector3d &twb() const noexcept { return twb_; }
    const Eigen::Quaterniond &unitQwb() const noexcept { return Qwb_; }
    void setPose(const Eigen::Vector3d &twb, const Eigen::Quaterniond &Qwb) noexcept {
        twb_ = twb;
        Qwb_ = Qwb;
    }
    std::shared_ptr<CameraFrame> frame(int) const { return frame_; }
    int size() const { return 1; }
    static Eigen::Quaterniond unitQsb(int) { return Eigen::Quaterniond(1, 0, 0, 0); }
    static Eigen::Vector3d tsb(int) { return Eigen::Vector3d(0, 0, 0); }
    std::vector<double> optT;

private:
    std::shared_ptr<CameraFrame> frame_;
    double ts_;
    Eigen::Vector3d twb_;
    Eigen::Quaterniond Qwb_;
};
}  // namespace opengv2
#endif
