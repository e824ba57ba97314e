// The slice of the reference's object model that its hot-path signatures name — LandmarkBase (id + position), Bodyframe (time
// stamp + pose + the frame's features) and MapBase (key frames ordered by stamp, landmarks) — with the accessor names the
// reference's EventCalibSpline.cpp / EventCalibIni.cpp use (core/landmark/.../LandmarkBase.hpp:17-60,
// core/frame/.../Bodyframe.hpp, core/map/.../MapBase.hpp:18-110).  The map / tracking / viewer object model itself is out of
// scope; this is the minimum that lets `EventCalibSpline(MapBase::Ptr, EventContainer::Ptr, bool, bool, double,
// double)` and `LandmarkBase::Ptr CirclesEventFrame::findCenter(const Eigen::Vector2d &)` exist with the reference's argument
// lists.  Differences: single-threaded (no shared mutexes), the camera lives in the map (`MapBase::camera`) instead of behind
// `EventContainer::camera`, a frame's features are plain (cx, cy, r) triples in board order.
#ifndef ECB_COMPAT_OPENGV2_LITE_HPP
#define ECB_COMPAT_OPENGV2_LITE_HPP

#include <array>
#include <map>
#include <memory>
#include <vector>

#include "../calib_init.hpp"
#include "eigen_lite.hpp"

namespace opengv2 {

class LandmarkBase {
public:
    typedef std::shared_ptr<LandmarkBase> Ptr;
    LandmarkBase(int id, const Eigen::Vector3d &position) : id_(id), position_(position) {}
    int id() const noexcept { return id_; }
    Eigen::Vector3d position() const noexcept { return position_; }

private:
    int id_;
    Eigen::Vector3d position_;
};

class Bodyframe {
public:
    typedef std::shared_ptr<Bodyframe> Ptr;
    Bodyframe(double timeStamp, const Eigen::Quaterniond &unitQwb, const Eigen::Vector3d &twb) : t_(timeStamp), q_(unitQwb), p_(twb) {}
    double timeStamp() const { return t_; }
    Eigen::Quaterniond unitQwb() const { return q_; }
    Eigen::Vector3d twb() const { return p_; }
    void setPose(const Eigen::Vector3d &twb, const Eigen::Quaterniond &unitQwb) {
        p_ = twb;
        q_ = unitQwb;
    }
    std::vector<std::array<double, 3>> circles;  // the CirclesEventFrame's features after rectifyFeatures: (cx, cy, r), r < 0 absent
    std::vector<double> optT;                    // Qbw (x y z w) + twb, like the reference's swap buffer (EventCalibSpline.cpp:296-316)

private:
    double t_;
    Eigen::Quaterniond q_;
    Eigen::Vector3d p_;
};

class MapBase {
public:
    typedef std::shared_ptr<MapBase> Ptr;
    void addFrame(const Bodyframe::Ptr &bf) { keyframes_[bf->timeStamp()] = bf; }
    void removeFrame(double timestamp) { keyframes_.erase(timestamp); }
    bool hasKeyframe(double timestamp) const { return keyframes_.count(timestamp) != 0; }
    Bodyframe::Ptr keyframe(double timestamp) const {
        const auto it = keyframes_.find(timestamp);
        return it == keyframes_.end() ? nullptr : it->second;
    }
    size_t frameNum() const { return keyframes_.size(); }
    const std::map<double, Bodyframe::Ptr> &keyframes() const { return keyframes_; }
    void addLandmark(const LandmarkBase::Ptr &lm) { landmarks_[lm->id()] = lm; }
    LandmarkBase::Ptr landmark(int id) const {
        const auto it = landmarks_.find(id);
        return it == landmarks_.end() ? nullptr : it->second;
    }
    const std::map<int, LandmarkBase::Ptr> &landmarks() const { return landmarks_; }
    void keyframeLockShared() {}
    void keyframeUnlockShared() {}

    ecb::CameraModel camera;                  // K + distCoeffs (k1 k2 p1 p2 k3) of the initialisation (PinholeCamera)
    std::array<double, 5> inverseRadialPoly{};  // written by EventCalibSpline::updateMap (EventCalibSpline.cpp:262)

private:
    std::map<double, Bodyframe::Ptr> keyframes_;
    std::map<int, LandmarkBase::Ptr> landmarks_;
};

}  // namespace opengv2
#endif  // ECB_COMPAT_OPENGV2_LITE_HPP
