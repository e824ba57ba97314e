// TEST INFRASTRUCTURE.  Stand-in for the reference's map (core/map/include/opengv2/map/MapBase.hpp): key frames ordered by
// time stamp with the accessors EventCalibSpline.cpp uses (the shared-mutex calls are no-ops in this single-threaded wrapper).
#ifndef ECB_ORACLE_MAPBASE_SHIM
#define ECB_ORACLE_MAPBASE_SHIM
#include <map>
#include <memory>
#include <opengv2/frame/Bodyframe.hpp>
#include <opengv2/landmark/LandmarkBase.hpp>
#include <vector>
namespace opengv2 {
class MapBase {
public:
    typedef std::shared_ptr<MapBase> Ptr;
    void addFrame(Bodyframe::Ptr bf) { kf_[bf->timeStamp()] = bf; }
    void removeFrame(double id) { kf_.erase(id); }
    size_t frameNum() const { return kf_.size(); }
    const std::map<double, Bodyframe::Ptr> &keyframes() const { return kf_; }
    Bodyframe::Ptr keyframe(double id) const {
        const auto it = kf_.find(id);
        return it == kf_.end() ? nullptr : it->second;
    }
    bool empty() const { return kf_.empty(); }
    Bodyframe::Ptr firstKeyframe() const { return kf_.empty() ? nullptr : kf_.begin()->second; }
    Bodyframe::Ptr lastKeyframe() const { return kf_.empty() ? nullptr : kf_.rbegin()->second; }
    std::map<double, Bodyframe::Ptr> copyKeyframes() const { return kf_; }
    void addLandmark(LandmarkBase::Ptr lm) { landmarks_.push_back(lm); }
    const std::vector<LandmarkBase::Ptr> &landmarks() const { return landmarks_; }
    void keyframeLockShared() {}
    void keyframeUnlockShared() {}

private:
    std::map<double, Bodyframe::Ptr> kf_;
    std::vector<LandmarkBase::Ptr> landmarks_;
};
}  // namespace opengv2
#endif
