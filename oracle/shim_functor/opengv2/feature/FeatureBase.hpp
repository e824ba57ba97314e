// TEST INFRASTRUCTURE.  Stand-in for core/feature/include/opengv2/feature/FeatureBase.hpp: a 2-d location and a weak
// landmark link (same accessors as the reference's class).
#ifndef ECB_ORACLE_FEATUREBASE_SHIM
#define ECB_ORACLE_FEATUREBASE_SHIM
#include <Eigen/Eigen>
#include <memory>
#include <opengv2/landmark/LandmarkBase.hpp>
namespace opengv2 {
class FeatureBase {
public:
    typedef std::shared_ptr<FeatureBase> Ptr;
    explicit FeatureBase(const Eigen::Vector2d &loc) : loc_(loc) {}
    virtual ~FeatureBase() {}
    LandmarkBase::Ptr landmark() const noexcept { return landmark_.lock(); }
    void setLandmark(const LandmarkBase::Ptr &lm) noexcept { landmark_ = lm; }
    const Eigen::Vector2d &location() const noexcept { return loc_; }
    virtual void setLocation(const Eigen::Vector2d &loc) noexcept { loc_ = loc; }

protected:
    std::weak_ptr<LandmarkBase> landmark_;
    Eigen::Vector2d loc_;
};
}  // namespace opengv2
#endif
