// TEST INFRASTRUCTURE.  Stand-in for the reference's feature class (core/feature): this oracle build only needs a 2-d image
// location and an optional link to a landmark, reachable through the accessor names CirclesEventFrame.cpp uses.
#ifndef ECB_ORACLE_FEATUREBASE_SHIM
#define ECB_ORACLE_FEATUREBASE_SHIM
#include <memory>

#include <Eigen/Eigen>
#include <opengv2/landmark/LandmarkBase.hpp>

namespace opengv2 {
class FeatureBase {
    typedef std::weak_ptr<LandmarkBase> Link;

public:
    typedef std::shared_ptr<FeatureBase> Ptr;

    explicit FeatureBase(const Eigen::Vector2d &xy) : where_(xy) {}
    virtual ~FeatureBase() = default;

    const Eigen::Vector2d &location() const { return where_; }
    virtual void setLocation(const Eigen::Vector2d &xy) { where_ = xy; }

    std::shared_ptr<LandmarkBase> landmark() const { return link_.lock(); }
    void setLandmark(const std::shared_ptr<LandmarkBase> &target) { link_ = target; }

private:
    Eigen::Vector2d where_;
    Link link_;
};
}  // namespace opengv2
#endif
