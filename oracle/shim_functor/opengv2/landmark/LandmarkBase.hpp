// TEST INFRASTRUCTURE.  Stand-in for core/landmark/include/opengv2/landmark/LandmarkBase.hpp (id + position).
#ifndef ECB_ORACLE_LANDMARKBASE_SHIM
#define ECB_ORACLE_LANDMARKBASE_SHIM
#include <Eigen/Eigen>
#include <memory>
namespace opengv2 {
struct FeatureIdentifier;
class LandmarkBase {
public:
    typedef std::shared_ptr<LandmarkBase> Ptr;
    LandmarkBase(int id, const Eigen::Vector3d &p) : id_(id), position_(p) {}
    Eigen::Vector3d position() const noexcept { return position_; }
    int id() const noexcept { return id_; }
    void addObservation(const FeatureIdentifier &) { ++observations_; }
    int observations() const { return observations_; }

private:
    int id_;
    Eigen::Vector3d position_;
    int observations_ = 0;
};
}  // namespace opengv2
#endif
