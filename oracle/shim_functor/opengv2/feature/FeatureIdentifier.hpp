// TEST INFRASTRUCTURE.  Stand-in for core/feature/include/opengv2/feature/FeatureIdentifier.hpp (an observation record).
#ifndef ECB_ORACLE_FEATUREIDENTIFIER_SHIM
#define ECB_ORACLE_FEATUREIDENTIFIER_SHIM
#include <opengv2/feature/FeatureBase.hpp>
namespace opengv2 {
struct FeatureIdentifier {
    FeatureIdentifier(double timeStamp, int frameIdx, int featureIdx, FeatureBase::Ptr f)
        : timeStamp(timeStamp), frameIdx(frameIdx), featureIdx(featureIdx), feature(f) {}
    double timeStamp;
    int frameIdx, featureIdx;
    std::weak_ptr<FeatureBase> feature;
};
}  // namespace opengv2
#endif
