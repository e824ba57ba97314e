// TEST INFRASTRUCTURE.  Stand-in for the reference's map object model (not on the hot path): EventCalibSpline only holds a Ptr.
#ifndef ECB_ORACLE_MAPBASE_SHIM
#define ECB_ORACLE_MAPBASE_SHIM
#include <memory>
namespace opengv2 {
struct MapBase {
    typedef std::shared_ptr<MapBase> Ptr;
};
}  // namespace opengv2
#endif
