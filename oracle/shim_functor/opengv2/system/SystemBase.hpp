// TEST INFRASTRUCTURE.  Stand-in for core/system/include/opengv2/system/SystemBase.hpp: the two members EventCalibIni.cpp uses.
#ifndef ECB_ORACLE_SYSTEMBASE_SHIM
#define ECB_ORACLE_SYSTEMBASE_SHIM
#include <memory>
#include <opengv2/map/MapBase.hpp>
namespace opengv2 {
class ViewerBase {
public:
    typedef std::shared_ptr<ViewerBase> Ptr;
    virtual ~ViewerBase() {}
    virtual void updateViewer() {}
};
class SystemBase {
public:
    MapBase::Ptr map;
    ViewerBase::Ptr viewer;
};
}  // namespace opengv2
#endif
