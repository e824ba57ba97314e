// TEST INFRASTRUCTURE.  Stand-in for core/tracking/include/opengv2/tracking/TrackingBase.hpp.  process() restates the state
// machine of core/tracking/src/TrackingBase.cpp:16-46 (first frame -> initialization(), later frames -> track()) without its
// mutexes; that file also holds triangulation code that is not on the path.
#ifndef ECB_ORACLE_TRACKINGBASE_SHIM
#define ECB_ORACLE_TRACKINGBASE_SHIM
#include <memory>
#include <opengv2/map/MapBase.hpp>
namespace opengv2 {
class SystemBase;
enum TrackingState { NOT_INITIALIZED, OK, LOST };
class TrackingBase {
public:
    typedef std::shared_ptr<TrackingBase> Ptr;
    explicit TrackingBase(MapBase::Ptr map) : state(NOT_INITIALIZED), map_(std::move(map)), system_(nullptr) {}
    virtual ~TrackingBase() {}
    virtual bool process(Bodyframe::Ptr bodyframe) {
        if (state == NOT_INITIALIZED) {
            const bool ok = initialization(bodyframe);
            if (ok) state = OK;
            return ok;
        }
        if (state == OK) return track(bodyframe);
        return false;
    }
    void setSystem(SystemBase *system) { system_ = system; }
    TrackingState state;

protected:
    virtual bool track(Bodyframe::Ptr bodyframe) = 0;
    virtual bool initialization(Bodyframe::Ptr bodyframe) = 0;
    MapBase::Ptr map_;
    SystemBase *system_;
};
}  // namespace opengv2
#endif
