// TEST INFRASTRUCTURE.  Stand-in for event/include/opengv2/event/EventContainer.hpp: EventCalibSpline only holds a Ptr.
#ifndef ECB_ORACLE_EVENTCONTAINER_SHIM
#define ECB_ORACLE_EVENTCONTAINER_SHIM
#include <memory>
namespace opengv2 {
struct EventContainer {
    typedef std::shared_ptr<EventContainer> Ptr;
};
}  // namespace opengv2
#endif
