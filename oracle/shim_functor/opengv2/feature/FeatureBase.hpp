// TEST INFRASTRUCTURE.  Empty stand-in (the object model is not on the hot path).
#ifndef ECB_ORACLE_FEATUREBASE_SHIM
#define ECB_ORACLE_FEATUREBASE_SHIM
#endif
