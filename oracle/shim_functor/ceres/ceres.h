// TEST INFRASTRUCTURE: see rotation.h in this directory
#include "rotation.h"
