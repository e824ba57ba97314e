// TEST INFRASTRUCTURE: see ../opencv.hpp
#include "../opencv.hpp"
