// Test helper: host build of include/ecb/circles_grid.hpp
#include "../../include/ecb/circles_grid.hpp"
extern "C" int grid_order(const double *xy, int n, int rows, int cols, int *order) {
    std::vector<ecb::Pt2> p((size_t) n);
    for (int i = 0; i < n; ++i) p[(size_t) i] = ecb::Pt2{xy[2 * i], xy[2 * i + 1]};
    std::vector<int> o;
    if (!ecb::find_asymmetric_circles_grid(p, rows, cols, o)) return 0;
    for (size_t i = 0; i < o.size(); ++i) order[i] = o[i];
    return 1;
}
