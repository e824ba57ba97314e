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
// the reference's call sequence (CirclesEventFrame.cpp:332-336): plain finder, then the CALIB_CB_CLUSTERING retry;
// returns 1 / 2 = which attempt found the grid, 0 = none.  grid_cluster_select: the selection stage alone.
extern "C" int grid_order_with_retry(const double *xy, int n, int rows, int cols, int *order) {
    std::vector<ecb::Pt2> p((size_t) n);
    for (int i = 0; i < n; ++i) p[(size_t) i] = ecb::Pt2{xy[2 * i], xy[2 * i + 1]};
    std::vector<int> o;
    int which = 1;
    if (!ecb::find_asymmetric_circles_grid(p, rows, cols, o)) {
        which = 2;
        if (!ecb::find_asymmetric_circles_grid_clustering(p, rows, cols, o)) return 0;
    }
    for (size_t i = 0; i < o.size(); ++i) order[i] = o[i];
    return which;
}
extern "C" int grid_order_clustering(const double *xy, int n, int rows, int cols, int *order) {
    std::vector<ecb::Pt2> p((size_t) n);
    for (int i = 0; i < n; ++i) p[(size_t) i] = ecb::Pt2{xy[2 * i], xy[2 * i + 1]};
    std::vector<int> o;
    if (!ecb::find_asymmetric_circles_grid_clustering(p, rows, cols, o)) return 0;
    for (size_t i = 0; i < o.size(); ++i) order[i] = o[i];
    return 1;
}
extern "C" int grid_cluster_select(const double *xy, int n, int pn, int *sel) {
    std::vector<ecb::Pt2> p((size_t) n);
    for (int i = 0; i < n; ++i) p[(size_t) i] = ecb::Pt2{xy[2 * i], xy[2 * i + 1]};
    std::vector<int> s;
    if (!ecb::hierarchical_cluster_select(p, (size_t) pn, s)) return 0;
    for (size_t i = 0; i < s.size(); ++i) sel[i] = s[i];
    return (int) s.size();
}
