// Exercises the C++ façade (include/ecb/*.h) on the GPU: DBSCAN<T,Float>::Run known answers and status codes.
#include <cstdio>
#include <vector>

#include "../../include/ecb/dbscan.h"
#include "../../include/ecb/event_calib.hpp"

int main() {
    typedef opengv2::Vec2 V2;
    std::vector<V2, ECB_ALLOC(V2)> pts;
    for (int i = 0; i < 10; ++i) pts.push_back(V2{{4.0 * i, 0.0}});
    pts.push_back(V2{{100.0, 100.0}});
    DBSCAN<V2, double> db;
    int rc = db.Run(&pts, 2, 4.0, 2);
    if (rc != 0) {
        std::printf("FAIL rc=%d (%s)\n", rc, db.last_error().c_str());
        return 1;
    }
    // one cluster [1..8], Noise = [0, 9, 10]
    bool ok = db.Clusters.size() == 1 && db.Clusters[0].size() == 8 && db.Noise.size() == 3 && db.Noise[0] == 0 &&
              db.Noise[1] == 9 && db.Noise[2] == 10;
    for (size_t i = 0; ok && i < 8; ++i) ok = db.Clusters[0][i] == i + 1;
    std::vector<V2, ECB_ALLOC(V2)> empty;
    ok = ok && db.Run(&empty, 2, 4.0, 2) == 1 && db.Run(&pts, 2, 4.0, 0) == 1 && db.Run(&pts, 0, 4.0, 2) == 1;
    std::printf(ok ? "facade ok\n" : "FAIL results\n");
    return ok ? 0 : 1;
}
