// TEST INFRASTRUCTURE — thin C wrapper around the UNMODIFIED reference DBSCAN.
//
// This file is compiled together with the reference's own sources where they
// lie (/root/reference/modules/camera_calibration/dbscan/include/dbscan.h and
// .../dbscan/src/kdtree.cpp) into oracle/_ref/libref_dbscan.so by
// oracle/Makefile.  Nothing from the reference is copied into this repo; the
// only thing added is a ctypes-friendly entry point so tests can compare the
// CUDA path and the CPU restatement (oracle/ecb_oracle.cpp) against the real
// thing.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load the resulting library.
#include <dbscan.h>  // reference header, resolved through -I (see Makefile)

#include <cstring>
#include <vector>

// The same template for other point types / dimensions (dbscan.h:40-41: any T with operator[] convertible to double).
//   pts  n x dim doubles, dim in 1..4
template <int D>
struct PointN {
    double v[D];
    double operator[](int i) const { return v[i]; }
};

template <int D>
static int run_nd(const double *pts, int n, double eps, unsigned minpts, int *labels, int *n_clusters, unsigned *members,
                  unsigned *cluster_off, unsigned *noise, int *n_noise) {
    std::vector<PointN<D>, Eigen::aligned_allocator<PointN<D>>> V((size_t) (n > 0 ? n : 0));
    for (int i = 0; i < n; ++i)
        for (int d = 0; d < D; ++d) V[i].v[d] = pts[(size_t) i * D + d];
    DBSCAN<PointN<D>, double> db;
    int rc = db.Run(&V, D, eps, minpts);
    *n_clusters = 0;
    *n_noise = 0;
    if (rc != 0) return rc;
    for (int i = 0; i < n; ++i) labels[i] = -1;
    unsigned off = 0;
    cluster_off[0] = 0;
    for (size_t c = 0; c < db.Clusters.size(); ++c) {
        for (uint pid : db.Clusters[c]) {
            labels[pid] = (int) c;
            members[off++] = pid;
        }
        cluster_off[c + 1] = off;
    }
    *n_clusters = (int) db.Clusters.size();
    for (size_t i = 0; i < db.Noise.size(); ++i) noise[i] = db.Noise[i];
    *n_noise = (int) db.Noise.size();
    return rc;
}

extern "C" {

// Runs DBSCAN<Eigen::Vector2d,double>::Run exactly as
// CirclesEventFrame::extractFeatures does (CirclesEventFrame.cpp:66-70).
//   xy            n x 2 doubles, pid order
//   labels        out, n ints: cluster id (discovery order) or -1 for Noise
//   members       out, n uints: concatenation of Clusters[c] in reference order
//   cluster_off   out, (n+1) uints: Clusters[c] = members[off[c] .. off[c+1])
//   noise         out, n uints: the reference's Noise vector
// returns the reference's status (0 SUCCESS, 1 FAILED); n_clusters / n_noise
// receive the vector sizes.
int ref_dbscan_run(const double *xy, int n, double eps, unsigned minpts, int *labels, int *n_clusters,
                   unsigned *members, unsigned *cluster_off, unsigned *noise, int *n_noise) {
    std::vector<Eigen::Vector2d, Eigen::aligned_allocator<Eigen::Vector2d>> V((size_t) (n > 0 ? n : 0));
    for (int i = 0; i < n; ++i) V[i] = Eigen::Vector2d(xy[2 * i], xy[2 * i + 1]);
    DBSCAN<Eigen::Vector2d, double> db;
    int rc = db.Run(&V, 2, eps, minpts);
    *n_clusters = 0;
    *n_noise = 0;
    if (rc != 0) return rc;
    for (int i = 0; i < n; ++i) labels[i] = -1;
    unsigned off = 0;
    cluster_off[0] = 0;
    for (size_t c = 0; c < db.Clusters.size(); ++c) {
        for (uint pid : db.Clusters[c]) {
            labels[pid] = (int) c;
            members[off++] = pid;
        }
        cluster_off[c + 1] = off;
    }
    *n_clusters = (int) db.Clusters.size();
    for (size_t i = 0; i < db.Noise.size(); ++i) noise[i] = db.Noise[i];
    *n_noise = (int) db.Noise.size();
    return rc;
}

int ref_dbscan_run_nd(const double *pts, int n, int dim, double eps, unsigned minpts, int *labels, int *n_clusters,
                      unsigned *members, unsigned *cluster_off, unsigned *noise, int *n_noise) {
    switch (dim) {
        case 1: return run_nd<1>(pts, n, eps, minpts, labels, n_clusters, members, cluster_off, noise, n_noise);
        case 2: return run_nd<2>(pts, n, eps, minpts, labels, n_clusters, members, cluster_off, noise, n_noise);
        case 3: return run_nd<3>(pts, n, eps, minpts, labels, n_clusters, members, cluster_off, noise, n_noise);
        case 4: return run_nd<4>(pts, n, eps, minpts, labels, n_clusters, members, cluster_off, noise, n_noise);
    }
    return -1;
}

// Raw kd_nearest_range result order for one query (used to pin the restated
// traversal order): builds the tree by inserting xy in pid order like
// DBSCAN::buildKdtree (dbscan.h:185-196) and returns the hit pids in result-list order.
int ref_kd_range(const double *xy, int n, int q, double eps, unsigned *out) {
    kdtree *t = kd_create(2);
    for (int i = 0; i < n; ++i) kd_insert(t, xy + 2 * i, (void *) (xy + 2 * i));
    kdres *r = kd_nearest_range(t, xy + 2 * q, eps);
    int k = 0;
    while (!kd_res_end(r)) {
        const double *p = (const double *) kd_res_item(r, nullptr);
        out[k++] = (unsigned) ((p - xy) / 2);
        kd_res_next(r);
    }
    kd_res_free(r);
    kd_free(t);
    return k;
}
}
