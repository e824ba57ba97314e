// TEST INFRASTRUCTURE — CPU restatement of the EventCalib front end.
//
// This is the parity ORACLE for the CUDA path, not product code: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load it.  Every function cites the reference file:line it restates
// (paths relative to /root/reference/modules/camera_calibration/).
//
// Pinning: the DBSCAN restatement below is checked against the UNMODIFIED
// reference compiled in place (oracle/_ref/libref_dbscan.so, see
// ref_dbscan_capi.cpp) as ordered lists in tests/test_oracle_dbscan.py, and the
// toy known-answer of SURVEY.md Appendix E.  The window/dedupe order uses the
// real libstdc++ std::unordered_set with the reference's hash, so it is the
// reference's order by construction (known answers: SURVEY.md Appendix E), and
// it is checked element for element against the reference's own EventFrame.cpp
// compiled in place (oracle/_ref/libref_functor.so, ref_functor_capi.cpp) in
// tests/test_oracle_reference_source.py.
// extractFeatures / fitCircle / rectifyFeatures have no reference test or golden
// vector, and the reference's own build needs OpenCV / Eigen / nanoflann; they
// are pinned instead against the reference's CirclesEventFrame.cpp compiled in
// place with stand-in headers and hooks (oracle/shim_functor/,
// ref_functor_capi.cpp -> oracle/_ref/libref_functor.so): same candidate lists,
// bit-identical features / rectified features / verdicts on raw events
// (tests/test_oracle_reference_source.py).  Still unpinned: the external pieces
// the stand-ins restate (Eigen PartialPivLU, nanoflann tie order, OpenCV).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <limits>
#include <queue>
#include <set>
#include <unordered_set>
#include <array>
#include <vector>

namespace {

struct P2 {
    double x, y;
    bool operator==(const P2 &o) const { return x == o.x && y == o.y; }
    double norm() const { return std::sqrt(x * x + y * y); }  // Eigen norm(): sqrt(sum of squares)
};

// core/utility/include/opengv2/utility/utility.hpp:38-51 (boost hash_combine over std::hash<double>)
struct P2Hash {
    size_t operator()(const P2 &p) const {
        size_t seed = 0;
        const double e[2] = {p.x, p.y};
        for (int i = 0; i < 2; ++i) seed ^= std::hash<double>()(e[i]) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
        return seed;
    }
};

// ---------------------------------------------------------------------------------------------
// DBSCAN restatement: dbscan/include/dbscan.h:115-265 + dbscan/src/kdtree.cpp:106-179,344-365,469-486
// ---------------------------------------------------------------------------------------------
struct KdTree {
    const double *xy;
    int n;
    std::vector<int> left, right, dir;
    explicit KdTree(const double *xy_, int n_) : xy(xy_), n(n_), left(n_, -1), right(n_, -1), dir(n_, 0) {
        // kd_insert in pid order, dbscan.h:185-196; insert_rec kdtree.cpp:106-132
        for (int i = 1; i < n; ++i) {
            int node = 0;
            for (;;) {
                int d = dir[node];
                int &child = (xy[2 * i + d] < xy[2 * node + d]) ? left[node] : right[node];
                if (child < 0) {
                    child = i;
                    dir[i] = (d + 1) % 2;
                    break;
                }
                node = child;
            }
        }
    }
    // find_nearest kdtree.cpp:148-179; hits appended in visit order
    void query(int node, const double *pos, double range, std::vector<int> &visit) const {
        if (node < 0) return;
        double dist_sq = 0;
        for (int i = 0; i < 2; ++i) {
            double d = xy[2 * node + i] - pos[i];
            dist_sq += d * d;
        }
        if (dist_sq <= range * range) visit.push_back(node);
        double dx = pos[dir[node]] - xy[2 * node + dir[node]];
        query(dx <= 0.0 ? left[node] : right[node], pos, range, visit);
        if (std::fabs(dx) < range) query(dx <= 0.0 ? right[node] : left[node], pos, range, visit);
    }
    // regionQuery dbscan.h:198-227: result list is built by head insertion (kdtree.cpp:469-486),
    // so it is the visit order reversed; self excluded (dbscan.h:218)
    std::vector<unsigned> region(int pid, double eps) const {
        std::vector<int> visit;
        if (n > 0) query(0, xy + 2 * pid, eps, visit);
        std::vector<unsigned> out;
        for (auto it = visit.rbegin(); it != visit.rend(); ++it)
            if (*it != pid) out.push_back((unsigned) *it);
        return out;
    }
};

struct DbscanResult {
    std::vector<std::vector<unsigned>> clusters;
    std::vector<unsigned> noise;
};

#ifdef ECB_USE_REF_DBSCAN
}  // namespace
// oracle/_ref build: the clustering step is the UNMODIFIED reference class (compiled in place from
// /root/reference, see oracle/Makefile) so that the "reference" CPU baseline is
// "reference DBSCAN verbatim + restated glue".
#include <dbscan.h>
namespace {
int dbscan_run(const double *xy, int n, double eps, unsigned minpts, DbscanResult &res) {
    std::vector<Eigen::Vector2d, Eigen::aligned_allocator<Eigen::Vector2d>> V((size_t) (n > 0 ? n : 0));
    for (int i = 0; i < n; ++i) V[i] = Eigen::Vector2d(xy[2 * i], xy[2 * i + 1]);
    DBSCAN<Eigen::Vector2d, double> db;
    int rc = db.Run(&V, 2, eps, minpts);
    res.clusters = std::move(db.Clusters);
    res.noise = std::move(db.Noise);
    return rc;
}
#else
int dbscan_run(const double *xy, int n, double eps, unsigned minpts, DbscanResult &res) {
    res.clusters.clear();
    res.noise.clear();
    if (n < 1 || minpts < 1) return 1;  // dbscan.h:121-123 (dim is fixed to 2 here)
    KdTree tree(xy, n);
    std::vector<char> visited(n, 0), assigned(n, 0);
    std::set<unsigned> borderset;
    for (int pid = 0; pid < n; ++pid) {  // dbscan.h:143-162
        borderset.clear();
        if (visited[pid]) continue;
        visited[pid] = 1;
        std::vector<unsigned> nb = tree.region(pid, eps);
        if (nb.size() < minpts) continue;
        unsigned cid = (unsigned) res.clusters.size();
        res.clusters.emplace_back();
        borderset.insert((unsigned) pid);
        res.clusters[cid].push_back((unsigned) pid);
        assigned[pid] = 1;
        // expandCluster dbscan.h:229-259
        std::queue<unsigned> border;
        for (unsigned p : nb) border.push(p);
        for (unsigned p : nb) borderset.insert(p);
        while (!border.empty()) {
            unsigned p = border.front();
            border.pop();
            if (visited[p]) continue;
            visited[p] = 1;
            std::vector<unsigned> pn = tree.region((int) p, eps);
            if (pn.size() >= minpts) {
                res.clusters[cid].push_back(p);
                assigned[p] = 1;
                for (unsigned q : pn)
                    if (!borderset.count(q)) {
                        border.push(q);
                        borderset.insert(q);
                    }
            }
        }
    }
    for (int pid = 0; pid < n; ++pid)
        if (!assigned[pid]) res.noise.push_back((unsigned) pid);  // dbscan.h:164-168
    return 0;
}
#endif

// ---------------------------------------------------------------------------------------------
// fitCircle: event_camera_calib/src/CirclesEventFrame.cpp:361-415
// 3x3 solve follows Eigen's PartialPivLU (unblocked, first-max pivot) [external, not in /root/reference]
// ---------------------------------------------------------------------------------------------
void lu_solve3(double A[3][3], const double b_in[3], double x[3]) {
    int perm[3] = {0, 1, 2};
    for (int k = 0; k < 3; ++k) {
        int piv = k;
        double big = std::fabs(A[k][k]);
        for (int i = k + 1; i < 3; ++i)
            if (std::fabs(A[i][k]) > big) {
                big = std::fabs(A[i][k]);
                piv = i;
            }
        if (big != 0.0) {
            if (piv != k) {
                for (int j = 0; j < 3; ++j) std::swap(A[k][j], A[piv][j]);
                std::swap(perm[k], perm[piv]);
            }
            for (int i = k + 1; i < 3; ++i) A[i][k] /= A[k][k];
        }
        for (int i = k + 1; i < 3; ++i)
            for (int j = k + 1; j < 3; ++j) A[i][j] -= A[i][k] * A[k][j];
    }
    double y[3] = {b_in[perm[0]], b_in[perm[1]], b_in[perm[2]]};
    y[1] -= A[1][0] * y[0];
    y[2] -= (A[2][0] * y[0] + A[2][1] * y[1]);
    y[2] /= A[2][2];
    y[1] -= A[1][2] * y[2];
    y[1] /= A[1][1];
    y[0] -= (A[0][1] * y[1] + A[0][2] * y[2]);
    y[0] /= A[0][0];
    x[0] = y[0];
    x[1] = y[1];
    x[2] = y[2];
}

void fit_circle(const std::vector<P2> &pos, const std::vector<P2> &neg, const std::vector<unsigned> &pSet,
                const std::vector<unsigned> &nSet, double center[2], double &radius) {
    double sum_x = 0, sum_y = 0, sum_xx = 0, sum_yy = 0, sum_xy = 0;
    double sum_xxx = 0, sum_yyy = 0, sum_xyy = 0, sum_xxy = 0;
    auto acc = [&](const P2 &s) {
        sum_x += s.x;
        sum_y += s.y;
        double xx = s.x * s.x, yy = s.y * s.y, xy = s.x * s.y;
        sum_xx += xx;
        sum_yy += yy;
        sum_xy += xy;
        sum_xxx += xx * s.x;
        sum_yyy += yy * s.y;
        sum_xyy += xy * s.y;
        sum_xxy += s.x * xy;
    };
    for (unsigned i : pSet) acc(pos[i]);
    for (unsigned i : nSet) acc(neg[i]);
    double A[3][3] = {{2 * sum_x, 2 * sum_y, (double) (pSet.size() + nSet.size())},
                      {2 * sum_xx, 2 * sum_xy, sum_x},
                      {2 * sum_xy, 2 * sum_yy, sum_y}};
    double b[3] = {sum_xx + sum_yy, sum_xxx + sum_xyy, sum_xxy + sum_yyy};
    double x[3];
    lu_solve3(A, b, x);
    center[0] = x[0];
    center[1] = x[1];
    radius = std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2]);
}

// k nearest (squared L2) by brute force, ascending distance, ties by index (nanoflann tie order is
// unpinned — SURVEY.md §8c).  Replaces KDTreeVectorOfVectorsAdaptor::query, CirclesEventFrame.cpp:160-168.
void knn(const std::vector<P2> &pts, const P2 &q, size_t k, std::vector<size_t> &idx, std::vector<double> &d2) {
    std::vector<std::pair<double, size_t>> all(pts.size());
    for (size_t i = 0; i < pts.size(); ++i) {
        double dx = q.x - pts[i].x, dy = q.y - pts[i].y;
        all[i] = {dx * dx + dy * dy, i};
    }
    std::stable_sort(all.begin(), all.end());
    for (size_t i = 0; i < k && i < all.size(); ++i) {
        idx[i] = all[i].second;
        d2[i] = all[i].first;
    }
}

double fit_err(const std::vector<P2> &pos, const std::vector<P2> &neg, const std::vector<unsigned> &pc,
               const std::vector<unsigned> &nc, const double c[2], double r) {
    double e = 0;
    for (unsigned i : pc) e += std::abs(std::sqrt((pos[i].x - c[0]) * (pos[i].x - c[0]) + (pos[i].y - c[1]) * (pos[i].y - c[1])) - r);
    for (unsigned i : nc) e += std::abs(std::sqrt((neg[i].x - c[0]) * (neg[i].x - c[0]) + (neg[i].y - c[1]) * (neg[i].y - c[1])) - r);
    e /= (pc.size() + nc.size()) * r;
    return e;
}

struct Candidate {
    unsigned pi, ni;
    double cx, cy, r;
};

struct FrameResult {
    std::vector<P2> pos, neg;
    DbscanResult pdb, ndb;                                 // raw DBSCAN output
    std::vector<std::vector<unsigned>> pcl, ncl;           // after the clusterMinSample filter (pClusters_/nClusters_)
    std::vector<unsigned> pmed, nmed;                      // median-by-norm member of each kept cluster
    std::vector<Candidate> cand;
    int enough = 0;
};

// CirclesEventFrame::extractFeatures up to (not including) findCirclesGrid: CirclesEventFrame.cpp:61-312
void extract(FrameResult &f, double eps, unsigned minS, unsigned clusterMin, int knn_num, int fitCircleFlag,
             double Rthr, unsigned rows_cols, bool canonical_median) {
    f.cand.clear();
    f.pcl.clear();
    f.ncl.clear();
    f.pmed.clear();
    f.nmed.clear();
    f.enough = 0;
    f.pdb.clusters.clear();
    f.pdb.noise.clear();
    f.ndb.clusters.clear();
    f.ndb.noise.clear();
    if (f.pos.empty() || f.neg.empty()) return;  // :62-64
    dbscan_run(&f.pos[0].x, (int) f.pos.size(), eps, minS, f.pdb);
    dbscan_run(&f.neg[0].x, (int) f.neg.size(), eps, minS, f.ndb);
    for (auto &c : f.pdb.clusters)
        if (c.size() >= clusterMin) f.pcl.push_back(c);  // :89-117
    for (auto &c : f.ndb.clusters)
        if (c.size() >= clusterMin) f.ncl.push_back(c);
    if (f.pcl.size() < rows_cols || f.ncl.size() < rows_cols) return;  // :127-129
    f.enough = 1;
    // medians :137-147 (std::nth_element permutes the member lists in place, like the reference).
    // canonical_median: order-independent variant used to check the CUDA path's documented tie rule
    // (median slot of the members sorted by (norm^2, pid)).
    auto p_clusters = f.pcl, n_clusters = f.ncl;
    auto med = [&](std::vector<std::vector<unsigned>> &cl, const std::vector<P2> &pts, std::vector<unsigned> &out) {
        for (auto &c : cl) {
            if (canonical_median) {
                std::sort(c.begin(), c.end(), [&](unsigned a, unsigned b) {
                    double na = pts[a].x * pts[a].x + pts[a].y * pts[a].y, nb = pts[b].x * pts[b].x + pts[b].y * pts[b].y;
                    return na < nb || (na == nb && a < b);
                });
            } else {
                std::nth_element(c.begin(), c.begin() + c.size() / 2, c.end(),
                                 [&](unsigned a, unsigned b) { return pts[a].norm() < pts[b].norm(); });
            }
            out.push_back(c[c.size() / 2]);
        }
    };
    med(p_clusters, f.pos, f.pmed);
    med(n_clusters, f.neg, f.nmed);
    std::vector<P2> pC(f.pmed.size()), nC(f.nmed.size());
    for (size_t i = 0; i < pC.size(); ++i) pC[i] = f.pos[f.pmed[i]];
    for (size_t i = 0; i < nC.size(); ++i) nC[i] = f.neg[f.nmed[i]];

    const size_t K = fitCircleFlag ? (size_t) knn_num : 1;
    std::vector<size_t> n_idx(K), p_idx(K);
    std::vector<double> d2(K), radius(K), fitErrs;
    std::vector<std::array<double, 2>> centers(K);
    if (fitCircleFlag) {  // :180-281
        for (size_t pi = 0; pi < pC.size(); ++pi) {
            size_t real = K;
            knn(nC, pC[pi], K, n_idx, d2);
            for (size_t oi = 0; oi < d2.size(); ++oi)
                if (d2[oi] > d2[0] * 4 || d2[oi] > 4 * Rthr * Rthr) {
                    real = oi;
                    break;
                }
            if (real == 0) continue;
            fitErrs.assign(real, 0);
            for (size_t j = 0; j < real; ++j) {
                fit_circle(f.pos, f.neg, p_clusters[pi], n_clusters[n_idx[j]], centers[j].data(), radius[j]);
                double dx = pC[pi].x - nC[n_idx[j]].x, dy = pC[pi].y - nC[n_idx[j]].y;
                double approx = std::sqrt(dx * dx + dy * dy) / 2;
                if (radius[j] > Rthr || radius[j] > 2 * approx)
                    fitErrs[j] = std::numeric_limits<double>::max();
                else
                    fitErrs[j] = fit_err(f.pos, f.neg, p_clusters[pi], n_clusters[n_idx[j]], centers[j].data(), radius[j]);
            }
            size_t n_min = std::min_element(fitErrs.begin(), fitErrs.end()) - fitErrs.begin();
            const double gaussianNoise = 2 / radius[n_min];
            if (!(fitErrs[n_min] < gaussianNoise)) continue;
            real = K;
            knn(pC, nC[n_idx[n_min]], K, p_idx, d2);
            for (size_t oi = 0; oi < d2.size(); ++oi)
                if (d2[oi] > d2[0] * 4 || d2[oi] > 4 * Rthr * Rthr) {
                    real = oi;
                    break;
                }
            if (real == 0) continue;
            fitErrs.assign(real, 0);
            for (size_t i = 0; i < real; ++i) {
                fit_circle(f.pos, f.neg, p_clusters[p_idx[i]], n_clusters[n_idx[n_min]], centers[i].data(), radius[i]);
                double dx = pC[p_idx[i]].x - nC[n_idx[n_min]].x, dy = pC[p_idx[i]].y - nC[n_idx[n_min]].y;
                double approx = std::sqrt(dx * dx + dy * dy) / 2;
                if (radius[i] > Rthr || radius[i] > 2 * approx)
                    fitErrs[i] = std::numeric_limits<double>::max();
                else
                    fitErrs[i] = fit_err(f.pos, f.neg, p_clusters[p_idx[i]], n_clusters[n_idx[n_min]], centers[i].data(), radius[i]);
            }
            size_t p_min = std::min_element(fitErrs.begin(), fitErrs.end()) - fitErrs.begin();
            if (p_idx[p_min] == pi)
                f.cand.push_back({(unsigned) pi, (unsigned) n_idx[n_min], centers[p_min][0], centers[p_min][1], radius[p_min]});
        }
    } else {  // :282-312
        for (size_t pi = 0; pi < pC.size(); ++pi) {
            knn(nC, pC[pi], K, n_idx, d2);
            if (d2[0] > 4 * Rthr * Rthr) continue;
            knn(pC, nC[n_idx[0]], K, p_idx, d2);
            if (p_idx[0] != pi) continue;
            double c[2] = {(pC[pi].x + nC[n_idx[0]].x) / 2, (pC[pi].y + nC[n_idx[0]].y) / 2};
            double dx = pC[pi].x - nC[n_idx[0]].x, dy = pC[pi].y - nC[n_idx[0]].y;
            double r = std::sqrt(dx * dx + dy * dy) / 2;
            double e = fit_err(f.pos, f.neg, p_clusters[pi], n_clusters[n_idx[0]], c, r);
            if (e < 10 / r) f.cand.push_back({(unsigned) pi, (unsigned) n_idx[0], c[0], c[1], r});
        }
    }
}

// EventFrame ctor: event/src/EventFrame.cpp:10-36.  Events must be time-sorted (the reference's
// multimap orders them; equal keys keep file order).  Window is CLOSED: [lower_bound(first), upper_bound(second)).
void event_frame(const double *t, const double *x, const double *y, const uint8_t *pol, int64_t n, double t0,
                 double t1, std::vector<P2> &pos, std::vector<P2> &neg, int64_t *lo_out, int64_t *hi_out) {
    int64_t lo = std::lower_bound(t, t + n, t0) - t;
    int64_t hi = std::upper_bound(t, t + n, t1) - t;
    if (lo_out) *lo_out = lo;
    if (hi_out) *hi_out = hi;
    std::unordered_set<P2, P2Hash> P, N;
    for (int64_t i = lo; i < hi; ++i) (pol[i] ? P : N).insert(P2{x[i], y[i]});
    for (auto it = P.begin(); it != P.end();) {
        auto f = N.find(*it);
        if (f == N.end())
            ++it;
        else {
            N.erase(f);
            it = P.erase(it);
        }
    }
    pos.assign(P.begin(), P.end());
    neg.assign(N.begin(), N.end());
}

double radius_threshold(double W, double H, int rows, int cols, int asym, double square, double radius) {
    // CirclesEventFrame ctor, CirclesEventFrame.cpp:16-33 (camera->size() is a double vector)
    double c2 = asym ? 2.0 * cols : (double) cols;
    return std::min(std::max(W, H) / std::max((double) rows, c2), std::min(W, H) / std::min((double) rows, c2)) /
           square * radius * 1.5;
}

}  // namespace

extern "C" {

// ---- DBSCAN (ordered) ----
int orc_dbscan_run(const double *xy, int n, double eps, unsigned minpts, int *labels, int *n_clusters,
                   unsigned *members, unsigned *cluster_off, unsigned *noise, int *n_noise) {
    DbscanResult r;
    *n_clusters = 0;
    *n_noise = 0;
    int rc = dbscan_run(xy, n, eps, minpts, r);
    if (rc) return rc;
    for (int i = 0; i < n; ++i) labels[i] = -1;
    unsigned off = 0;
    cluster_off[0] = 0;
    for (size_t c = 0; c < r.clusters.size(); ++c) {
        for (unsigned p : r.clusters[c]) {
            labels[p] = (int) c;
            members[off++] = p;
        }
        cluster_off[c + 1] = off;
    }
    *n_clusters = (int) r.clusters.size();
    for (size_t i = 0; i < r.noise.size(); ++i) noise[i] = r.noise[i];
    *n_noise = (int) r.noise.size();
    return 0;
}

int orc_kd_range(const double *xy, int n, int q, double eps, unsigned *out) {
    KdTree t(xy, n);
    auto r = t.region(q, eps);
    // region() drops self; the raw kd result keeps it — re-insert at its visit position for comparison
    std::vector<int> visit;
    t.query(0, xy + 2 * q, eps, visit);
    int k = 0;
    for (auto it = visit.rbegin(); it != visit.rend(); ++it) out[k++] = (unsigned) *it;
    return k;
}

// kd tie flags as the CUDA path defines them: bit d set iff an ancestor A with A.dir==d has A[d]==p[d]
void orc_kd_flags(const double *xy, int n, uint8_t *flags) {
    KdTree t(xy, n);
    for (int i = 0; i < n; ++i) flags[i] = 0;
    for (int i = 1; i < n; ++i) {
        int node = 0;
        while (node != i) {
            int d = t.dir[node];
            if (xy[2 * i + d] == xy[2 * node + d]) flags[i] |= (uint8_t) (1 << d);
            node = (xy[2 * i + d] < xy[2 * node + d]) ? t.left[node] : t.right[node];
        }
    }
}

// ---- hashing known-answers ----
uint64_t orc_hash_double(double v) { return (uint64_t) std::hash<double>()(v); }
uint64_t orc_hash_p2(double x, double y) { return (uint64_t) P2Hash()(P2{x, y}); }

// ---- window + dedupe + cancel ----
// out_pos / out_neg: capacity (hi-lo) points each (x,y interleaved). Returns 0.
int orc_event_frame(const double *t, const double *x, const double *y, const uint8_t *pol, int64_t n, double t0,
                    double t1, double *out_pos, int *n_pos, double *out_neg, int *n_neg, int64_t *lo, int64_t *hi) {
    std::vector<P2> P, N;
    event_frame(t, x, y, pol, n, t0, t1, P, N, lo, hi);
    *n_pos = (int) P.size();
    *n_neg = (int) N.size();
    if (out_pos) memcpy(out_pos, P.data(), P.size() * sizeof(P2));
    if (out_neg) memcpy(out_neg, N.data(), N.size() * sizeof(P2));
    return 0;
}

// insertion-order -> iteration-order permutation of std::unordered_set<P2,P2Hash> for a list of
// DISTINCT points (used to pin the GPU emulation of the libstdc++ order).
int orc_uset_order(const double *xy, int n, double *out_xy) {
    std::unordered_set<P2, P2Hash> S;
    for (int i = 0; i < n; ++i) S.insert(P2{xy[2 * i], xy[2 * i + 1]});
    int k = 0;
    for (const P2 &p : S) {
        out_xy[2 * k] = p.x;
        out_xy[2 * k + 1] = p.y;
        ++k;
    }
    return k;
}

double orc_radius_threshold(double W, double H, int rows, int cols, int asym, double square, double radius) {
    return radius_threshold(W, H, rows, cols, asym, square, radius);
}

void orc_fit_circle(const double *pxy, int np, const double *nxy, int nn, double *out3) {
    std::vector<P2> P(np), N(nn);
    std::vector<unsigned> pi(np), ni(nn);
    for (int i = 0; i < np; ++i) P[i] = {pxy[2 * i], pxy[2 * i + 1]}, pi[i] = i;
    for (int i = 0; i < nn; ++i) N[i] = {nxy[2 * i], nxy[2 * i + 1]}, ni[i] = i;
    fit_circle(P, N, pi, ni, out3, out3[2]);
}

// ---- full per-frame front end from explicit point sets (V order given) ----
// Outputs (caller allocates; caps given):
//   p_labels[np], n_labels[nn]      raw DBSCAN labels (-1 noise)
//   p_members/p_off, n_members/n_off reference-ordered Clusters (raw)
//   kept_p[<=np], kept_n             raw cluster id of each kept cluster (size filter)
//   med_p, med_n                     median member (pid) of each kept cluster
//   cand[5*cap]                      (pi, ni, cx, cy, r) per candidate, pi/ni index KEPT clusters
// info[0]=n raw p clusters, [1]=n raw n clusters, [2]=kept p, [3]=kept n, [4]=enough, [5]=n candidates
int orc_extract(const double *pxy, int np, const double *nxy, int nn, double eps, unsigned minS, unsigned clusterMin,
                int knn_num, int fitCircleFlag, double Rthr, unsigned rows_cols, int canonical_median, int *p_labels,
                int *n_labels, unsigned *p_members, unsigned *p_off, unsigned *n_members, unsigned *n_off,
                int *kept_p, int *kept_n, int *med_p, int *med_n, double *cand, int cand_cap, int *info) {
    FrameResult f;
    f.pos.resize(np);
    f.neg.resize(nn);
    for (int i = 0; i < np; ++i) f.pos[i] = {pxy[2 * i], pxy[2 * i + 1]};
    for (int i = 0; i < nn; ++i) f.neg[i] = {nxy[2 * i], nxy[2 * i + 1]};
    extract(f, eps, minS, clusterMin, knn_num, fitCircleFlag, Rthr, rows_cols, canonical_median != 0);
    auto dump = [](const DbscanResult &r, int n, int *labels, unsigned *members, unsigned *off) {
        for (int i = 0; i < n; ++i) labels[i] = -1;
        unsigned o = 0;
        off[0] = 0;
        for (size_t c = 0; c < r.clusters.size(); ++c) {
            for (unsigned p : r.clusters[c]) {
                labels[p] = (int) c;
                members[o++] = p;
            }
            off[c + 1] = o;
        }
    };
    dump(f.pdb, np, p_labels, p_members, p_off);
    dump(f.ndb, nn, n_labels, n_members, n_off);
    auto keptmap = [](const DbscanResult &r, unsigned clusterMin, int *kept) {
        int k = 0;
        for (size_t c = 0; c < r.clusters.size(); ++c)
            if (r.clusters[c].size() >= clusterMin) kept[k++] = (int) c;
        return k;
    };
    info[0] = (int) f.pdb.clusters.size();
    info[1] = (int) f.ndb.clusters.size();
    info[2] = keptmap(f.pdb, clusterMin, kept_p);
    info[3] = keptmap(f.ndb, clusterMin, kept_n);
    info[4] = f.enough;
    for (size_t i = 0; i < f.pmed.size(); ++i) med_p[i] = (int) f.pmed[i];
    for (size_t i = 0; i < f.nmed.size(); ++i) med_n[i] = (int) f.nmed[i];
    int nc = 0;
    for (const Candidate &c : f.cand) {
        if (nc >= cand_cap) break;
        cand[5 * nc + 0] = c.pi;
        cand[5 * nc + 1] = c.ni;
        cand[5 * nc + 2] = c.cx;
        cand[5 * nc + 3] = c.cy;
        cand[5 * nc + 4] = c.r;
        ++nc;
    }
    info[5] = nc;
    return 0;
}

}  // extern "C"

// ---- rectifyFeatures ----
// CirclesEventFrame::rectifyFeatures, CirclesEventFrame.cpp:417-609, from the projected image points on (cv::projectPoints
// at :449 is the caller's, like in the reference it needs the OpenCV initialisation): img = n_feat x 5 x 2 doubles
// (centre, then the four quadrant points of :434-440).  out = n_feat x 3 (cx, cy, r); r < 0: feature deleted.
// Returns the frame verdict (:596-609): 1 keep, 0 drop.  std::pow(x, 2) is evaluated as x*x.
static int rectify(const FrameResult &f, const double *img, int n_feat, double W, double H, int rows, int cols, int asym,
                   int fitCircleFlag, double *out) {
    const double inlierThreshold = 3;
    std::vector<int> pS2S(f.pos.size(), -1), nS2S(f.neg.size(), -1);  // :523-534 (each sample is in at most one cluster)
    for (size_t i = 0; i < f.pcl.size(); ++i)
        for (unsigned j : f.pcl[i]) pS2S[j] = (int) i;
    for (size_t i = 0; i < f.ncl.size(); ++i)
        for (unsigned j : f.ncl[i]) nS2S[j] = (int) i;
    std::vector<char> gone((size_t) n_feat, 0);
    for (int k = 0; k < n_feat; ++k) {
        const double *ip = img + (size_t) k * 10;
        double *o = out + 3 * (size_t) k;
        o[0] = o[1] = 0;
        o[2] = -1;
        gone[(size_t) k] = 1;
        if (ip[0] >= W || ip[1] >= H || ip[0] < 0 || ip[1] < 0) continue;  // :459-463
        double radius[4], maxRadius = 0;
        for (int i = 1; i < 5; ++i) {  // :465-473
            const double dx = ip[2 * i] - ip[0], dy = ip[2 * i + 1] - ip[1];
            radius[i - 1] = std::sqrt(dx * dx + dy * dy);
            if (radius[i - 1] > maxRadius) maxRadius = radius[i - 1];
        }
        const double R2 = (maxRadius + inlierThreshold) * (maxRadius + inlierThreshold);
        std::set<unsigned> pSet, nSet;
        auto collect = [&](const std::vector<P2> &pts, const std::vector<int> &s2s, std::set<unsigned> &sets) {
            for (size_t i = 0; i < pts.size(); ++i) {  // radiusSearch: d2 < R2 (nanoflann RadiusResultSet) :476-481
                const double dx = pts[i].x - ip[0], dy = pts[i].y - ip[1];
                const double d2 = dx * dx + dy * dy;
                if (!(d2 < R2)) continue;
                const double distance = std::sqrt(d2);
                int idx = 0;  // :487-497
                if (dx >= 0 && dy >= 0) idx = 0;
                else if (dx >= 0 && dy <= 0) idx = 1;
                else if (dx <= 0 && dy <= 0) idx = 2;
                else if (dx <= 0 && dy >= 0) idx = 3;
                if (std::abs(distance - radius[idx]) <= inlierThreshold && s2s[i] >= 0) sets.insert((unsigned) s2s[i]);  // :499-501,536-543
            }
        };
        collect(f.pos, pS2S, pSet);
        collect(f.neg, nS2S, nSet);
        std::vector<unsigned> pAll, nAll;  // :544-555
        for (unsigned c : pSet) pAll.insert(pAll.end(), f.pcl[c].begin(), f.pcl[c].end());
        for (unsigned c : nSet) nAll.insert(nAll.end(), f.ncl[c].begin(), f.ncl[c].end());
        if (pAll.size() < 5 || nAll.size() < 5) continue;  // :558-561
        double c[2], r;
        fit_circle(f.pos, f.neg, pAll, nAll, c, r);  // :563-566
        std::nth_element(radius, radius + 2, radius + 4);  // :568
        const double ex = c[0] - ip[0], ey = c[1] - ip[1];
        if (std::sqrt(ex * ex + ey * ey) > 2 * inlierThreshold || std::abs(r - radius[2]) > 1.5 * inlierThreshold) continue;  // :570-574
        o[0] = c[0];
        o[1] = c[1];
        o[2] = r;
        gone[(size_t) k] = 0;
    }
    // edge scores and the 20 % rule :585-609
    int score[4] = {0, 0, 0, 0}, size_[4] = {0, 0, 0, 0}, counter = 0;
    auto on_edge = [&](int e, int i) -> bool {
        const int step = (asym ? 2 : 1) * cols;
        if (e == 0) return i < cols;
        if (e == 1) return i >= (rows - 1) * cols && i < rows * cols;
        if (e == 2) return i < rows * cols && i % step == 0;
        const int first = asym ? 2 * cols - 1 : cols - 1;
        return i >= first && i < rows * cols && (i - first) % step == 0;
    };
    for (int e = 0; e < 4; ++e)
        for (int i = 0; i < rows * cols; ++i) size_[e] += on_edge(e, i);
    for (int i = 0; i < n_feat; ++i)
        if (gone[(size_t) i]) {
            for (int e = 0; e < 4; ++e) score[e] += on_edge(e, i);
            ++counter;
        }
    if (!fitCircleFlag)
        for (int e = 0; e < 4; ++e)
            if (score[e] >= size_[e] - 1) return 0;
    if (counter >= 0.2 * (cols * rows)) return 0;
    return 1;
}

extern "C" int orc_rectify(const double *pxy, int np, const double *nxy, int nn, double eps, unsigned minS, unsigned clusterMin,
                           const double *img, int n_feat, double W, double H, int rows, int cols, int asym, int fitCircleFlag,
                           double *out) {
    FrameResult f;
    f.pos.resize(np);
    f.neg.resize(nn);
    for (int i = 0; i < np; ++i) f.pos[i] = {pxy[2 * i], pxy[2 * i + 1]};
    for (int i = 0; i < nn; ++i) f.neg[i] = {nxy[2 * i], nxy[2 * i + 1]};
    extract(f, eps, minS, clusterMin, 1, 0, 1e30, 0xFFFFFFFFu, true);  // DBSCAN + the clusterMinSample filter only
    return rectify(f, img, n_feat, W, H, rows, cols, asym, fitCircleFlag, out);
}

extern "C" {
// ---- CPU baseline: the reference-shaped front end over a list of windows, `threads` std::threads ----
// (event/src/EventFrame.cpp:10-36 + CirclesEventFrame.cpp:61-312, worker model of
//  event_camera_calib/test/eventCameraCalib.cpp:172-190).  Returns total candidates found; n_events_out =
//  sum of window event counts (the unit of the events/s metric).
}  // extern "C"

#include <atomic>
#include <thread>

// Same loop with the per-window results kept (bench.py's full-size parity check): counts[w] = points -, points +, raw
// clusters -, +, kept clusters -, + ; cand_out[w][k] = pi, ni, cx, cy, r of the first cand_cap candidates.
static int64_t frontend_windows_impl(const double *t, const double *x, const double *y, const uint8_t *pol,
                                     int64_t n, const double *win, int n_win, double eps, unsigned minS,
                                     unsigned clusterMin, int knn_num, int fitCircleFlag, double Rthr,
                                     unsigned rows_cols, int threads, int64_t *n_events_out, int *cand_per_win,
                                     int *counts, double *cand_out, int cand_cap);

extern "C" int64_t orc_frontend_windows_detail(const double *t, const double *x, const double *y, const uint8_t *pol,
                                               int64_t n, const double *win, int n_win, double eps, unsigned minS,
                                               unsigned clusterMin, int knn_num, int fitCircleFlag, double Rthr,
                                               unsigned rows_cols, int threads, int64_t *n_events_out, int *cand_per_win,
                                               int *counts, double *cand_out, int cand_cap) {
    return frontend_windows_impl(t, x, y, pol, n, win, n_win, eps, minS, clusterMin, knn_num, fitCircleFlag, Rthr, rows_cols,
                                 threads, n_events_out, cand_per_win, counts, cand_out, cand_cap);
}

extern "C" int64_t orc_frontend_windows(const double *t, const double *x, const double *y, const uint8_t *pol,
                                        int64_t n, const double *win, int n_win, double eps, unsigned minS,
                                        unsigned clusterMin, int knn_num, int fitCircleFlag, double Rthr,
                                        unsigned rows_cols, int threads, int64_t *n_events_out, int *cand_per_win) {
    return frontend_windows_impl(t, x, y, pol, n, win, n_win, eps, minS, clusterMin, knn_num, fitCircleFlag, Rthr, rows_cols,
                                 threads, n_events_out, cand_per_win, nullptr, nullptr, 0);
}

static int64_t frontend_windows_impl(const double *t, const double *x, const double *y, const uint8_t *pol,
                                     int64_t n, const double *win, int n_win, double eps, unsigned minS,
                                     unsigned clusterMin, int knn_num, int fitCircleFlag, double Rthr,
                                     unsigned rows_cols, int threads, int64_t *n_events_out, int *cand_per_win,
                                     int *counts, double *cand_out, int cand_cap) {
    std::atomic<int> next(0);
    std::atomic<int64_t> total(0), nev(0);
    auto work = [&]() {
        FrameResult f;
        for (;;) {
            int w = next.fetch_add(1);
            if (w >= n_win) break;
            int64_t lo, hi;
            event_frame(t, x, y, pol, n, win[2 * w], win[2 * w + 1], f.pos, f.neg, &lo, &hi);
            nev += hi - lo;
            extract(f, eps, minS, clusterMin, knn_num, fitCircleFlag, Rthr, rows_cols, false);
            total += (int64_t) f.cand.size();
            if (cand_per_win) cand_per_win[w] = (int) f.cand.size();
            if (counts) {
                int *c = counts + (size_t) 6 * w;
                c[0] = (int) f.neg.size();
                c[1] = (int) f.pos.size();
                c[2] = (int) f.ndb.clusters.size();
                c[3] = (int) f.pdb.clusters.size();
                c[4] = (int) f.ncl.size();
                c[5] = (int) f.pcl.size();
            }
            if (cand_out)
                for (size_t k = 0; k < f.cand.size() && (int) k < cand_cap; ++k) {
                    double *o = cand_out + ((size_t) w * cand_cap + k) * 5;
                    o[0] = f.cand[k].pi;
                    o[1] = f.cand[k].ni;
                    o[2] = f.cand[k].cx;
                    o[3] = f.cand[k].cy;
                    o[4] = f.cand[k].r;
                }
        }
    };
    if (threads <= 1)
        work();
    else {
        std::vector<std::thread> th;
        for (int i = 0; i < threads; ++i) th.emplace_back(work);
        for (auto &h : th) h.join();
    }
    if (n_events_out) *n_events_out = nev.load();
    return total.load();
}
