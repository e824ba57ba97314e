// Grid ordering of candidate circle centres for the asymmetric circle-grid pattern — the RESULT of the points-in
// cv::findCirclesGrid overload the reference calls at CirclesEventFrame.cpp:332-336 (cv_calib/src/cv_calib.cpp:8-88 +
// the vendored circlesgrid.cpp).  OpenCV's internal path (k-means of neighbour vectors with a per-thread RNG, graph
// growing, RANSAC homographies) cannot be matched step by step and is not needed: when the grid is found its output is
// canonical — the 36 pattern points row-major (index i*cols + j <-> board point ((2j + i%2) s, i s), EventCalibIni.cpp:103-106),
// labelled so that board -> image preserves orientation.  For rows odd that labelling is unique: the point set has a mirror
// symmetry (i <-> rows-1-i) but no rotational one.  This header finds it directly: lattice growing from a central seed,
// alignment of the labelled lattice with the board rectangle under the 8 lattice symmetries, then a homography check.
// Checked against cv2.findCirclesGrid on rendered candidates (tests/golden/circles_grid.npz).  Host only; <= ~100 points.
// Differences: the accept / reject decision in hard cases (clutter next to the grid, missing circles) is OpenCV's own
// heuristic and is NOT reproduced — this returns false whenever no complete, consistent labelling exists.  The reference's
// CALIB_CB_CLUSTERING retry is find_asymmetric_circles_grid_clustering() at the end of this file.
#ifndef ECB_CIRCLES_GRID_HPP
#define ECB_CIRCLES_GRID_HPP

#include <algorithm>
#include <array>
#include <cmath>
#include <map>
#include <queue>
#include <utility>
#include <vector>

namespace ecb {

struct Pt2 {
    double x, y;
};

namespace grid_detail {

// homography board -> image from >= 4 correspondences (normalised DLT via the 8x8 normal equations, h33 = 1)
inline bool fit_homography(const std::vector<Pt2> &b, const std::vector<Pt2> &p, double H[9]) {
    const size_t n = b.size();
    if (n < 4) return false;
    auto normalise = [](const std::vector<Pt2> &v, double &cx, double &cy, double &s) {
        cx = cy = 0;
        for (auto &q : v) cx += q.x, cy += q.y;
        cx /= v.size();
        cy /= v.size();
        double d = 0;
        for (auto &q : v) d += std::sqrt((q.x - cx) * (q.x - cx) + (q.y - cy) * (q.y - cy));
        s = d > 0 ? std::sqrt(2.0) * v.size() / d : 1.0;
    };
    double bx, by, bs, px, py, ps;
    normalise(b, bx, by, bs);
    normalise(p, px, py, ps);
    double A[8][9] = {{0}};  // normal equations [A^T A | A^T r]
    for (size_t k = 0; k < n; ++k) {
        const double X = (b[k].x - bx) * bs, Y = (b[k].y - by) * bs, u = (p[k].x - px) * ps, v = (p[k].y - py) * ps;
        const double r1[8] = {X, Y, 1, 0, 0, 0, -u * X, -u * Y}, r2[8] = {0, 0, 0, X, Y, 1, -v * X, -v * Y};
        for (int i = 0; i < 8; ++i) {
            for (int j = 0; j < 8; ++j) A[i][j] += r1[i] * r1[j] + r2[i] * r2[j];
            A[i][8] += r1[i] * u + r2[i] * v;
        }
    }
    for (int c = 0; c < 8; ++c) {  // Gaussian elimination with partial pivoting
        int piv = c;
        for (int r = c + 1; r < 8; ++r)
            if (std::fabs(A[r][c]) > std::fabs(A[piv][c])) piv = r;
        if (std::fabs(A[piv][c]) < 1e-14) return false;
        if (piv != c)
            for (int j = 0; j < 9; ++j) std::swap(A[c][j], A[piv][j]);
        for (int r = 0; r < 8; ++r)
            if (r != c) {
                const double f = A[r][c] / A[c][c];
                for (int j = c; j < 9; ++j) A[r][j] -= f * A[c][j];
            }
    }
    double h[9];
    for (int i = 0; i < 8; ++i) h[i] = A[i][8] / A[i][i];
    h[8] = 1;
    // de-normalise: H = Tp^-1 * Hn * Tb
    const double Tb[9] = {bs, 0, -bs * bx, 0, bs, -bs * by, 0, 0, 1}, Tpi[9] = {1 / ps, 0, px, 0, 1 / ps, py, 0, 0, 1};
    double M[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M[3 * i + j] = h[3 * i] * Tb[j] + h[3 * i + 1] * Tb[3 + j] + h[3 * i + 2] * Tb[6 + j];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) H[3 * i + j] = Tpi[3 * i] * M[j] + Tpi[3 * i + 1] * M[3 + j] + Tpi[3 * i + 2] * M[6 + j];
    return true;
}
inline Pt2 apply_h(const double H[9], double x, double y) {
    const double w = H[6] * x + H[7] * y + H[8];
    return Pt2{(H[0] * x + H[1] * y + H[2]) / w, (H[3] * x + H[4] * y + H[5]) / w};
}

}  // namespace grid_detail

// points: candidate centres; order_out[i*cols + j] = index into points.  Returns true when the asymmetric rows x cols grid
// was found.  max_err: largest accepted reprojection error of the final homography, in units of the local circle spacing.
inline bool find_asymmetric_circles_grid(const std::vector<Pt2> &points, int rows, int cols, std::vector<int> &order_out,
                                         double max_err = 0.25) {
    using namespace grid_detail;
    const int K = (int) points.size(), need = rows * cols;
    order_out.clear();
    if (K < need || rows < 2 || cols < 2) return false;
    // lattice coordinates (X, Y), X + Y even:  board point (i, j) = (2j + i%2, i)
    const int WX = 2 * cols - 1, WY = rows - 1;  // board rectangle [0, WX] x [0, WY]
    Pt2 cen{0, 0};
    for (auto &p : points) cen.x += p.x / K, cen.y += p.y / K;
    std::vector<int> seeds(K);
    for (int k = 0; k < K; ++k) seeds[k] = k;
    std::sort(seeds.begin(), seeds.end(), [&](int a, int b) {
        const double da = (points[a].x - cen.x) * (points[a].x - cen.x) + (points[a].y - cen.y) * (points[a].y - cen.y);
        const double db = (points[b].x - cen.x) * (points[b].x - cen.x) + (points[b].y - cen.y) * (points[b].y - cen.y);
        return da < db || (da == db && a < b);
    });
    auto d2 = [&](int a, int b) {
        return (points[a].x - points[b].x) * (points[a].x - points[b].x) + (points[a].y - points[b].y) * (points[a].y - points[b].y);
    };
    const int step[8][2] = {{1, 1}, {1, -1}, {-1, 1}, {-1, -1}, {2, 0}, {-2, 0}, {0, 2}, {0, -2}};
    for (int si = 0; si < std::min(K, 8); ++si) {
        const int s = seeds[si];
        std::vector<int> nb;  // the seed's nearest neighbours: candidates for the lattice steps (1,1) and (1,-1)
        for (int k = 0; k < K; ++k)
            if (k != s) nb.push_back(k);
        std::sort(nb.begin(), nb.end(), [&](int a, int b) { return d2(a, s) < d2(b, s) || (d2(a, s) == d2(b, s) && a < b); });
        const int nn = std::min((int) nb.size(), 6);
        for (int ia = 0; ia < nn; ++ia)
            for (int ib = 0; ib < nn; ++ib) {
                if (ia == ib) continue;
                const int a = nb[ia], b = nb[ib];
                // hypothesis: a = seed + (1,1), b = seed + (1,-1)  ->  lattice-to-image affine map around the seed
                const double e1x = points[a].x - points[s].x, e1y = points[a].y - points[s].y;
                const double e2x = points[b].x - points[s].x, e2y = points[b].y - points[s].y;
                // columns of A: image displacement per unit X and per unit Y
                const double ax = (e1x + e2x) / 2, ay = (e1y + e2y) / 2, bx = (e1x - e2x) / 2, by = (e1y - e2y) / 2;
                const double det = ax * by - ay * bx;
                if (!(det > 0)) continue;  // board -> image must preserve orientation (OpenCV's labelling)
                const double l1 = std::sqrt(e1x * e1x + e1y * e1y), l2 = std::sqrt(e2x * e2x + e2y * e2y);
                if (l1 > 1.6 * l2 || l2 > 1.6 * l1) continue;
                // grow the lattice: node -> candidate
                std::map<std::pair<int, int>, int> node;
                std::vector<char> used((size_t) K, 0);
                std::queue<std::pair<int, int>> q;
                node[{0, 0}] = s;
                used[(size_t) s] = 1;
                q.push({0, 0});
                while (!q.empty()) {
                    const auto cur = q.front();
                    q.pop();
                    const int ck = node[cur];
                    for (int t = 0; t < 8; ++t) {
                        const std::pair<int, int> nx{cur.first + step[t][0], cur.second + step[t][1]};
                        if (node.count(nx)) continue;
                        if (std::abs(nx.first) > 2 * WX || std::abs(nx.second) > 2 * WY) continue;
                        // local affine map: finite differences of already labelled neighbours of `cur`, else the seed's
                        double lax = ax, lay = ay, lbx = bx, lby = by;
                        {
                            auto it1 = node.find({cur.first + 1, cur.second + 1}), it2 = node.find({cur.first - 1, cur.second - 1});
                            auto it3 = node.find({cur.first + 1, cur.second - 1}), it4 = node.find({cur.first - 1, cur.second + 1});
                            double d1x = 0, d1y = 0, d2x = 0, d2y = 0;
                            bool h1 = false, h2 = false;
                            if (it1 != node.end()) d1x = points[it1->second].x - points[ck].x, d1y = points[it1->second].y - points[ck].y, h1 = true;
                            else if (it2 != node.end()) d1x = points[ck].x - points[it2->second].x, d1y = points[ck].y - points[it2->second].y, h1 = true;
                            if (it3 != node.end()) d2x = points[it3->second].x - points[ck].x, d2y = points[it3->second].y - points[ck].y, h2 = true;
                            else if (it4 != node.end()) d2x = points[ck].x - points[it4->second].x, d2y = points[ck].y - points[it4->second].y, h2 = true;
                            if (h1 && h2) {
                                lax = (d1x + d2x) / 2, lay = (d1y + d2y) / 2;
                                lbx = (d1x - d2x) / 2, lby = (d1y - d2y) / 2;
                            }
                        }
                        const double px = points[ck].x + lax * step[t][0] + lbx * step[t][1];
                        const double py = points[ck].y + lay * step[t][0] + lby * step[t][1];
                        const double sp2 = (lax + lbx) * (lax + lbx) + (lay + lby) * (lay + lby);  // |(1,1) step|^2
                        int best = -1;
                        double bd = 0.16 * sp2;  // within 0.4 of the diagonal spacing
                        for (int k = 0; k < K; ++k) {
                            if (used[(size_t) k]) continue;
                            const double dd = (points[k].x - px) * (points[k].x - px) + (points[k].y - py) * (points[k].y - py);
                            if (dd < bd) bd = dd, best = k;
                        }
                        if (best < 0) continue;
                        node[nx] = best;
                        used[(size_t) best] = 1;
                        q.push(nx);
                    }
                }
                if ((int) node.size() < need) continue;
                // align with the board: the 4 lattice symmetries that keep orientation (rotations by 0/90/180/270 deg) x translation
                for (int rot = 0; rot < 4; ++rot) {
                    auto rotate = [&](int X, int Y, int &RX, int &RY) {
                        switch (rot) {
                            case 0: RX = X, RY = Y; break;
                            case 1: RX = -Y, RY = X; break;
                            case 2: RX = -X, RY = -Y; break;
                            default: RX = Y, RY = -X; break;
                        }
                    };
                    std::map<std::pair<int, int>, int> rn;
                    for (auto &kv : node) {
                        int RX, RY;
                        rotate(kv.first.first, kv.first.second, RX, RY);
                        rn[{RX, RY}] = kv.second;
                    }
                    for (auto &origin : rn) {  // a labelled node that would be board point (0, 0)
                        const int ox = origin.first.first, oy = origin.first.second;
                        std::vector<int> order((size_t) need, -1);
                        bool ok = true;
                        for (int i = 0; i < rows && ok; ++i)
                            for (int j = 0; j < cols; ++j) {
                                auto it = rn.find({ox + 2 * j + i % 2, oy + i});
                                if (it == rn.end()) {
                                    ok = false;
                                    break;
                                }
                                order[(size_t) (i * cols + j)] = it->second;
                            }
                        if (!ok) continue;
                        // no labelled node may sit right outside the rectangle in the same lattice (the grid would be larger)
                        std::vector<Pt2> bp, ip;
                        for (int i = 0; i < rows; ++i)
                            for (int j = 0; j < cols; ++j) {
                                bp.push_back(Pt2{(double) (2 * j + i % 2), (double) i});
                                ip.push_back(points[(size_t) order[(size_t) (i * cols + j)]]);
                            }
                        double H[9];
                        if (!fit_homography(bp, ip, H)) continue;
                        // orientation and fit quality
                        const Pt2 c0 = apply_h(H, WX / 2.0, WY / 2.0), cx1 = apply_h(H, WX / 2.0 + 1, WY / 2.0), cy1 = apply_h(H, WX / 2.0, WY / 2.0 + 1);
                        const double jd = (cx1.x - c0.x) * (cy1.y - c0.y) - (cx1.y - c0.y) * (cy1.x - c0.x);
                        if (!(jd > 0)) continue;
                        double worst = 0;
                        for (size_t k = 0; k < bp.size(); ++k) {
                            const Pt2 r = apply_h(H, bp[k].x, bp[k].y), r1 = apply_h(H, bp[k].x + 1, bp[k].y + 1);
                            const double sp = std::sqrt((r1.x - r.x) * (r1.x - r.x) + (r1.y - r.y) * (r1.y - r.y));
                            const double e = std::sqrt((r.x - ip[k].x) * (r.x - ip[k].x) + (r.y - ip[k].y) * (r.y - ip[k].y)) / std::max(sp, 1e-9);
                            worst = std::max(worst, e);
                        }
                        if (worst > max_err) continue;
                        order_out = order;
                        return true;
                    }
                }
            }
    }
    return false;
}

// The reference's SECOND attempt, `findCirclesGrid(..., CALIB_CB_ASYMMETRIC_GRID | CALIB_CB_CLUSTERING)`
// (CirclesEventFrame.cpp:334-336 -> cv_calib.cpp:24-31 -> CirclesGridClusterFinder::findGrid, circlesgrid.cpp:132-176).
// Its first stage decides WHICH candidates form the pattern and is restated step by step (hierarchicalClustering,
// circlesgrid.cpp:72-130): single-linkage agglomeration on the float distance matrix — always the globally smallest
// remaining distance (first one in row-major order, like minMaxLoc), the higher index merged into the lower — until the
// cluster just grown holds rows * cols points; more than that means "not found".  The following stages of OpenCV (convex
// hull corners, rectification, parsing) only label the selected points; the label set is canonical, so the selected points
// are labelled by the lattice finder above.
inline bool hierarchical_cluster_select(const std::vector<Pt2> &points, size_t pn, std::vector<int> &selected) {
    const int n = (int) points.size();
    selected.clear();
    if (pn >= points.size()) {
        if (pn == points.size())
            for (int i = 0; i < n; ++i) selected.push_back(i);
        return !selected.empty();
    }
    std::vector<float> dists((size_t) n * n, 0.0f);
    std::vector<unsigned char> mask((size_t) n * n, 0);
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j) {
            // norm(Point2f - Point2f): float difference, double accumulation and sqrt, stored as float
            const float dx = (float) points[(size_t) i].x - (float) points[(size_t) j].x, dy = (float) points[(size_t) i].y - (float) points[(size_t) j].y;
            const float d = (float) std::sqrt((double) dx * dx + (double) dy * dy);
            dists[(size_t) i * n + j] = dists[(size_t) j * n + i] = d;
            mask[(size_t) i * n + j] = mask[(size_t) j * n + i] = 255;
        }
    std::vector<std::vector<int>> clusters((size_t) n);
    for (int i = 0; i < n; ++i) clusters[(size_t) i].push_back(i);
    int pattern = 0;
    while (clusters[(size_t) pattern].size() < pn) {
        int br = -1, bc = -1;
        float best = 0;
        for (int r = 0; r < n; ++r)
            for (int c = 0; c < n; ++c)
                if (mask[(size_t) r * n + c] && (br < 0 || dists[(size_t) r * n + c] < best)) best = dists[(size_t) r * n + c], br = r, bc = c;
        if (br < 0) return false;
        const int lo = std::min(br, bc), hi = std::max(br, bc);
        for (int k = 0; k < n; ++k) mask[(size_t) hi * n + k] = mask[(size_t) k * n + hi] = 0;
        for (int k = 0; k < n; ++k) {  // row lo = min(row minLoc.x, row minLoc.y), mirrored into column lo
            const float v = std::min(dists[(size_t) bc * n + k], dists[(size_t) br * n + k]);
            dists[(size_t) lo * n + k] = v;
        }
        for (int k = 0; k < n; ++k) dists[(size_t) k * n + lo] = dists[(size_t) lo * n + k];
        clusters[(size_t) lo].insert(clusters[(size_t) lo].end(), clusters[(size_t) hi].begin(), clusters[(size_t) hi].end());
        clusters[(size_t) hi].clear();
        pattern = lo;
    }
    if (clusters[(size_t) pattern].size() != pn) return false;  // the cluster overshot the pattern size
    selected = clusters[(size_t) pattern];
    return true;
}

inline bool find_asymmetric_circles_grid_clustering(const std::vector<Pt2> &points, int rows, int cols, std::vector<int> &order_out,
                                                    double max_err = 0.25) {
    std::vector<int> sel;
    order_out.clear();
    if (points.empty() || !hierarchical_cluster_select(points, (size_t) rows * (size_t) cols, sel)) return false;
    std::vector<Pt2> sub;
    for (int k : sel) sub.push_back(points[(size_t) k]);
    std::vector<int> o;
    if (!find_asymmetric_circles_grid(sub, rows, cols, o, max_err)) return false;
    for (int k : o) order_out.push_back(sel[(size_t) k]);
    return true;
}

}  // namespace ecb
#endif  // ECB_CIRCLES_GRID_HPP
