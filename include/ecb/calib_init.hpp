// include/ecb/calib_init.hpp — SURVEY §8 row f-4: the initialisation stage between the circle detection and the spline
// optimisation, without OpenCV.  Host only (small, sequential: ≤ 200 views × 36 points); it feeds the GPU cost evaluation.
//
// Replaces, with the same argument meaning (reference paths relative to modules/camera_calibration/event_camera_calib):
//   * cv::calibrateCamera as called by EventCalibIni::cvCalibration (src/EventCalibIni.cpp:143-210: planar board, flags
//     from CalibrationSetting::validate, include/opengv2/event_camera_calib/parameters.hpp:49-60, 5-coefficient model
//     k1 k2 p1 p2 k3)  ->  ecb::calibrateCamera
//   * cv::projectPoints (EventCalibIni.cpp:115-141 computeReprojectionErrors, src/CirclesEventFrame.cpp:431-456) -> ecb::projectPoints
//   * cv::solvePnPRansac(..., 50, 4.0, 0.99, inliers, SOLVEPNP_IPPE) (EventCalibIni.cpp:255-256) -> ecb::solvePnPPlanar
//   * cv::Rodrigues, Eigen::Quaterniond(R) (:263-270) -> ecb::rodrigues, ecb::rotationToQuaternion
//   * EventCalibIni::checkPose (:328-346) -> ecb::checkPose
//
// PARITY UNPINNED: OpenCV is an external dependency that is not in /root/reference and not in this image as a C++ library.
// The published algorithms are restated (Zhang's closed-form initialisation from plane homographies with the principal point
// at the image centre, Levenberg-Marquardt on the reprojection error with the masked parameters of the flags); the maximum-
// likelihood optimum they converge to is unique, and tests/test_calib_init.py checks it against the cv2 4.13 wheel's
// calibrateCamera / solvePnP / projectPoints on the same float32 inputs (fixtures in tests/golden/, written by
// tests/golden/make_calib_init_golden.py).  Known deviations: OpenCV stops its LM after 30 iterations, this one iterates to
// convergence; solvePnPRansac draws random minimal sets and ends with the non-iterative IPPE solution on the inliers, here
// the pose is the reprojection-error minimum over the points within the 4 px gate (deterministic), which agrees with IPPE
// to the noise level of the centres.
#ifndef ECB_CALIB_INIT_HPP
#define ECB_CALIB_INIT_HPP
#include <algorithm>
#include <array>
#include <cmath>
#include <cstddef>
#include <vector>

namespace ecb {

// ---- forward-mode dual numbers (the Jacobians of the 15-parameter projection) ----
template <int N>
struct Dual {
    double v;
    double d[N];
    Dual() : v(0) { for (int i = 0; i < N; ++i) d[i] = 0; }
    Dual(double x) : v(x) { for (int i = 0; i < N; ++i) d[i] = 0; }
    static Dual var(double x, int i) { Dual r(x); r.d[i] = 1; return r; }
};
template <int N> inline Dual<N> operator+(const Dual<N> &a, const Dual<N> &b) { Dual<N> r; r.v = a.v + b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
template <int N> inline Dual<N> operator-(const Dual<N> &a, const Dual<N> &b) { Dual<N> r; r.v = a.v - b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
template <int N> inline Dual<N> operator-(const Dual<N> &a) { Dual<N> r; r.v = -a.v; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }
template <int N> inline Dual<N> operator*(const Dual<N> &a, const Dual<N> &b) { Dual<N> r; r.v = a.v * b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <int N> inline Dual<N> operator/(const Dual<N> &a, const Dual<N> &b) { Dual<N> r; const double ib = 1.0 / b.v; r.v = a.v * ib; for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib; return r; }
template <int N> inline Dual<N> sqrt(const Dual<N> &a) { Dual<N> r; r.v = std::sqrt(a.v); const double s = 0.5 / r.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * s; return r; }
template <int N> inline Dual<N> sin(const Dual<N> &a) { Dual<N> r; r.v = std::sin(a.v); const double c = std::cos(a.v); for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * c; return r; }
template <int N> inline Dual<N> cos(const Dual<N> &a) { Dual<N> r; r.v = std::cos(a.v); const double s = -std::sin(a.v); for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * s; return r; }
inline double value_of(double x) { return x; }
template <int N> inline double value_of(const Dual<N> &x) { return x.v; }

// ---- rotations ----
// cv::Rodrigues, vector -> matrix (row-major R[9]); first-order branch at the origin keeps the derivative finite
template <class T>
inline void rodrigues(const T r[3], T R[9]) {
    using std::cos;
    using std::sin;
    using std::sqrt;
    const T th2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
    if (value_of(th2) < 1e-24) {
        R[0] = T(1.0); R[1] = -r[2]; R[2] = r[1];
        R[3] = r[2]; R[4] = T(1.0); R[5] = -r[0];
        R[6] = -r[1]; R[7] = r[0]; R[8] = T(1.0);
        return;
    }
    const T th = sqrt(th2), c = cos(th), s = sin(th), c1 = T(1.0) - c;
    const T x = r[0] / th, y = r[1] / th, z = r[2] / th;
    R[0] = c + c1 * x * x;     R[1] = c1 * x * y - s * z; R[2] = c1 * x * z + s * y;
    R[3] = c1 * x * y + s * z; R[4] = c + c1 * y * y;     R[5] = c1 * y * z - s * x;
    R[6] = c1 * x * z - s * y; R[7] = c1 * y * z + s * x; R[8] = c + c1 * z * z;
}
// Eigen::Quaterniond(R): x y z w (the branch structure of Eigen's quaternionbase_assign_impl<..., 3, 3>)
inline void rotationToQuaternion(const double R[9], double q[4]) {
    double t = R[0] + R[4] + R[8];
    if (t > 0) {
        t = std::sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (R[7] - R[5]) * t;
        q[1] = (R[2] - R[6]) * t;
        q[2] = (R[3] - R[1]) * t;
    } else {
        int i = 0;
        if (R[4] > R[0]) i = 1;
        if (R[8] > R[4 * i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        q[3] = (R[3 * k + j] - R[3 * j + k]) * t;
        q[j] = (R[3 * j + i] + R[3 * i + j]) * t;
        q[k] = (R[3 * k + i] + R[3 * i + k]) * t;
    }
}
// cv::Rodrigues, matrix -> vector (through the unit quaternion; angle in [0, pi])
inline void rodriguesInverse(const double R[9], double r[3]) {
    double q[4];
    rotationToQuaternion(R, q);
    const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 4; ++i) q[i] /= n;
    if (q[3] < 0)
        for (int i = 0; i < 4; ++i) q[i] = -q[i];
    const double s = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
    const double k = s < 1e-12 ? 2.0 : 2.0 * std::atan2(s, q[3]) / s;
    for (int i = 0; i < 3; ++i) r[i] = k * q[i];
}

// ---- the OpenCV pinhole + (k1 k2 p1 p2 k3) model ----
struct CameraModel {
    double fx = 1, fy = 1, cx = 0, cy = 0;
    double dist[5] = {0, 0, 0, 0, 0};  // k1 k2 p1 p2 k3
};
template <class T>
inline void projectPoint(const T intr[4], const T dist[5], const T rv[3], const T tv[3], const double X[3], T uv[2]) {
    T R[9];
    rodrigues(rv, R);
    const T x = R[0] * T(X[0]) + R[1] * T(X[1]) + R[2] * T(X[2]) + tv[0];
    const T y = R[3] * T(X[0]) + R[4] * T(X[1]) + R[5] * T(X[2]) + tv[1];
    const T z = R[6] * T(X[0]) + R[7] * T(X[1]) + R[8] * T(X[2]) + tv[2];
    const T a = x / z, b = y / z, r2 = a * a + b * b, r4 = r2 * r2, r6 = r4 * r2;
    const T radial = T(1.0) + dist[0] * r2 + dist[1] * r4 + dist[4] * r6;
    const T xd = a * radial + T(2.0) * dist[2] * a * b + dist[3] * (r2 + T(2.0) * a * a);
    const T yd = b * radial + dist[2] * (r2 + T(2.0) * b * b) + T(2.0) * dist[3] * a * b;
    uv[0] = intr[0] * xd + intr[2];
    uv[1] = intr[1] * yd + intr[3];
}
// cv::projectPoints: n object points (xyz) -> n image points (uv), double precision (OpenCV returns Point2f when asked to;
// round at the call site where the reference does)
inline void projectPoints(const double *obj, int n, const double rvec[3], const double tvec[3], const CameraModel &cam, double *img) {
    const double intr[4] = {cam.fx, cam.fy, cam.cx, cam.cy};
    for (int i = 0; i < n; ++i) projectPoint<double>(intr, cam.dist, rvec, tvec, obj + 3 * i, img + 2 * i);
}
// cv::undistortPoints without R/P: pixel -> ideal normalised coordinates (fixed point iteration like OpenCV's)
inline void undistortPoint(const CameraModel &cam, const double uv[2], double xy[2]) {
    const double x0 = (uv[0] - cam.cx) / cam.fx, y0 = (uv[1] - cam.cy) / cam.fy;
    double x = x0, y = y0;
    for (int it = 0; it < 50; ++it) {
        const double r2 = x * x + y * y;
        const double icd = 1.0 / (1 + ((cam.dist[4] * r2 + cam.dist[1]) * r2 + cam.dist[0]) * r2);
        const double dx = 2 * cam.dist[2] * x * y + cam.dist[3] * (r2 + 2 * x * x);
        const double dy = cam.dist[2] * (r2 + 2 * y * y) + 2 * cam.dist[3] * x * y;
        const double xn = (x0 - dx) * icd, yn = (y0 - dy) * icd;
        const double ch = std::abs(xn - x) + std::abs(yn - y);
        x = xn;
        y = yn;
        if (ch < 1e-15) break;
    }
    xy[0] = x;
    xy[1] = y;
}

// ---- small dense linear algebra ----
// A x = b by Gaussian elimination with partial pivoting (A: n x n row-major, destroyed); false when singular
inline bool solveLinear(int n, std::vector<double> &A, std::vector<double> &b) {
    for (int c = 0; c < n; ++c) {
        int p = c;
        for (int r = c + 1; r < n; ++r)
            if (std::abs(A[(size_t) r * n + c]) > std::abs(A[(size_t) p * n + c])) p = r;
        if (!(std::abs(A[(size_t) p * n + c]) > 0)) return false;
        if (p != c) {
            for (int k = 0; k < n; ++k) std::swap(A[(size_t) p * n + k], A[(size_t) c * n + k]);
            std::swap(b[(size_t) p], b[(size_t) c]);
        }
        const double ip = 1.0 / A[(size_t) c * n + c];
        for (int r = c + 1; r < n; ++r) {
            const double f = A[(size_t) r * n + c] * ip;
            if (f == 0) continue;
            for (int k = c; k < n; ++k) A[(size_t) r * n + k] -= f * A[(size_t) c * n + k];
            b[(size_t) r] -= f * b[(size_t) c];
        }
    }
    for (int r = n - 1; r >= 0; --r) {
        double v = b[(size_t) r];
        for (int k = r + 1; k < n; ++k) v -= A[(size_t) r * n + k] * b[(size_t) k];
        b[(size_t) r] = v / A[(size_t) r * n + r];
    }
    return true;
}
// eigenvector of the smallest eigenvalue of a symmetric n x n matrix (cyclic Jacobi)
inline void smallestEigenvector(int n, std::vector<double> M, std::vector<double> &vec) {
    std::vector<double> V((size_t) n * n, 0.0);
    for (int i = 0; i < n; ++i) V[(size_t) i * n + i] = 1;
    for (int sweep = 0; sweep < 100; ++sweep) {
        double off = 0, dg = 0;
        for (int p = 0; p < n; ++p) {
            dg += M[(size_t) p * n + p] * M[(size_t) p * n + p];
            for (int q = p + 1; q < n; ++q) off += M[(size_t) p * n + q] * M[(size_t) p * n + q];
        }
        if (off <= 1e-60 * dg || off == 0) break;
        for (int p = 0; p < n; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = M[(size_t) p * n + q];
                if (apq == 0) continue;
                const double th = (M[(size_t) q * n + q] - M[(size_t) p * n + p]) / (2 * apq);
                const double t = (th >= 0 ? 1.0 : -1.0) / (std::abs(th) + std::sqrt(th * th + 1));
                const double c = 1 / std::sqrt(t * t + 1), s = t * c;
                for (int k = 0; k < n; ++k) {
                    const double a = M[(size_t) k * n + p], b = M[(size_t) k * n + q];
                    M[(size_t) k * n + p] = c * a - s * b;
                    M[(size_t) k * n + q] = s * a + c * b;
                }
                for (int k = 0; k < n; ++k) {
                    const double a = M[(size_t) p * n + k], b = M[(size_t) q * n + k];
                    M[(size_t) p * n + k] = c * a - s * b;
                    M[(size_t) q * n + k] = s * a + c * b;
                }
                for (int k = 0; k < n; ++k) {
                    const double a = V[(size_t) k * n + p], b = V[(size_t) k * n + q];
                    V[(size_t) k * n + p] = c * a - s * b;
                    V[(size_t) k * n + q] = s * a + c * b;
                }
            }
    }
    int m = 0;
    for (int k = 1; k < n; ++k)
        if (M[(size_t) k * n + k] < M[(size_t) m * n + m]) m = k;
    vec.resize((size_t) n);
    for (int k = 0; k < n; ++k) vec[(size_t) k] = V[(size_t) k * n + m];
}

// plane homography dst ~ H src (n >= 4 correspondences, xy pairs) by the normalised DLT, then Gauss-Newton on the transfer
// error (what cv::findHomography(method 0) does after its linear estimate)
inline bool findHomography(const double *src, const double *dst, int n, double H[9]) {
    if (n < 4) return false;
    double ms[2] = {0, 0}, md[2] = {0, 0}, ss = 0, sd = 0;
    for (int i = 0; i < n; ++i) {
        ms[0] += src[2 * i], ms[1] += src[2 * i + 1];
        md[0] += dst[2 * i], md[1] += dst[2 * i + 1];
    }
    for (int c = 0; c < 2; ++c) ms[c] /= n, md[c] /= n;
    for (int i = 0; i < n; ++i) {
        ss += std::hypot(src[2 * i] - ms[0], src[2 * i + 1] - ms[1]);
        sd += std::hypot(dst[2 * i] - md[0], dst[2 * i + 1] - md[1]);
    }
    if (!(ss > 0) || !(sd > 0)) return false;
    ss = std::sqrt(2.0) * n / ss;
    sd = std::sqrt(2.0) * n / sd;
    std::vector<double> M(81, 0.0);
    for (int i = 0; i < n; ++i) {
        const double x = (src[2 * i] - ms[0]) * ss, y = (src[2 * i + 1] - ms[1]) * ss;
        const double u = (dst[2 * i] - md[0]) * sd, v = (dst[2 * i + 1] - md[1]) * sd;
        const double r1[9] = {x, y, 1, 0, 0, 0, -u * x, -u * y, -u}, r2[9] = {0, 0, 0, x, y, 1, -v * x, -v * y, -v};
        for (int a = 0; a < 9; ++a)
            for (int b = 0; b < 9; ++b) M[(size_t) a * 9 + b] += r1[a] * r1[b] + r2[a] * r2[b];
    }
    std::vector<double> h;
    smallestEigenvector(9, M, h);
    // Gauss-Newton refinement of the transfer error in the normalised frame, h[8] free, scale fixed by |h| = 1 per step
    for (int it = 0; it < 10; ++it) {
        std::vector<double> A(81, 0.0), g(9, 0.0);
        for (int i = 0; i < n; ++i) {
            const double x = (src[2 * i] - ms[0]) * ss, y = (src[2 * i + 1] - ms[1]) * ss;
            const double u = (dst[2 * i] - md[0]) * sd, v = (dst[2 * i + 1] - md[1]) * sd;
            const double w = h[6] * x + h[7] * y + h[8], iw = 1 / w;
            const double pu = (h[0] * x + h[1] * y + h[2]) * iw, pv = (h[3] * x + h[4] * y + h[5]) * iw;
            const double ju[9] = {x * iw, y * iw, iw, 0, 0, 0, -pu * x * iw, -pu * y * iw, -pu * iw};
            const double jv[9] = {0, 0, 0, x * iw, y * iw, iw, -pv * x * iw, -pv * y * iw, -pv * iw};
            for (int a = 0; a < 9; ++a) {
                for (int b = 0; b < 9; ++b) A[(size_t) a * 9 + b] += ju[a] * ju[b] + jv[a] * jv[b];
                g[(size_t) a] += ju[a] * (u - pu) + jv[a] * (v - pv);
            }
        }
        // gauge: the step is orthogonal to h (add h h^T to the singular normal matrix)
        for (int a = 0; a < 9; ++a)
            for (int b = 0; b < 9; ++b) A[(size_t) a * 9 + b] += h[(size_t) a] * h[(size_t) b];
        if (!solveLinear(9, A, g)) break;
        double nn = 0, st = 0;
        for (int a = 0; a < 9; ++a) h[(size_t) a] += g[(size_t) a], st += g[(size_t) a] * g[(size_t) a];
        for (int a = 0; a < 9; ++a) nn += h[(size_t) a] * h[(size_t) a];
        nn = std::sqrt(nn);
        for (int a = 0; a < 9; ++a) h[(size_t) a] /= nn;
        if (st < 1e-28) break;
    }
    // H = Td^-1 Hn Ts
    const double Hn[9] = {h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8]};
    double A[9];  // Hn * Ts,  Ts = [ss 0 -ss ms0; 0 ss -ss ms1; 0 0 1]
    for (int r = 0; r < 3; ++r) {
        A[3 * r] = Hn[3 * r] * ss;
        A[3 * r + 1] = Hn[3 * r + 1] * ss;
        A[3 * r + 2] = Hn[3 * r + 2] - ss * (Hn[3 * r] * ms[0] + Hn[3 * r + 1] * ms[1]);
    }
    for (int c = 0; c < 3; ++c) {  // Td^-1 = [1/sd 0 md0; 0 1/sd md1; 0 0 1]
        H[c] = A[c] / sd + md[0] * A[6 + c];
        H[3 + c] = A[3 + c] / sd + md[1] * A[6 + c];
        H[6 + c] = A[6 + c];
    }
    if (H[8] != 0) {
        const double s = 1 / H[8];
        for (int k = 0; k < 9; ++k) H[k] *= s;
    }
    return true;
}

// nearest rotation to M (row-major 3x3) by Newton's polar iteration; det forced positive by the caller's sign choice
inline void nearestRotation(double M[9]) {
    for (int it = 0; it < 60; ++it) {
        const double det = M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
        const double id = 1 / det;
        const double invT[9] = {(M[4] * M[8] - M[5] * M[7]) * id, (M[5] * M[6] - M[3] * M[8]) * id, (M[3] * M[7] - M[4] * M[6]) * id,
                                (M[2] * M[7] - M[1] * M[8]) * id, (M[0] * M[8] - M[2] * M[6]) * id, (M[1] * M[6] - M[0] * M[7]) * id,
                                (M[1] * M[5] - M[2] * M[4]) * id, (M[2] * M[3] - M[0] * M[5]) * id, (M[0] * M[4] - M[1] * M[3]) * id};
        double ch = 0;
        for (int k = 0; k < 9; ++k) {
            const double nv = 0.5 * (M[k] + invT[k]);
            ch += std::abs(nv - M[k]);
            M[k] = nv;
        }
        if (ch < 1e-15) break;
    }
}

// pose of the z = 0 board from the homography board(x,y) -> normalised image coordinates: H ~ [r1 r2 t]
inline void poseFromHomography(const double H[9], double rvec[3], double tvec[3]) {
    double h1[3] = {H[0], H[3], H[6]}, h2[3] = {H[1], H[4], H[7]}, h3[3] = {H[2], H[5], H[8]};
    const double n1 = std::sqrt(h1[0] * h1[0] + h1[1] * h1[1] + h1[2] * h1[2]);
    const double n2 = std::sqrt(h2[0] * h2[0] + h2[1] * h2[1] + h2[2] * h2[2]);
    double s = 2.0 / (n1 + n2);
    if (h3[2] * s < 0) s = -s;  // the board lies in front of the camera
    for (int k = 0; k < 3; ++k) h1[k] *= s, h2[k] *= s, h3[k] *= s;
    double R[9] = {h1[0], h2[0], h1[1] * h2[2] - h1[2] * h2[1],
                   h1[1], h2[1], h1[2] * h2[0] - h1[0] * h2[2],
                   h1[2], h2[2], h1[0] * h2[1] - h1[1] * h2[0]};
    nearestRotation(R);
    rodriguesInverse(R, rvec);
    for (int k = 0; k < 3; ++k) tvec[k] = h3[k];
}

// ---- bundle of views: Levenberg-Marquardt on the reprojection error ----
struct CalibFlags {  // the cv::CALIB_* bits CalibrationSetting::validate builds (parameters.hpp:49-60)
    bool fixPrincipalPoint = false;  // CALIB_FIX_PRINCIPAL_POINT
    bool zeroTangentDist = false;    // CALIB_ZERO_TANGENT_DIST
    bool fixAspectRatio = false;     // CALIB_FIX_ASPECT_RATIO: fx = aspectRatio * fy throughout
    bool fixK1 = false, fixK2 = false, fixK3 = false;  // k4..k6 do not exist in the 5-coefficient model the reference keeps
    bool fixIntrinsics = false;      // pose-only refinement (solvePnP)
    double aspectRatio = 1.0;        // cameraMatrix(0,0) / cameraMatrix(1,1) on entry (EventCalibIni.cpp:148-149)
};
struct View {
    std::vector<double> img;  // n x 2
    double rvec[3] = {0, 0, 0}, tvec[3] = {0, 0, 1};
};
namespace detail {
// sum of squared reprojection errors over all views
inline double reprojectionSSE(const std::vector<double> &obj, const std::vector<View> &views, const CameraModel &cam) {
    const int n = (int) obj.size() / 3;
    double sse = 0, uv[2];
    const double intr[4] = {cam.fx, cam.fy, cam.cx, cam.cy};
    for (const View &v : views)
        for (int i = 0; i < n; ++i) {
            projectPoint<double>(intr, cam.dist, v.rvec, v.tvec, &obj[(size_t) 3 * i], uv);
            const double ex = uv[0] - v.img[(size_t) 2 * i], ey = uv[1] - v.img[(size_t) 2 * i + 1];
            sse += ex * ex + ey * ey;
        }
    return sse;
}
// LM over (free intrinsics | 6 per view) with the arrow structure eliminated view by view (Schur complement).
// mask: optional per-view per-point weights (0/1) — null = all points
inline double refine(const std::vector<double> &obj, std::vector<View> &views, CameraModel &cam, const CalibFlags &fl,
                     int max_iter = 200) {
    typedef Dual<15> D;
    const int n = (int) obj.size() / 3, nv = (int) views.size();
    // free shared parameters -> columns of the 9-vector (fx fy cx cy k1 k2 p1 p2 k3)
    bool free9[9] = {!fl.fixAspectRatio, true, !fl.fixPrincipalPoint, !fl.fixPrincipalPoint,
                     !fl.fixK1, !fl.fixK2, !fl.zeroTangentDist, !fl.zeroTangentDist, !fl.fixK3};
    if (fl.fixIntrinsics)
        for (bool &f : free9) f = false;
    int col[9], na = 0;
    for (int k = 0; k < 9; ++k) col[k] = free9[k] ? na++ : -1;
    double lambda = 1e-3;
    double sse = reprojectionSSE(obj, views, cam);
    for (int iter = 0; iter < max_iter; ++iter) {
        std::vector<double> A((size_t) na * na, 0.0), ga((size_t) na, 0.0);
        std::vector<double> B((size_t) nv * na * 6, 0.0), C((size_t) nv * 36, 0.0), gv((size_t) nv * 6, 0.0);
        D intr[4] = {D::var(cam.fx, 0), D::var(cam.fy, 1), D::var(cam.cx, 2), D::var(cam.cy, 3)}, dist[5];
        for (int k = 0; k < 5; ++k) dist[k] = D::var(cam.dist[k], 4 + k);
        for (int vi = 0; vi < nv; ++vi) {
            const View &v = views[(size_t) vi];
            D rv[3], tv[3], uv[2];
            for (int k = 0; k < 3; ++k) rv[k] = D::var(v.rvec[k], 9 + k), tv[k] = D::var(v.tvec[k], 12 + k);
            for (int i = 0; i < n; ++i) {
                projectPoint<D>(intr, dist, rv, tv, &obj[(size_t) 3 * i], uv);
                for (int c = 0; c < 2; ++c) {
                    const double e = v.img[(size_t) 2 * i + c] - uv[c].v;  // residual = observed - predicted
                    double ja[9], *jp = uv[c].d + 9;
                    int m = 0;
                    for (int k = 0; k < 9; ++k)
                        if (col[k] >= 0) ja[m++] = uv[c].d[k] + (k == 1 && fl.fixAspectRatio ? fl.aspectRatio * uv[c].d[0] : 0.0);
                    for (int a = 0; a < na; ++a) {
                        for (int b = 0; b < na; ++b) A[(size_t) a * na + b] += ja[a] * ja[b];
                        ga[(size_t) a] += ja[a] * e;
                        for (int b = 0; b < 6; ++b) B[((size_t) vi * na + a) * 6 + b] += ja[a] * jp[b];
                    }
                    for (int a = 0; a < 6; ++a) {
                        for (int b = 0; b < 6; ++b) C[(size_t) vi * 36 + a * 6 + b] += jp[a] * jp[b];
                        gv[(size_t) vi * 6 + a] += jp[a] * e;
                    }
                }
            }
        }
        bool accepted = false;
        for (int attempt = 0; attempt < 30 && !accepted; ++attempt) {
            // Schur complement with the diagonals scaled by (1 + lambda) like CvLevMarq
            std::vector<double> S = A, rhs = ga, CiB((size_t) nv * 6 * std::max(na, 1)), Cig((size_t) nv * 6);
            for (int a = 0; a < na; ++a) S[(size_t) a * na + a] *= 1 + lambda;
            bool ok = true;
            for (int vi = 0; vi < nv && ok; ++vi) {
                // [C^-1 B^T | C^-1 g] through one elimination per column
                for (int c = 0; c <= na; ++c) {
                    std::vector<double> M(C.begin() + (size_t) vi * 36, C.begin() + (size_t) (vi + 1) * 36), b(6);
                    for (int a = 0; a < 6; ++a) M[(size_t) a * 6 + a] *= 1 + lambda;
                    for (int a = 0; a < 6; ++a) b[(size_t) a] = c < na ? B[((size_t) vi * na + c) * 6 + a] : gv[(size_t) vi * 6 + a];
                    if (!solveLinear(6, M, b)) {
                        ok = false;
                        break;
                    }
                    for (int a = 0; a < 6; ++a) (c < na ? CiB[((size_t) vi * 6 + a) * na + c] : Cig[(size_t) vi * 6 + a]) = b[(size_t) a];
                }
                if (!ok) break;
                for (int a = 0; a < na; ++a) {
                    for (int b = 0; b < na; ++b) {
                        double s = 0;
                        for (int k = 0; k < 6; ++k) s += B[((size_t) vi * na + a) * 6 + k] * CiB[((size_t) vi * 6 + k) * na + b];
                        S[(size_t) a * na + b] -= s;
                    }
                    double s = 0;
                    for (int k = 0; k < 6; ++k) s += B[((size_t) vi * na + a) * 6 + k] * Cig[(size_t) vi * 6 + k];
                    rhs[(size_t) a] -= s;
                }
            }
            if (ok && na > 0) ok = solveLinear(na, S, rhs);
            if (!ok) {
                lambda *= 10;
                continue;
            }
            CameraModel trial = cam;
            double *p9[9] = {&trial.fx, &trial.fy, &trial.cx, &trial.cy, &trial.dist[0], &trial.dist[1], &trial.dist[2], &trial.dist[3], &trial.dist[4]};
            double step2 = 0, par2 = 0;
            for (int k = 0; k < 9; ++k)
                if (col[k] >= 0) {
                    *p9[k] += rhs[(size_t) col[k]];
                    step2 += rhs[(size_t) col[k]] * rhs[(size_t) col[k]];
                    par2 += *p9[k] * *p9[k];
                }
            if (fl.fixAspectRatio && !fl.fixIntrinsics) trial.fx = trial.fy * fl.aspectRatio;
            std::vector<View> tviews = views;
            for (int vi = 0; vi < nv; ++vi)
                for (int a = 0; a < 6; ++a) {
                    double d = Cig[(size_t) vi * 6 + a];
                    for (int c = 0; c < na; ++c) d -= CiB[((size_t) vi * 6 + a) * na + c] * rhs[(size_t) c];
                    (a < 3 ? tviews[(size_t) vi].rvec[a] : tviews[(size_t) vi].tvec[a - 3]) += d;
                    step2 += d * d;
                    const double pv = a < 3 ? tviews[(size_t) vi].rvec[a] : tviews[(size_t) vi].tvec[a - 3];
                    par2 += pv * pv;
                }
            const double tsse = reprojectionSSE(obj, tviews, trial);
            if (tsse <= sse) {
                const bool tiny = step2 <= 1e-26 * (par2 + 1e-26) || sse - tsse <= 1e-16 * sse;
                cam = trial;
                views.swap(tviews);
                sse = tsse;
                lambda = std::max(lambda * 0.1, 1e-12);
                accepted = true;
                if (tiny) return sse;
            } else {
                lambda *= 10;
            }
        }
        if (!accepted) break;
    }
    return sse;
}
}  // namespace detail

// cv::calibrateCamera for views of ONE planar board (object points z = 0, the same for every view), no intrinsic guess:
// the principal point starts (and with fixPrincipalPoint stays) at ((w-1)/2, (h-1)/2), the focal lengths come from the
// orthogonality constraints of the view homographies, every view's pose from its homography, then LM over everything.
// imagePoints[v] = n x 2.  Returns the RMS reprojection error (what calibrateCamera returns); < 0 on failure.
inline double calibrateCamera(const std::vector<double> &objectPoints, const std::vector<std::vector<double>> &imagePoints,
                              int width, int height, const CalibFlags &flags, CameraModel &cam,
                              std::vector<std::array<double, 3>> &rvecs, std::vector<std::array<double, 3>> &tvecs) {
    const int n = (int) objectPoints.size() / 3, nv = (int) imagePoints.size();
    if (n < 4 || nv < 1) return -1;
    std::vector<double> oxy((size_t) 2 * n);
    for (int i = 0; i < n; ++i) oxy[(size_t) 2 * i] = objectPoints[(size_t) 3 * i], oxy[(size_t) 2 * i + 1] = objectPoints[(size_t) 3 * i + 1];
    cam = CameraModel();
    cam.cx = (width - 1) * 0.5;
    cam.cy = (height - 1) * 0.5;
    // focal lengths (cvInitIntrinsicParams2D): for H' = T(-c) H with columns h, v: h.v = 0 and |h| = |v| under diag(1/fx^2, 1/fy^2, 1)
    double M[4] = {0, 0, 0, 0}, rh[2] = {0, 0};
    std::vector<std::array<double, 9>> Hs((size_t) nv);
    for (int vi = 0; vi < nv; ++vi) {
        double *H = Hs[(size_t) vi].data();
        if (!findHomography(oxy.data(), imagePoints[(size_t) vi].data(), n, H)) return -1;
        double G[9];
        for (int c = 0; c < 3; ++c) G[c] = H[c] - cam.cx * H[6 + c], G[3 + c] = H[3 + c] - cam.cy * H[6 + c], G[6 + c] = H[6 + c];
        double h[3], v[3], d1[3], d2[3], nh = 0, nvv = 0, n1 = 0, n2 = 0;
        for (int k = 0; k < 3; ++k) {
            h[k] = G[3 * k], v[k] = G[3 * k + 1];
            d1[k] = (h[k] + v[k]) * 0.5, d2[k] = (h[k] - v[k]) * 0.5;
            nh += h[k] * h[k], nvv += v[k] * v[k], n1 += d1[k] * d1[k], n2 += d2[k] * d2[k];
        }
        nh = 1 / std::sqrt(nh), nvv = 1 / std::sqrt(nvv), n1 = 1 / std::sqrt(n1), n2 = 1 / std::sqrt(n2);
        for (int k = 0; k < 3; ++k) h[k] *= nh, v[k] *= nvv, d1[k] *= n1, d2[k] *= n2;
        const double rows[2][3] = {{h[0] * v[0], h[1] * v[1], -h[2] * v[2]}, {d1[0] * d2[0], d1[1] * d2[1], -d1[2] * d2[2]}};
        for (const auto &r : rows) {
            M[0] += r[0] * r[0], M[1] += r[0] * r[1], M[3] += r[1] * r[1];
            rh[0] += r[0] * r[2], rh[1] += r[1] * r[2];
        }
    }
    M[2] = M[1];
    const double det = M[0] * M[3] - M[1] * M[2];
    if (det == 0) return -1;
    const double f0 = (M[3] * rh[0] - M[1] * rh[1]) / det, f1 = (M[0] * rh[1] - M[2] * rh[0]) / det;
    cam.fx = std::sqrt(std::abs(1.0 / f0));
    cam.fy = std::sqrt(std::abs(1.0 / f1));
    if (flags.fixAspectRatio) {
        const double tf = (cam.fx + cam.fy * flags.aspectRatio) * 0.5;
        cam.fx = tf;
        cam.fy = tf / flags.aspectRatio;
    }
    if (!std::isfinite(cam.fx) || !std::isfinite(cam.fy)) return -1;
    // per-view extrinsics from the homography to normalised coordinates (zero distortion at this point)
    std::vector<View> views((size_t) nv);
    for (int vi = 0; vi < nv; ++vi) {
        const double *H = Hs[(size_t) vi].data();
        double Hn[9];
        for (int c = 0; c < 3; ++c) {
            Hn[c] = (H[c] - cam.cx * H[6 + c]) / cam.fx;
            Hn[3 + c] = (H[3 + c] - cam.cy * H[6 + c]) / cam.fy;
            Hn[6 + c] = H[6 + c];
        }
        views[(size_t) vi].img = imagePoints[(size_t) vi];
        poseFromHomography(Hn, views[(size_t) vi].rvec, views[(size_t) vi].tvec);
    }
    {  // each view's pose alone first (cvFindExtrinsicCameraParams2 refines too), then everything
        CalibFlags pf;
        pf.fixIntrinsics = true;
        for (int vi = 0; vi < nv; ++vi) {
            std::vector<View> one(1, views[(size_t) vi]);
            detail::refine(objectPoints, one, cam, pf, 30);
            views[(size_t) vi] = one[0];
        }
    }
    const double sse = detail::refine(objectPoints, views, cam, flags);
    rvecs.resize((size_t) nv);
    tvecs.resize((size_t) nv);
    for (int vi = 0; vi < nv; ++vi)
        for (int k = 0; k < 3; ++k) rvecs[(size_t) vi][(size_t) k] = views[(size_t) vi].rvec[k], tvecs[(size_t) vi][(size_t) k] = views[(size_t) vi].tvec[k];
    return std::sqrt(sse / ((double) n * nv));
}

// EventCalibIni.cpp:115-141: total RMS over all points and the per-view RMS
inline double computeReprojectionErrors(const std::vector<double> &objectPoints, const std::vector<std::vector<double>> &imagePoints,
                                        const std::vector<std::array<double, 3>> &rvecs, const std::vector<std::array<double, 3>> &tvecs,
                                        const CameraModel &cam, std::vector<float> &perViewErrors) {
    const int n = (int) objectPoints.size() / 3;
    perViewErrors.resize(imagePoints.size());
    double total = 0;
    size_t points = 0;
    std::vector<double> img((size_t) 2 * n);
    for (size_t v = 0; v < imagePoints.size(); ++v) {
        projectPoints(objectPoints.data(), n, rvecs[v].data(), tvecs[v].data(), cam, img.data());
        double e2 = 0;
        for (int i = 0; i < 2 * n; ++i) {
            const double d = (double) (float) img[(size_t) i] - imagePoints[v][(size_t) i];  // imagePoints2 is vector<Point2f>
            e2 += d * d;
        }
        perViewErrors[v] = (float) std::sqrt(e2 / n);
        total += e2;
        points += (size_t) n;
    }
    return std::sqrt(total / points);
}

// cv::solvePnPRansac(objectPoints, imagePoints, K, dist, rvec, tvec, false, 50, reprojectionError, 0.99, inliers, SOLVEPNP_IPPE)
// for the planar board: homography of the undistorted normalised points -> pose -> LM on the reprojection error, inliers =
// points within reprojectionError px, re-solved on the inliers until the set is stable.  false when fewer than 4 inliers.
inline bool solvePnPPlanar(const std::vector<double> &objectPoints, const std::vector<double> &imagePoints, const CameraModel &cam,
                           double reprojectionError, double rvec[3], double tvec[3], std::vector<int> &inliers) {
    const int n = (int) objectPoints.size() / 3;
    if (n < 4 || (int) imagePoints.size() != 2 * n) return false;
    std::vector<char> use((size_t) n, 1);
    const double intr[4] = {cam.fx, cam.fy, cam.cx, cam.cy};
    CalibFlags pf;
    pf.fixIntrinsics = true;
    for (int round = 0; round < 5; ++round) {
        std::vector<double> o2, o3, nm, im;
        for (int i = 0; i < n; ++i)
            if (use[(size_t) i]) {
                double xy[2];
                undistortPoint(cam, &imagePoints[(size_t) 2 * i], xy);
                o2.insert(o2.end(), {objectPoints[(size_t) 3 * i], objectPoints[(size_t) 3 * i + 1]});
                o3.insert(o3.end(), {objectPoints[(size_t) 3 * i], objectPoints[(size_t) 3 * i + 1], objectPoints[(size_t) 3 * i + 2]});
                nm.insert(nm.end(), {xy[0], xy[1]});
                im.insert(im.end(), {imagePoints[(size_t) 2 * i], imagePoints[(size_t) 2 * i + 1]});
            }
        const int m = (int) o2.size() / 2;
        if (m < 4) return false;
        double H[9];
        if (!findHomography(o2.data(), nm.data(), m, H)) return false;
        std::vector<View> one(1);
        one[0].img = im;
        poseFromHomography(H, one[0].rvec, one[0].tvec);
        CameraModel c = cam;
        detail::refine(o3, one, c, pf, 100);
        for (int k = 0; k < 3; ++k) rvec[k] = one[0].rvec[k], tvec[k] = one[0].tvec[k];
        bool changed = false;
        for (int i = 0; i < n; ++i) {
            double uv[2];
            projectPoint<double>(intr, cam.dist, rvec, tvec, &objectPoints[(size_t) 3 * i], uv);
            const double ex = uv[0] - imagePoints[(size_t) 2 * i], ey = uv[1] - imagePoints[(size_t) 2 * i + 1];
            const char in = ex * ex + ey * ey <= reprojectionError * reprojectionError;
            if (in != use[(size_t) i]) changed = true;
            use[(size_t) i] = in;
        }
        if (!changed) break;
    }
    inliers.clear();
    for (int i = 0; i < n; ++i)
        if (use[(size_t) i]) inliers.push_back(i);
    return inliers.size() >= 4;
}

// body pose of a key frame from the PnP result (EventCalibIni.cpp:260-275; identity sensor extrinsics):
// Qwb = Qsw^-1 (x y z w), twb = -Rsw^T tsw
inline void bodyPoseFromPnP(const double rvec[3], const double tvec[3], double unitQwb[4], double twb[3]) {
    double R[9], q[4];
    rodrigues<double>(rvec, R);
    rotationToQuaternion(R, q);
    const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    unitQwb[0] = -q[0] / n, unitQwb[1] = -q[1] / n, unitQwb[2] = -q[2] / n, unitQwb[3] = q[3] / n;
    for (int k = 0; k < 3; ++k) twb[k] = -(R[k] * tvec[0] + R[3 + k] * tvec[1] + R[6 + k] * tvec[2]);
}

// EventCalibIni::checkPose (:328-346): linear and angular speed against the last key frame of the map
inline bool checkPose(double refStamp, const double refQwb[4], const double refTwb[3], double curStamp, const double curQwb[4],
                      const double curTwb[3], double motionTimeStep) {
    const double duration = curStamp - refStamp;
    // the rotation by refQwb^-1 keeps the norm
    const double dt[3] = {curTwb[0] - refTwb[0], curTwb[1] - refTwb[1], curTwb[2] - refTwb[2]};
    const double v_t = std::sqrt(dt[0] * dt[0] + dt[1] * dt[1] + dt[2] * dt[2]) / duration;
    // Eigen angularDistance: 2 atan2(|vec(d)|, |w(d)|), d = cur * ref^-1
    const double *a = curQwb, *b = refQwb;
    const double w = a[3] * b[3] + a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
    const double x = -a[3] * b[0] + a[0] * b[3] - a[1] * b[2] + a[2] * b[1];
    const double y = -a[3] * b[1] + a[1] * b[3] - a[2] * b[0] + a[0] * b[2];
    const double z = -a[3] * b[2] + a[2] * b[3] - a[0] * b[1] + a[1] * b[0];
    const double v_R = std::abs(2 * std::atan2(std::sqrt(x * x + y * y + z * z), std::abs(w)) / duration);
    return v_t < (2.5e-1 / motionTimeStep) * 2 && v_R < (5e-4 * M_PI) * 2 / motionTimeStep;
}

}  // namespace ecb
#endif  // ECB_CALIB_INIT_HPP
