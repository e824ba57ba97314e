// TEST INFRASTRUCTURE.  Stand-in for <opencv2/opencv.hpp> (external, absent).  EventFrame.cpp / CirclesEventFrame.cpp paint
// debug images (kept as sparse pictures / no-ops); the three OpenCV calls with an effect on the results are HOOKS the test
// wrapper fills in (oracle/ref_functor_capi.cpp): findCirclesGrid (grid order of the candidate centres), projectPoints (the
// 5 image points per board circle of rectifyFeatures) — eigen2cv / Rodrigues only shuttle data to projectPoints and are no-ops.
#ifndef ECB_ORACLE_OPENCV_SHIM
#define ECB_ORACLE_OPENCV_SHIM
#include <cmath>
#include <cstdint>
#include <Eigen/Eigen>
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>
#define CV_8UC3 16
#define CV_64F 6
#define CV_32F 5
namespace cv {
struct Vec3b {
    uint8_t v[3];
    Vec3b() : v{0, 0, 0} {}
    Vec3b(int a, int b, int c) : v{(uint8_t) a, (uint8_t) b, (uint8_t) c} {}
};
struct Point {
    int x, y;
    Point() : x(0), y(0) {}
    Point(double x_, double y_) : x((int) x_), y((int) y_) {}
};
struct Point2f {
    float x, y;
    Point2f() : x(0), y(0) {}
    Point2f(double x_, double y_) : x((float) x_), y((float) y_) {}
};
struct Point3f {
    float x, y, z;
    Point3f() : x(0), y(0), z(0) {}
    Point3f(double x_, double y_, double z_) : x((float) x_), y((float) y_), z((float) z_) {}
};
struct Size {
    int width, height;
    Size(int w = 0, int h = 0) : width(w), height(h) {}
};
enum { CALIB_CB_SYMMETRIC_GRID = 1, CALIB_CB_ASYMMETRIC_GRID = 2, CALIB_CB_CLUSTERING = 4 };
enum { NORM_L2 = 4, SOLVEPNP_ITERATIVE = 0, SOLVEPNP_IPPE = 6, CALIB_USE_LU = (1 << 17) };
enum { CALIB_FIX_ASPECT_RATIO = 2, CALIB_FIX_PRINCIPAL_POINT = 4, CALIB_ZERO_TANGENT_DIST = 8, CALIB_FIX_K1 = 32, CALIB_FIX_K2 = 64,
       CALIB_FIX_K3 = 128, CALIB_FIX_K4 = 2048, CALIB_FIX_K5 = 4096, CALIB_FIX_K6 = 8192 };
namespace fisheye {
enum { CALIB_RECOMPUTE_EXTRINSIC = 2, CALIB_FIX_SKEW = 8, CALIB_FIX_K1 = 16, CALIB_FIX_K2 = 32, CALIB_FIX_K3 = 64, CALIB_FIX_K4 = 128,
       CALIB_FIX_PRINCIPAL_POINT = 512 };
}
// cv::FileStorage as a key -> number table: the reference's own parameter constructors read their fields from it
struct FileNode {
    double value;
    bool present;
};
template <class T> inline void operator>>(const FileNode &n, T &v) {
    if (n.present) v = (T) n.value;
}
struct FileStorage {
    std::map<std::string, double> kv;
    FileNode operator[](const char *k) const {
        const auto it = kv.find(k);
        return it == kv.end() ? FileNode{0.0, false} : FileNode{it->second, true};
    }
};
class Mat {  // a sparse picture (at<Vec3b>(Point) = colour) or a small dense matrix of doubles (at<double>(i, j))
public:
    Mat() {}
    static Mat eye(int r, int c, int) {
        Mat m = zeros(r, c, 0);
        for (int i = 0; i < r && i < c; ++i) m.d[(size_t) (i * c + i)] = 1.0;
        return m;
    }
    static Mat zeros(int r, int c, int) {
        Mat m;
        m.rows = r, m.cols = c;
        m.d.assign((size_t) (r * c), 0.0);
        return m;
    }
    Mat row(int i) const {
        Mat m = zeros(1, cols, 0);
        for (int j = 0; j < cols; ++j) m.d[(size_t) j] = d[(size_t) (i * cols + j)];
        return m;
    }
    bool empty() const { return d.empty(); }
    std::vector<double> d;  // row-major
    explicit Mat(const std::vector<Point2f> &) {}
    Mat(int rows_, int cols_, int, const Vec3b & = Vec3b()) : rows(rows_), cols(cols_), px(std::make_shared<std::map<std::pair<int, int>, Vec3b>>()) {}
    template <class T> T &at(int i, int j) { return d[(size_t) (i * cols + j)]; }
    template <class T> const T &at(int i, int j) const { return d[(size_t) (i * cols + j)]; }
    template <class T> T &at(const Point &p) {
        if (!px) px = std::make_shared<std::map<std::pair<int, int>, Vec3b>>();
        return (*px)[std::make_pair(p.y, p.x)];
    }
    Mat clone() const {
        Mat m = *this;
        if (px) m.px = std::make_shared<std::map<std::pair<int, int>, Vec3b>>(*px);
        return m;
    }
    int rows = 0, cols = 0;

private:
    std::shared_ptr<std::map<std::pair<int, int>, Vec3b>> px;
};
inline void circle(Mat &, const Point &, double, const Vec3b &) {}
inline void drawChessboardCorners(Mat &, Size, const Mat &, bool) {}
// Eigen <-> Mat shuttles (dense doubles)
template <class T, int N> inline void eigen2cv(const Eigen::Matrix<T, N, 1> &v, Mat &m) {
    m = Mat::zeros(N, 1, 0);
    for (int i = 0; i < N; ++i) m.d[(size_t) i] = v[i];
}
template <class T> inline void eigen2cv(const Eigen::Matrix<T, 3, 3> &a, Mat &m) {
    m = Mat::zeros(3, 3, 0);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) m.d[(size_t) (3 * i + j)] = a(i, j);
}
inline void eigen2cv(const Eigen::VectorXd &v, Mat &m) {
    m = Mat::zeros((int) v.size(), 1, 0);
    for (long i = 0; i < v.size(); ++i) m.d[(size_t) i] = v[i];
}
inline void cv2eigen(const Mat &m, Eigen::Matrix<double, 3, 3> &a) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) a(i, j) = m.d[(size_t) (3 * i + j)];
}
inline void cv2eigen(const Mat &m, Eigen::Matrix<double, 3, 1> &v) {
    for (int i = 0; i < 3; ++i) v[i] = m.d[(size_t) i];
}
inline void cv2eigen(const Mat &m, Eigen::VectorXd &v) {
    v = Eigen::VectorXd((long) m.d.size());
    for (size_t i = 0; i < m.d.size(); ++i) v[(long) i] = m.d[i];
}
inline bool checkRange(const Mat &m) {
    for (double x : m.d)
        if (!(x == x) || x > 1e300 || x < -1e300) return false;
    return true;
}
inline double norm(const std::vector<Point2f> &a, const std::vector<Point2f> &b, int) {
    double s = 0;
    for (size_t i = 0; i < a.size(); ++i) {
        const double dx = (double) a[i].x - (double) b[i].x, dy = (double) a[i].y - (double) b[i].y;
        s += dx * dx + dy * dy;
    }
    return std::sqrt(s);
}
// HOOKS, defined by the test wrapper (oracle/ref_functor_capi.cpp) on top of the product's host header include/ecb/calib_init.hpp
void Rodrigues(const Mat &src, Mat &dst);  // 3-vector -> 3x3 matrix, 3x3 matrix -> 3-vector
void projectPoints(const std::vector<Point3f> &objectPoints, const Mat &rvec, const Mat &tvec, const Mat &cameraMatrix,
                   const Mat &distCoeffs, std::vector<Point2f> &imagePoints);
double calibrateCamera(const std::vector<std::vector<Point3f>> &objectPoints, const std::vector<std::vector<Point2f>> &imagePoints,
                       Size imageSize, Mat &cameraMatrix, Mat &distCoeffs, std::vector<Mat> &rvecs, std::vector<Mat> &tvecs, int flags);
bool solvePnPRansac(const std::vector<Point3f> &objectPoints, const std::vector<Point2f> &imagePoints, const Mat &cameraMatrix,
                    const Mat &distCoeffs, Mat &rvec, Mat &tvec, bool useExtrinsicGuess, int iterationsCount, float reprojectionError,
                    double confidence, std::vector<int> &inliers, int flags);
namespace fisheye {  // Calibrate_UseFisheyeModel: 1 is outside this build; the calls only have to compile
inline double calibrate(const std::vector<std::vector<Point3f>> &, const std::vector<std::vector<Point2f>> &, Size, Mat &, Mat &, Mat &,
                        Mat &, int) {
    return -1.0;
}
inline void projectPoints(const std::vector<Point3f> &, std::vector<Point2f> &, const Mat &, const Mat &, const Mat &, const Mat &) {}
}  // namespace fisheye
}  // namespace cv
#endif
