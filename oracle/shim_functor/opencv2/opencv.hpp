// TEST INFRASTRUCTURE.  Stand-in for <opencv2/opencv.hpp> (external, absent).  EventFrame.cpp / CirclesEventFrame.cpp paint
// debug images (kept as sparse pictures / no-ops); the three OpenCV calls with an effect on the results are HOOKS the test
// wrapper fills in (oracle/ref_functor_capi.cpp): findCirclesGrid (grid order of the candidate centres), projectPoints (the
// 5 image points per board circle of rectifyFeatures) — eigen2cv / Rodrigues only shuttle data to projectPoints and are no-ops.
#ifndef ECB_ORACLE_OPENCV_SHIM
#define ECB_ORACLE_OPENCV_SHIM
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>
#define CV_8UC3 16
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
class Mat {  // a sparse picture: enough for at<Vec3b>(Point) = colour
public:
    Mat() {}
    explicit Mat(const std::vector<Point2f> &) {}
    Mat(int rows_, int cols_, int, const Vec3b & = Vec3b()) : rows(rows_), cols(cols_), px(std::make_shared<std::map<std::pair<int, int>, Vec3b>>()) {}
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
template <class E> inline void eigen2cv(const E &, Mat &) {}
inline void Rodrigues(const Mat &, Mat &) {}
// HOOK: defined by the test wrapper
void projectPoints(const std::vector<Point3f> &objectPoints, const Mat &rvec, const Mat &tvec, const Mat &cameraMatrix,
                   const Mat &distCoeffs, std::vector<Point2f> &imagePoints);
}  // namespace cv
#endif
