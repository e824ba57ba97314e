// TEST INFRASTRUCTURE.  Stand-in for <opencv2/opencv.hpp> (external, absent): EventFrame.cpp only paints debug images.
#ifndef ECB_ORACLE_OPENCV_SHIM
#define ECB_ORACLE_OPENCV_SHIM
#include <cstdint>
#include <map>
#include <memory>
#include <utility>
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
class Mat {  // a sparse picture: enough for at<Vec3b>(Point) = colour
public:
    Mat() {}
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
}  // namespace cv
#endif
