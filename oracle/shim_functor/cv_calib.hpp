// TEST INFRASTRUCTURE.  Stand-in for camera_calibration/cv_calib/include/cv_calib.hpp (the reference's points-in
// findCirclesGrid, vendored OpenCV code that needs OpenCV itself): a HOOK the test wrapper defines — it records the candidate
// centres it is handed and orders them with include/ecb/circles_grid.hpp (the product's grid finder).
#ifndef ECB_ORACLE_CV_CALIB_SHIM
#define ECB_ORACLE_CV_CALIB_SHIM
#include <opencv2/opencv.hpp>
namespace cv {
bool findCirclesGrid(const std::vector<Point2f> &points_, Size patternSize, std::vector<Point2f> &centers, int flags);
}
#endif
