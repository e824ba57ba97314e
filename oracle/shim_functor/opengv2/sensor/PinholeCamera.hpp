// TEST INFRASTRUCTURE.  Stand-in for core/sensor/include/opengv2/sensor/PinholeCamera.hpp: K, the 5 OpenCV distortion
// coefficients and the 5-term inverse radial polynomial.  inverseRadialDistortion restates core/sensor/src/PinholeCamera.cpp:70-95
// (that file needs OpenCV); it is pinned by the reference's own known answer (unit_test_inverseDistortion, tests/test_oracle_cost.py).
#ifndef ECB_ORACLE_PINHOLECAMERA_SHIM
#define ECB_ORACLE_PINHOLECAMERA_SHIM
#include <opengv2/sensor/CameraBase.hpp>
namespace opengv2 {
class PinholeCamera : public CameraBase {
public:
    explicit PinholeCamera(const Eigen::Vector2d &size) : CameraBase(size), dist_(5) {}
    const Eigen::Matrix3d &K() const { return K_; }
    void setK(const Eigen::Matrix3d &K) { K_ = K; }
    Eigen::VectorXd &distCoeffs() { return dist_; }
    const Eigen::VectorXd &distCoeffs() const { return dist_; }
    Eigen::Matrix<double, 5, 1> &inverseRadialPoly() { return inv_; }
    static Eigen::VectorXd inverseRadialDistortion(const Eigen::Vector4d &k) {
        const double k00 = k[0] * k[0], k000 = k[0] * k00, k0000 = k[0] * k000, k00000 = k[0] * k0000;
        const double k01 = k[0] * k[1], k001 = k[0] * k01, k0001 = k[0] * k001, k11 = k[1] * k[1], k011 = k[0] * k11;
        const double k02 = k[0] * k[2], k002 = k[0] * k02, k12 = k[1] * k[2], k03 = k[0] * k[3];
        Eigen::VectorXd b(5);
        b[0] = -k[0];
        b[1] = 3 * k00 - k[1];
        b[2] = -12 * k000 + 8 * k01 - k[2];
        b[3] = 55 * k0000 - 55 * k001 + 5 * k11 + 10 * k02 - k[3];
        b[4] = -273 * k00000 + 364 * k0001 - 78 * k011 - 78 * k002 + 12 * k12 + 12 * k03;
        return b;
    }

private:
    Eigen::Matrix3d K_;
    Eigen::VectorXd dist_;
    Eigen::Matrix<double, 5, 1> inv_;
};
}  // namespace opengv2
#endif
