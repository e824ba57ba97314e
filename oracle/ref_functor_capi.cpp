// TEST INFRASTRUCTURE — C wrapper around the reference's OWN residual functor and B-spline, compiled where they lie under
// /root/reference (never copied) into oracle/_ref/libref_functor.so by oracle/Makefile:
//
//   opengv2::EventCalibSpline::CalibReprojectionError::operator()<T>   event_camera_calib/include/opengv2/event_camera_calib/
//   opengv2::EventCalibSpline::unDistort<T>                            EventCalibSpline.hpp:36-63,158-250
//   opengv2::BsplineReal<dim> (ctor fit, findSpan, dersBasisFuns, evaluate)   core/spline/include/opengv2/spline/BsplineReal.hpp
//   opengv2::EventFrame::EventFrame (window, per-polarity pixel sets, +/- cancellation, hash-set order)
//                                                                      camera_calibration/event/src/EventFrame.cpp:10-36 with
//                                                                      EigenMatrixHash (core/utility/.../utility.hpp:38-51)
//   opengv2::Event operator>> (the 25-byte record reader)              camera_calibration/event/include/opengv2/event/Event.hpp:41-47
//   opengv2::CirclesEventFrame (ctor, extractFeatures, fitCircle, rectifyFeatures, findCenter) with the reference's DBSCAN
//                                                                      camera_calibration/event_camera_calib/src/CirclesEventFrame.cpp,
//                                                                      .../include/opengv2/event_camera_calib/CirclesEventFrame.hpp,
//                                                                      camera_calibration/dbscan/{include/dbscan.h,src/kdtree.cpp}
//   opengv2::EventCalibSpline (constructor: reduceMap segmentation, time extension, cpNum rule, spline fits, intrinsics with the
//   inverse radial polynomial; optimize(): association loop, Ceres problem assembly; updateMap())
//                                                                      camera_calibration/event_camera_calib/src/EventCalibSpline.cpp
//   opengv2::EventCalibIni (track = the tracking gate, checkPose, cvCalibration)
//                                                                      camera_calibration/event_camera_calib/src/EventCalibIni.cpp
//     hooks: calibrateCamera / solvePnPRansac / projectPoints / Rodrigues = the product's include/ecb/calib_init.hpp, so the
//     reference's own control flow (frame selection, pose conversion, checkPose chain, rectifyFeatures, counters) runs on the
//     product's numerics; JacobiSVD of the row fits is restated in the stand-in Eigen
//     ceres::Problem records what the reference adds and ceres::Solve is a no-op hook (Ceres is absent): the ASSEMBLED problem
//     (residual list, parameter blocks, loss, solver options) is what is compared, not the solve
//     hooks (OpenCV is absent): findCirclesGrid = the product's grid finder (include/ecb/circles_grid.hpp) on the candidate
//     centres the reference hands over; projectPoints = the 5 image points per board circle supplied by the caller
//
// against the stand-in headers of oracle/shim_functor/ (Eigen / Ceres / Sophus are external and absent: what the two reference
// headers need from Eigen is restated there; nothing of the functor's or the spline's own arithmetic is).  T = double gives the
// residual, T = Jet<37> (ceres/jet.h semantics: value + 37 partials, restated below like in ecb_oracle_cost.cpp) gives the
// 1 x 37 ambient Jacobian Ceres' AutoDiffCostFunction would produce (EventCalibSpline.hpp:233-239).
// Only tests/ use this library: it pins the oracle's restatement (ecb_oracle_cost.cpp) to the reference's source text.
#include <cmath>
#include <cstring>
#include <memory>
#include <vector>

namespace jet {
template <int NP>
struct Jet {
    double a;
    double v[NP];
    Jet() : a(0) { std::memset(v, 0, sizeof v); }
    Jet(double s) : a(s) { std::memset(v, 0, sizeof v); }  // NOLINT
};
#define JET_BIN(op, expr_a, expr_v)                                              \
    template <int NP> Jet<NP> operator op(const Jet<NP> &f, const Jet<NP> &g) { \
        Jet<NP> h;                                                               \
        h.a = expr_a;                                                            \
        for (int i = 0; i < NP; ++i) h.v[i] = expr_v;                            \
        return h;                                                                \
    }
JET_BIN(+, f.a + g.a, f.v[i] + g.v[i])
JET_BIN(-, f.a - g.a, f.v[i] - g.v[i])
JET_BIN(*, f.a *g.a, f.a *g.v[i] + f.v[i] * g.a)
template <int NP> Jet<NP> operator/(const Jet<NP> &f, const Jet<NP> &g) {  // ceres/jet.h: g_a_inverse, f_a_by_g_a
    Jet<NP> h;
    const double gi = 1.0 / g.a, fg = f.a * gi;
    h.a = fg;
    for (int i = 0; i < NP; ++i) h.v[i] = (f.v[i] - fg * g.v[i]) * gi;
    return h;
}
template <int NP> Jet<NP> operator*(double s, const Jet<NP> &f) {
    Jet<NP> h;
    h.a = s * f.a;
    for (int i = 0; i < NP; ++i) h.v[i] = s * f.v[i];
    return h;
}
template <int NP> Jet<NP> operator*(const Jet<NP> &f, double s) { return s * f; }
template <int NP> Jet<NP> operator-(const Jet<NP> &f) {
    Jet<NP> h;
    h.a = -f.a;
    for (int i = 0; i < NP; ++i) h.v[i] = -f.v[i];
    return h;
}
template <int NP> Jet<NP> operator-(double s, const Jet<NP> &f) { return Jet<NP>(s) - f; }
template <int NP> Jet<NP> operator-(const Jet<NP> &f, double s) { return f - Jet<NP>(s); }
template <int NP> Jet<NP> operator+(double s, const Jet<NP> &f) { return Jet<NP>(s) + f; }
template <int NP> Jet<NP> operator+(const Jet<NP> &f, double s) { return Jet<NP>(s) + f; }
template <int NP> Jet<NP> &operator*=(Jet<NP> &f, const Jet<NP> &g) { return f = f * g; }
template <int NP> Jet<NP> &operator+=(Jet<NP> &f, const Jet<NP> &g) { return f = f + g; }
template <int NP> bool operator>(const Jet<NP> &f, const Jet<NP> &g) { return f.a > g.a; }
template <int NP> bool operator!=(const Jet<NP> &f, const Jet<NP> &g) { return f.a != g.a; }
template <int NP> bool operator<(const Jet<NP> &f, const Jet<NP> &g) { return f.a < g.a; }
template <int NP> Jet<NP> operator/(const Jet<NP> &f, double s) { return f / Jet<NP>(s); }
template <int NP> Jet<NP> operator/(double s, const Jet<NP> &f) { return Jet<NP>(s) / f; }
template <int NP> Jet<NP> &operator/=(Jet<NP> &f, const Jet<NP> &g) { return f = f / g; }
// ceres/jet.h: abs, sin, cos, atan of a dual number
template <int NP> Jet<NP> abs(const Jet<NP> &f) { return f.a < 0.0 ? -f : f; }
template <int NP> Jet<NP> sin(const Jet<NP> &f) {
    Jet<NP> h;
    h.a = std::sin(f.a);
    const double c = std::cos(f.a);
    for (int i = 0; i < NP; ++i) h.v[i] = c * f.v[i];
    return h;
}
template <int NP> Jet<NP> cos(const Jet<NP> &f) {
    Jet<NP> h;
    h.a = std::cos(f.a);
    const double s = -std::sin(f.a);
    for (int i = 0; i < NP; ++i) h.v[i] = s * f.v[i];
    return h;
}
template <int NP> Jet<NP> atan(const Jet<NP> &f) {
    Jet<NP> h;
    h.a = std::atan(f.a);
    const double t = 1.0 / (1.0 + f.a * f.a);
    for (int i = 0; i < NP; ++i) h.v[i] = t * f.v[i];
    return h;
}
template <int NP> Jet<NP> sqrt(const Jet<NP> &f) {  // ceres/jet.h
    Jet<NP> h;
    h.a = std::sqrt(f.a);
    const double t = 1.0 / (2.0 * h.a);
    for (int i = 0; i < NP; ++i) h.v[i] = f.v[i] * t;
    return h;
}
}  // namespace jet

#include <opengv2/event_camera_calib/EventCalibSpline.hpp>  // the reference's header, resolved through -I by the Makefile

using opengv2::EventCalibSpline;
typedef EventCalibSpline::CalibReprojectionError Functor;

namespace {
template <class T>
T run_functor(const T *intr, const T *rcp16, const T *tcp12, const double *obs, const double *lm, double radius, const double *basis) {
    const Eigen::Vector2d o(obs[0], obs[1]);
    const Eigen::Vector3d l(lm[0], lm[1], lm[2]);
    const Eigen::Quaterniond Qbs(1, 0, 0, 0);
    const Eigen::Vector3d tbs(0, 0, 0);
    auto rB = std::make_shared<std::vector<std::vector<double>>>(1, std::vector<double>(basis, basis + 4));
    auto tB = std::make_shared<std::vector<std::vector<double>>>(1, std::vector<double>(basis, basis + 4));
    Functor f(&o, &l, &radius, &Qbs, &tbs, rB, tB);
    T res;
    f(intr, rcp16, rcp16 + 4, rcp16 + 8, rcp16 + 12, tcp12, tcp12 + 3, tcp12 + 6, tcp12 + 9, &res);
    return res;
}
}  // namespace

extern "C" {

// value of the reference functor with T = double
double ref_residual(const double *intr, const double *rcp16, const double *tcp12, const double *obs, const double *lm, double radius,
                    const double *basis) {
    return run_functor<double>(intr, rcp16, tcp12, obs, lm, radius, basis);
}

// value and 1 x 37 ambient Jacobian (9 intrinsics | 4 x 4 rotation control points x y z w | 4 x 3 translation control points)
double ref_residual_jac(const double *intr, const double *rcp16, const double *tcp12, const double *obs, const double *lm,
                        double radius, const double *basis, double *jac37) {
    typedef jet::Jet<37> J;
    J p[37];
    for (int k = 0; k < 9; ++k) p[k] = J(intr[k]);
    for (int k = 0; k < 16; ++k) p[9 + k] = J(rcp16[k]);
    for (int k = 0; k < 12; ++k) p[25 + k] = J(tcp12[k]);
    for (int k = 0; k < 37; ++k) p[k].v[k] = 1.0;
    const J r = run_functor<J>(p, p + 9, p + 25, obs, lm, radius, basis);
    for (int k = 0; k < 37; ++k) jac37[k] = r.v[k];
    return r.a;
}

// EventCalibSpline::unDistort<double>
void ref_undistort(const double *intr, const double *obs, double *Xc3) {
    Sophus::Vector3<double> Xc;
    EventCalibSpline::unDistort<double>(intr[0], intr[1], intr[2], intr[3], intr[4], intr[5], intr[6], intr[7], intr[8],
                                        Eigen::Vector2d(obs[0], obs[1]), Xc);
    for (int k = 0; k < 3; ++k) Xc3[k] = Xc[k];
}
}

// ---- a11: the SO(3) variant (useSO3: 1) — CalibReprojectionError_SO3 (EventCalibSpline.hpp:65-156), the cumulative basis of
// BsplineSO3::derBasisFuns (core/spline/src/BsplineSO3.cpp:73-109, compiled where it lies) and LocalParameterizationSO3
// (core/spline/include/opengv2/spline/BsplineSO3.hpp:190-221), against the stand-in Sophus of shim_functor/sophus/so3.hpp ----
namespace {
template <class T>
T run_functor_so3(const T *intr, const T *rcp16, const T *tcp12, const double *obs, const double *lm, double radius,
                  const double *beta3, const double *basis4) {
    const Eigen::Vector2d o(obs[0], obs[1]);
    const Eigen::Vector3d l(lm[0], lm[1], lm[2]);
    const Eigen::Quaterniond Qbs(1, 0, 0, 0);
    const Eigen::Vector3d tbs(0, 0, 0);
    auto rB = std::make_shared<std::vector<std::vector<double>>>(1, std::vector<double>(beta3, beta3 + 3));
    auto tB = std::make_shared<std::vector<std::vector<double>>>(1, std::vector<double>(basis4, basis4 + 4));
    EventCalibSpline::CalibReprojectionError_SO3 f(&o, &l, &radius, &Qbs, &tbs, rB, tB);
    T res;
    f(intr, rcp16, rcp16 + 4, rcp16 + 8, rcp16 + 12, tcp12, tcp12 + 3, tcp12 + 6, tcp12 + 9, &res);
    return res;
}
struct SO3SplineProbe : opengv2::BsplineSO3 {
    SO3SplineProbe() : opengv2::BsplineSO3(3) {}
    void setKnots(const double *k, int nk) { knotVector_.assign(k, k + nk); }
};
}  // namespace

extern "C" {
double ref_residual_so3(const double *intr, const double *rcp16, const double *tcp12, const double *obs, const double *lm,
                        double radius, const double *beta3, const double *basis4) {
    return run_functor_so3<double>(intr, rcp16, tcp12, obs, lm, radius, beta3, basis4);
}
// value and 1 x 37 ambient Jacobian (9 intrinsics | 4 x 4 SO3 control-point coefficients x y z w | 4 x 3 translation control points)
double ref_residual_jac_so3(const double *intr, const double *rcp16, const double *tcp12, const double *obs, const double *lm,
                            double radius, const double *beta3, const double *basis4, double *jac37) {
    typedef jet::Jet<37> J;
    J p[37];
    for (int k = 0; k < 9; ++k) p[k] = J(intr[k]);
    for (int k = 0; k < 16; ++k) p[9 + k] = J(rcp16[k]);
    for (int k = 0; k < 12; ++k) p[25 + k] = J(tcp12[k]);
    for (int k = 0; k < 37; ++k) p[k].v[k] = 1.0;
    const J r = run_functor_so3<J>(p, p + 9, p + 25, obs, lm, radius, beta3, basis4);
    for (int k = 0; k < 37; ++k) jac37[k] = r.v[k];
    return r.a;
}
// BsplineSO3::findSpan + derBasisFuns(u, span, 0) on a given knot vector: the 3 cumulative basis values beta_{k,k-p+j}
void ref_so3_basis(const double *knots, int nk, double u, int *span, double *beta3) {
    SO3SplineProbe sp;
    sp.setKnots(knots, nk);
    const size_t s = sp.findSpan(u);
    std::vector<std::vector<double>> ders;
    sp.derBasisFuns(u, s, 0, ders);
    *span = (int) s;
    for (int j = 0; j < 3; ++j) beta3[j] = ders[0][(size_t) j];
}
// LocalParameterizationSO3::Plus (T * exp(delta)) and ::ComputeJacobian (4 x 3, row major)
void ref_so3_plus(const double *x4, const double *delta3, double *out4) {
    opengv2::LocalParameterizationSO3 lp;
    lp.Plus(x4, delta3, out4);
}
void ref_so3_plus_jacobian(const double *x4, double *J12) {
    opengv2::LocalParameterizationSO3 lp;
    lp.ComputeJacobian(x4, J12);
}
int ref_so3_sizes(void) {
    opengv2::LocalParameterizationSO3 lp;
    return lp.GlobalSize() * 10 + lp.LocalSize();
}
}

// ---- BsplineReal ----
namespace {
template <int dim>
struct Probe : opengv2::BsplineReal<dim> {
    typedef Eigen::Matrix<double, dim, 1> V;
    typedef std::vector<V, Eigen::aligned_allocator<V>> VV;
    Probe() : opengv2::BsplineReal<dim>(3) {}
    Probe(const VV &Q, int cpNum, const std::vector<double> &u) : opengv2::BsplineReal<dim>(3, Q, cpNum, u) {}
    void setKnots(const double *k, int nk) { this->knotVector.assign(k, k + nk); }
};
template <int dim>
int fit(const double *us, const double *data, int n, int cp_num, double *knots, double *cp) {
    typename Probe<dim>::VV Q((size_t) n);
    for (int i = 0; i < n; ++i)
        for (int c = 0; c < dim; ++c) Q[(size_t) i][c] = data[(size_t) i * dim + c];
    Probe<dim> sp(Q, cp_num, std::vector<double>(us, us + n));
    const std::vector<double> &kv = sp.getKnotVector();
    for (size_t i = 0; i < kv.size(); ++i) knots[i] = kv[i];
    auto &C = sp.getCP();
    for (size_t i = 0; i < C.size(); ++i)
        for (int c = 0; c < dim; ++c) cp[i * dim + (size_t) c] = C[i][c];
    return (int) C.size();
}
}  // namespace

extern "C" {
// BsplineReal<dim>(degree 3, samples, cpNum, timestamps): knot vector (cpNum + 4) and control points (cpNum x dim); returns the
// number of control points the reference produced (0 when its fit failed)
int ref_spline_fit(int dim, const double *us, const double *data, int n, int cp_num, double *knots, double *cp) {
    return dim == 3 ? fit<3>(us, data, n, cp_num, knots, cp) : dim == 4 ? fit<4>(us, data, n, cp_num, knots, cp) : -1;
}
// findSpan + dersBasisFuns(u, span, 0) on a given knot vector
void ref_basis(const double *knots, int nk, double u, int *span, double *N4) {
    Probe<3> sp;
    sp.setKnots(knots, nk);
    const size_t s = sp.findSpan(u);
    std::vector<std::vector<double>> ders;
    sp.dersBasisFuns(u, s, 0, ders);
    *span = (int) s;
    for (int j = 0; j < 4; ++j) N4[j] = ders[0][(size_t) j];
}
}

// ---- EventFrame (a2) and the record reader (a1) ----
#include <fstream>
#include <unordered_set>

#include <opengv2/event/Event.hpp>
#include <opengv2/event/EventFrame.hpp>

namespace {
struct FrameProbe : opengv2::EventFrame {
    FrameProbe(opengv2::EventContainer::Ptr c, const std::pair<double, double> &d) : opengv2::EventFrame(c, d) {}
    bool rectifyFeatures(const std::unordered_set<int> &, const Eigen::Ref<const Eigen::Matrix3d> &,
                         const Eigen::Ref<const Eigen::Vector3d> &) override {
        return true;
    }
    const opengv2::vectorofEigenMatrix<Eigen::Vector2d> &pos() const { return positiveEvents_; }
    const opengv2::vectorofEigenMatrix<Eigen::Vector2d> &neg() const { return negativeEvents_; }
};
}  // namespace

extern "C" {
// The reference's EventFrame constructor over events (t, x, y, polarity) inserted like its load loop does
// (event_camera_calib/test/eventCameraCalib.cpp:158-160): closed window [t0, t1]; pixel lists in the hash sets' iteration order
void ref_event_frame(const double *t, const double *x, const double *y, const unsigned char *pol, long long n, double t0, double t1,
                     double *pos_xy, double *neg_xy, long long *n_pos, long long *n_neg) {
    auto container = std::make_shared<opengv2::EventContainer>();
    container->camera = std::make_shared<opengv2::CameraBase>(Eigen::Vector2d(346, 260));
    for (long long i = 0; i < n; ++i)
        container->container.emplace(t[i], opengv2::Event_loc_pol(Eigen::Vector2d(x[i], y[i]), pol[i] != 0));
    FrameProbe f(container, std::make_pair(t0, t1));
    *n_pos = (long long) f.pos().size();
    *n_neg = (long long) f.neg().size();
    for (size_t i = 0; i < f.pos().size(); ++i) pos_xy[2 * i] = f.pos()[i][0], pos_xy[2 * i + 1] = f.pos()[i][1];
    for (size_t i = 0; i < f.neg().size(); ++i) neg_xy[2 * i] = f.neg()[i][0], neg_xy[2 * i + 1] = f.neg()[i][1];
}

// the reference's record reader: Event's operator>> until the stream fails; returns the number of records read
long long ref_read_bin(const char *path, long long cap, double *t, double *x, double *y, unsigned char *pol) {
    std::ifstream is(path, std::ifstream::binary | std::ifstream::in);
    long long n = 0;
    opengv2::Event e;
    while (n < cap && (is >> e) && is.good()) {
        t[n] = e.timeStamp();
        x[n] = e.location()[0];
        y[n] = e.location()[1];
        pol[n] = e.polarity() ? 1 : 0;
        ++n;
    }
    return n;
}
}

// ---- CirclesEventFrame (a4 extractFeatures, a5 fitCircle, a6 rectifyFeatures, a7 findCenter) ----
#include <cv_calib.hpp>
#include <opengv2/event_camera_calib/CirclesEventFrame.hpp>
#include <opengv2/sensor/PinholeCamera.hpp>

#include "../include/ecb/calib_init.hpp"    // the product's OpenCV-free calibration, used as the calibrateCamera / PnP hooks
#include "../include/ecb/circles_grid.hpp"  // the product's grid finder, used as the findCirclesGrid hook

namespace hook {
std::vector<cv::Point2f> grid_points;      // candidate centres of the last findCirclesGrid call (cv::Point2f like the reference's)
int grid_calls = 0;
const double *image_points = nullptr;      // [features][5][2] for projectPoints (null: real projection)
int project_calls = 0;
int calibrate_calls = 0, calibrate_views = 0, calibrate_flags = 0;
}  // namespace hook

namespace cv {
bool findCirclesGrid(const std::vector<Point2f> &points_, Size patternSize, std::vector<Point2f> &centers, int flags) {
    ++hook::grid_calls;
    if (flags & CALIB_CB_CLUSTERING) return false;  // the reference's second attempt (CirclesEventFrame.cpp:335-336)
    hook::grid_points = points_;
    std::vector<ecb::Pt2> pts;
    for (const auto &p : points_) pts.push_back(ecb::Pt2{(double) p.x, (double) p.y});
    std::vector<int> order;
    if (!ecb::find_asymmetric_circles_grid(pts, patternSize.height, patternSize.width, order)) return false;
    centers.clear();
    for (int idx : order) centers.push_back(points_[(size_t) idx]);
    return true;
}
// projectPoints: with hook::image_points set, the caller's points (rectifyFeatures tests); otherwise the real projection through
// the product's host header (cvCalibration flow)
void projectPoints(const std::vector<Point3f> &objectPoints, const Mat &rvec, const Mat &tvec, const Mat &cameraMatrix,
                   const Mat &distCoeffs, std::vector<Point2f> &imagePoints) {
    imagePoints.clear();
    if (hook::image_points) {
        for (size_t i = 0; i < objectPoints.size(); ++i) {
            const double *p = hook::image_points + ((size_t) hook::project_calls * objectPoints.size() + i) * 2;
            imagePoints.emplace_back(p[0], p[1]);
        }
        ++hook::project_calls;
        return;
    }
    ecb::CameraModel cam;
    cam.fx = cameraMatrix.d[0], cam.fy = cameraMatrix.d[4], cam.cx = cameraMatrix.d[2], cam.cy = cameraMatrix.d[5];
    for (size_t k = 0; k < 5 && k < distCoeffs.d.size(); ++k) cam.dist[k] = distCoeffs.d[k];
    for (const auto &o : objectPoints) {
        const double X[3] = {o.x, o.y, o.z};
        double uv[2];
        ecb::projectPoints(X, 1, rvec.d.data(), tvec.d.data(), cam, uv);
        imagePoints.emplace_back(uv[0], uv[1]);
    }
}
void Rodrigues(const Mat &src, Mat &dst) {
    if (src.d.size() == 9) {
        dst = Mat::zeros(3, 1, 0);
        ecb::rodriguesInverse(src.d.data(), dst.d.data());
    } else {
        dst = Mat::zeros(3, 3, 0);
        ecb::rodrigues<double>(src.d.data(), dst.d.data());
    }
}
double calibrateCamera(const std::vector<std::vector<Point3f>> &objectPoints, const std::vector<std::vector<Point2f>> &imagePoints,
                       Size imageSize, Mat &cameraMatrix, Mat &distCoeffs, std::vector<Mat> &rvecs, std::vector<Mat> &tvecs, int flags) {
    ++hook::calibrate_calls;
    hook::calibrate_views = (int) imagePoints.size();
    hook::calibrate_flags = flags;
    std::vector<double> obj;
    for (const auto &o : objectPoints[0]) obj.insert(obj.end(), {(double) o.x, (double) o.y, (double) o.z});
    std::vector<std::vector<double>> img;
    for (const auto &v : imagePoints) {
        img.emplace_back();
        for (const auto &q : v) img.back().insert(img.back().end(), {(double) q.x, (double) q.y});
    }
    ecb::CalibFlags fl;
    fl.fixPrincipalPoint = flags & CALIB_FIX_PRINCIPAL_POINT;
    fl.zeroTangentDist = flags & CALIB_ZERO_TANGENT_DIST;
    fl.fixAspectRatio = flags & CALIB_FIX_ASPECT_RATIO;
    fl.fixK1 = flags & CALIB_FIX_K1, fl.fixK2 = flags & CALIB_FIX_K2, fl.fixK3 = flags & CALIB_FIX_K3;
    fl.aspectRatio = cameraMatrix.d[0] / cameraMatrix.d[4];
    ecb::CameraModel cam;
    std::vector<std::array<double, 3>> rv, tv;
    const double rms = ecb::calibrateCamera(obj, img, imageSize.width, imageSize.height, fl, cam, rv, tv);
    cameraMatrix = Mat::eye(3, 3, 0);
    cameraMatrix.d[0] = cam.fx, cameraMatrix.d[4] = cam.fy, cameraMatrix.d[2] = cam.cx, cameraMatrix.d[5] = cam.cy;
    distCoeffs = Mat::zeros(5, 1, 0);  // OpenCV trims the 8 coefficients to 5 without CALIB_RATIONAL_MODEL
    for (int k = 0; k < 5; ++k) distCoeffs.d[(size_t) k] = cam.dist[k];
    rvecs.clear(), tvecs.clear();
    for (size_t v = 0; v < rv.size(); ++v) {
        Mat r = Mat::zeros(3, 1, 0), t = Mat::zeros(3, 1, 0);
        for (int k = 0; k < 3; ++k) r.d[(size_t) k] = rv[v][(size_t) k], t.d[(size_t) k] = tv[v][(size_t) k];
        rvecs.push_back(r), tvecs.push_back(t);
    }
    return rms;
}
bool solvePnPRansac(const std::vector<Point3f> &objectPoints, const std::vector<Point2f> &imagePoints, const Mat &cameraMatrix,
                    const Mat &distCoeffs, Mat &rvec, Mat &tvec, bool, int, float reprojectionError, double, std::vector<int> &inliers,
                    int) {
    std::vector<double> obj, img;
    for (const auto &o : objectPoints) obj.insert(obj.end(), {(double) o.x, (double) o.y, (double) o.z});
    for (const auto &q : imagePoints) img.insert(img.end(), {(double) q.x, (double) q.y});
    ecb::CameraModel cam;
    cam.fx = cameraMatrix.d[0], cam.fy = cameraMatrix.d[4], cam.cx = cameraMatrix.d[2], cam.cy = cameraMatrix.d[5];
    for (size_t k = 0; k < 5 && k < distCoeffs.d.size(); ++k) cam.dist[k] = distCoeffs.d[k];
    rvec = Mat::zeros(3, 1, 0), tvec = Mat::zeros(3, 1, 0);
    return ecb::solvePnPPlanar(obj, img, cam, (double) reprojectionError, rvec.d.data(), tvec.d.data(), inliers);
}
}  // namespace cv

namespace {
struct CircleProbe : opengv2::CirclesEventFrame {
    CircleProbe(opengv2::EventContainer::Ptr c, const std::pair<double, double> &d, CirclePatternParameters::Ptr pattern, Params p)
        : opengv2::CirclesEventFrame(c, d, pattern, p) {}
    opengv2::vectorofEigenMatrix<Eigen::Vector2d> &pos() { return positiveEvents_; }
    opengv2::vectorofEigenMatrix<Eigen::Vector2d> &neg() { return negativeEvents_; }
    std::vector<opengv2::FeatureBase::Ptr> &feats() { return features_; }
    void fit(const std::vector<uint> &ps, const std::vector<uint> &ns, Eigen::Vector2d &c, double &r) { fitCircle(ps, ns, c, r); }
    double rthr() const { return circleRadiusThreshold_; }
};

std::shared_ptr<CircleProbe> make_frame(const double *t, const double *x, const double *y, const unsigned char *pol, long long n,
                                        double t0, double t1, int W, int H, const double *prm) {
    auto container = std::make_shared<opengv2::EventContainer>();
    container->camera = std::make_shared<opengv2::PinholeCamera>(Eigen::Vector2d(W, H));
    for (long long i = 0; i < n; ++i)
        container->container.emplace(t[i], opengv2::Event_loc_pol(Eigen::Vector2d(x[i], y[i]), pol[i] != 0));
    cv::FileStorage fs;  // the reference's own parameter constructors read these keys (parameters.hpp:15-21, CirclesEventFrame.cpp:42-48)
    fs.kv = {{"BoardSize_Cols", prm[0]}, {"BoardSize_Rows", prm[1]}, {"Square_Size", prm[2]}, {"Is_Pattern_Asymmetric", prm[3]},
             {"Circles_Radius", prm[4]}, {"dbscan_eps", prm[5]}, {"dbscan_startMinSample", prm[6]}, {"clusterMinSample", prm[7]},
             {"knn_num", prm[8]}, {"fitCircle", prm[9]}};
    auto pattern = std::make_shared<CirclePatternParameters>(fs);
    opengv2::CirclesEventFrame::Params params(fs);
    return std::make_shared<CircleProbe>(container, std::make_pair(t0, t1), pattern, params);
}
}  // namespace

extern "C" {
// prm = cols rows square asymmetric radius | eps startMinSample clusterMinSample knn fitCircle.
// Runs the reference's constructor + extractFeatures().  cand_f32[cap][2]: the candidate centres it handed to findCirclesGrid
// (cv::Point2f); features[rows*cols][3]: centre and radius of features_ in board order when the grid was found.
// Returns 1 found / 0 not found; *n_cand = -1 when findCirclesGrid was never reached (too few clusters, :127-129).
int ref_extract(const double *t, const double *x, const double *y, const unsigned char *pol, long long n, double t0, double t1, int W,
                int H, const double *prm, float *cand_f32, int cap, int *n_cand, double *features, double *rthr) {
    auto f = make_frame(t, x, y, pol, n, t0, t1, W, H, prm);
    hook::grid_points.clear();
    hook::grid_calls = 0;
    const bool ok = f->extractFeatures();
    *rthr = f->rthr();
    *n_cand = hook::grid_calls ? (int) hook::grid_points.size() : -1;
    for (size_t i = 0; i < hook::grid_points.size() && (int) i < cap; ++i)
        cand_f32[2 * i] = hook::grid_points[i].x, cand_f32[2 * i + 1] = hook::grid_points[i].y;
    if (ok)
        for (size_t i = 0; i < f->feats().size(); ++i) {
            auto c = dynamic_cast<opengv2::CalibCircle *>(f->feats()[i].get());
            features[3 * i] = c->location()[0], features[3 * i + 1] = c->location()[1], features[3 * i + 2] = c->radius;
        }
    return ok ? 1 : 0;
}

// the reference's fitCircle over explicit point sets (+ then -)
void ref_fit_circle(const double *pxy, int np, const double *nxy, int nn, double *out3) {
    const double t = 0, xx = 0, yy = 0;
    const unsigned char pp = 1;
    const double prm[10] = {4, 9, 5.5, 1, 1.75, 4, 2, 5, 3, 1};
    auto f = make_frame(&t, &xx, &yy, &pp, 0, 0, 1, 346, 260, prm);
    std::vector<uint> ps, ns;
    for (int i = 0; i < np; ++i) f->pos().emplace_back(pxy[2 * i], pxy[2 * i + 1]), ps.push_back((uint) i);
    for (int i = 0; i < nn; ++i) f->neg().emplace_back(nxy[2 * i], nxy[2 * i + 1]), ns.push_back((uint) i);
    Eigen::Vector2d c;
    double r;
    f->fit(ps, ns, c, r);
    out3[0] = c[0], out3[1] = c[1], out3[2] = r;
}

// extractFeatures() then rectifyFeatures() with the caller's projections image_points[rows*cols][5][2] (centre + 4 quadrant
// points per board circle, :431-456).  out[rows*cols][3]: rectified centre / radius per board circle, radius -1 = deleted.
// Returns -1 when extractFeatures() fails, else rectifyFeatures()'s verdict (0 / 1); find_xy[n_find][2] -> find_id[n_find]: the
// landmark findCenter() returns for a pixel afterwards (-1: none), only when the verdict is 1.
int ref_rectify(const double *t, const double *x, const double *y, const unsigned char *pol, long long n, double t0, double t1, int W,
                int H, const double *prm, const double *image_points, double *out, const double *find_xy, int n_find, int *find_id) {
    auto f = make_frame(t, x, y, pol, n, t0, t1, W, H, prm);
    hook::grid_calls = 0;
    if (!f->extractFeatures()) return -1;
    const int nc = (int) f->feats().size();
    std::vector<opengv2::LandmarkBase::Ptr> lms;  // the board points (EventCalibIni.cpp:99-113), kept alive here
    const int cols = (int) prm[0];
    for (int i = 0; i < nc; ++i) {
        const int r = i / cols, c = i % cols;
        lms.push_back(std::make_shared<opengv2::LandmarkBase>(i, Eigen::Vector3d((prm[3] != 0 ? (2 * c + r % 2) : c) * prm[2], r * prm[2], 0)));
        f->feats()[(size_t) i]->setLandmark(lms.back());
    }
    for (int i = 0; i < nc; ++i) out[3 * i] = out[3 * i + 1] = 0.0, out[3 * i + 2] = -1.0;
    hook::image_points = image_points;
    hook::project_calls = 0;
    Eigen::Matrix3d R;
    const bool ok = f->rectifyFeatures(std::unordered_set<int>(), R, Eigen::Vector3d(0, 0, 0));
    hook::image_points = nullptr;
    for (auto &fb : f->feats()) {
        auto c = dynamic_cast<opengv2::CalibCircle *>(fb.get());
        const int id = c->landmark()->id();
        out[3 * id] = c->location()[0], out[3 * id + 1] = c->location()[1], out[3 * id + 2] = c->radius;
    }
    if (ok)
        for (int i = 0; i < n_find; ++i) {
            auto lm = f->findCenter(Eigen::Vector2d(find_xy[2 * i], find_xy[2 * i + 1]));
            find_id[i] = lm ? lm->id() : -1;
        }
    return ok ? 1 : 0;
}
}

// ---- EventCalibSpline (a7 association, a12 problem assembly, spline set-up) ----
#include <ceres/ceres.h>
#include <opengv2/map/MapBase.hpp>

namespace hook {
struct SolveRecord {
    int calls = 0;
    double gradient_tolerance = 0, function_tolerance = 0, huber = 0;
    int linear_solver = -1, n_param_blocks = 0, n_quaternion_blocks = 0;
    std::vector<std::vector<double *>> residual_params;
} solve;
}  // namespace hook

namespace ceres {
void Solve(const Solver::Options &options, Problem *problem, Solver::Summary *) {
    hook::SolveRecord &r = hook::solve;
    ++r.calls;
    r.gradient_tolerance = options.gradient_tolerance;
    r.function_tolerance = options.function_tolerance;
    r.linear_solver = (int) options.linear_solver_type;
    r.n_param_blocks = (int) problem->parameters.size();
    r.n_quaternion_blocks = 0;
    for (const auto &pb : problem->parameters)
        if (pb.size == 4 && dynamic_cast<const EigenQuaternionParameterization *>(pb.local)) ++r.n_quaternion_blocks;
    r.residual_params.clear();
    r.huber = 0;
    for (const auto &rb : problem->residuals) {
        r.residual_params.push_back(rb.params);
        if (auto h = dynamic_cast<const HuberLoss *>(rb.loss)) r.huber = h->a_;
    }
}
}  // namespace ceres

namespace {
// a key frame in the state rectifyFeatures() leaves it in (CirclesEventFrame.cpp:628-637): features_ with landmarks, circles_,
// circleKdTree_
struct KeyFrameProbe : opengv2::CirclesEventFrame {
    KeyFrameProbe(opengv2::EventContainer::Ptr c, const std::pair<double, double> &d, CirclePatternParameters::Ptr pattern)
        : opengv2::CirclesEventFrame(c, d, pattern) {}
    void setCircles(const double *circ, int n, const std::vector<opengv2::LandmarkBase::Ptr> &lms) {
        for (int i = 0; i < n; ++i) {
            if (circ[3 * i + 2] < 0) continue;  // feature deleted by rectifyFeatures
            auto f = std::make_shared<opengv2::CalibCircle>(Eigen::Vector2d(circ[3 * i], circ[3 * i + 1]), circ[3 * i + 2]);
            f->setLandmark(lms[(size_t) i]);
            features_.push_back(f);
        }
        circles_.reserve(features_.size());
        for (const auto &f : features_) circles_.push_back(f->location());
        circleKdTree_ = std::make_shared<KDTreeVectorOfVectorsAdaptor<opengv2::vectorofEigenMatrix<Eigen::Vector2d>, double, 2>>(2, circles_, 10);
    }
};
struct SplineProbe : opengv2::EventCalibSpline {
    using opengv2::EventCalibSpline::EventCalibSpline;
    std::vector<opengv2::BsplineReal<3>> &tw() { return twbSplines_; }
    std::vector<opengv2::BsplineReal<4>> &qw() { return QwbSplines_; }
    const std::vector<RelationContainer> &relations() const { return relationContainer_; }
    const Eigen::Matrix<double, 9, 1> &intr() const { return intrinsics_; }
    const std::vector<std::pair<double, double>> &ranges() const { return time2splineIdx_; }
};
}  // namespace

extern "C" {
// Runs the reference's EventCalibSpline constructor (set-up + optimize() with the no-op Solve + updateMap()).
//  events (t, x, y, pol)[n]; key frames: stamp[K], pose q[K][4] (x y z w) / twb[K][3], circles[K][n_circ][3] (r < 0: absent);
//  board[n_circ][3]; cam = fx fy cx cy k1 k2 p1 p2 k3; W, H; step = MotionTimeStep; radius = Circles_Radius.
// Outputs (caller-sized): info[8] = n_splines, n_residuals, solve calls, parameter blocks, quaternion blocks, linear solver,
//  key frames left in the map, 0; ncp[n_splines]; knots / rot_cp / trans_cp concatenated over the splines; ranges[n_splines][2];
//  intr9; huber_tol[3] = Huber a, gradient tolerance, function tolerance; per residual: obs[2], lm[3], basis[4], span, spline,
//  and the FIRST control-point index of its 4 rotation / 4 translation blocks (res_cp[2]) as handed to AddResidualBlock;
//  kf_pose[K][8] = stamp, twb, q (x y z w) of the key frames after updateMap() (NaN rows for frames removed from the map).
// Returns 0, or -1 with what() in err (the constructor throws std::logic_error like the reference's).
int ref_calib_spline(const double *t, const double *x, const double *y, const unsigned char *pol, long long n, const double *stamp,
                     const double *kq, const double *kt, const double *circles, int K, int n_circ, const double *board,
                     const double *cam, int W, int H, double step, double radius, int *info, int *ncp, double *knots,
                     double *rot_cp, double *trans_cp, double *ranges, double *intr9, double *huber_tol, long long res_cap,
                     double *res_obs, double *res_lm, double *res_basis, int *res_span, int *res_spline, int *res_cp,
                     double *kf_pose, char *err, int err_cap) {
    auto container = std::make_shared<opengv2::EventContainer>();
    auto camera = std::make_shared<opengv2::PinholeCamera>(Eigen::Vector2d(W, H));
    Eigen::Matrix3d Km;
    Km << cam[0], 0, cam[2], 0, cam[1], cam[3], 0, 0, 1;
    camera->setK(Km);
    for (int i = 0; i < 5; ++i) camera->distCoeffs()[i] = cam[4 + i];
    container->camera = camera;
    for (long long i = 0; i < n; ++i)
        container->container.emplace(t[i], opengv2::Event_loc_pol(Eigen::Vector2d(x[i], y[i]), pol[i] != 0));
    cv::FileStorage fs;
    fs.kv = {{"BoardSize_Cols", 4}, {"BoardSize_Rows", 9}, {"Square_Size", 5.5}, {"Is_Pattern_Asymmetric", 1}, {"Circles_Radius", radius}};
    auto pattern = std::make_shared<CirclePatternParameters>(fs);
    std::vector<opengv2::LandmarkBase::Ptr> lms;
    for (int i = 0; i < n_circ; ++i)
        lms.push_back(std::make_shared<opengv2::LandmarkBase>(i, Eigen::Vector3d(board[3 * i], board[3 * i + 1], board[3 * i + 2])));
    auto map = std::make_shared<opengv2::MapBase>();
    for (int k = 0; k < K; ++k) {
        // the frame's own window is irrelevant here (its event sets are released after rectifyFeatures): an empty one
        auto frame = std::make_shared<KeyFrameProbe>(container, std::make_pair(-2.0, -1.0), pattern);
        frame->setCircles(circles + (size_t) k * n_circ * 3, n_circ, lms);
        map->addFrame(std::make_shared<opengv2::Bodyframe>(frame, stamp[k], Eigen::Vector3d(kt[3 * k], kt[3 * k + 1], kt[3 * k + 2]),
                                                            Eigen::Quaterniond(kq[4 * k + 3], kq[4 * k], kq[4 * k + 1], kq[4 * k + 2])));
    }
    hook::solve = hook::SolveRecord();
    std::unique_ptr<SplineProbe> sp;
    try {
        sp.reset(new SplineProbe(map, container, false, false, step, radius));
    } catch (const std::exception &e) {
        std::snprintf(err, (size_t) err_cap, "%s", e.what());
        return -1;
    }
    const int S = (int) sp->tw().size();
    info[0] = S;
    info[1] = (int) sp->relations().size();
    info[2] = hook::solve.calls;
    info[3] = hook::solve.n_param_blocks;
    info[4] = hook::solve.n_quaternion_blocks;
    info[5] = hook::solve.linear_solver;
    info[6] = (int) map->frameNum();
    info[7] = 0;
    size_t ko = 0, ro = 0, to = 0;
    for (int s = 0; s < S; ++s) {
        const auto &kv = sp->tw()[(size_t) s].getKnotVector();
        auto &tc = sp->tw()[(size_t) s].getCP();
        auto &qc = sp->qw()[(size_t) s].getCP();
        ncp[s] = (int) tc.size();
        for (double k : kv) knots[ko++] = k;
        for (auto &c : qc)
            for (int a = 0; a < 4; ++a) rot_cp[ro++] = c[a];
        for (auto &c : tc)
            for (int a = 0; a < 3; ++a) trans_cp[to++] = c[a];
        ranges[2 * s] = sp->ranges()[(size_t) s].first;
        ranges[2 * s + 1] = sp->ranges()[(size_t) s].second;
    }
    for (int i = 0; i < 9; ++i) intr9[i] = sp->intr()[i];
    huber_tol[0] = hook::solve.huber;
    huber_tol[1] = hook::solve.gradient_tolerance;
    huber_tol[2] = hook::solve.function_tolerance;
    long long r = 0;
    for (const auto &rel : sp->relations()) {
        if (r >= res_cap) break;
        res_obs[2 * r] = rel.obs[0], res_obs[2 * r + 1] = rel.obs[1];
        for (int a = 0; a < 3; ++a) res_lm[3 * r + a] = rel.lm[a];
        for (int a = 0; a < 4; ++a) res_basis[4 * r + a] = (*rel.rBasis)[0][(size_t) a];
        res_span[r] = (int) rel.rSpanIdx;
        res_spline[r] = (int) rel.splineIdx;
        const auto &ps = hook::solve.residual_params[(size_t) r];  // intrinsics, 4 rotation blocks, 4 translation blocks
        res_cp[2 * r] = (int) ((ps[1] - sp->qw()[rel.splineIdx].getCP()[0].data()) / 4);
        res_cp[2 * r + 1] = (int) ((ps[5] - sp->tw()[rel.splineIdx].getCP()[0].data()) / 3);
        ++r;
    }
    for (int k = 0; k < K; ++k) {
        auto bf = map->keyframe(stamp[k]);
        double *o = kf_pose + 8 * k;
        if (!bf) {
            for (int a = 0; a < 8; ++a) o[a] = std::nan("");
            continue;
        }
        o[0] = bf->timeStamp();
        for (int a = 0; a < 3; ++a) o[1 + a] = bf->twb()[a];
        for (int a = 0; a < 4; ++a) o[4 + a] = bf->unitQwb().coeffs()[a];
    }
    return 0;
}
}

// ---- EventCalibIni: tracking gate (f-1), checkPose and cvCalibration flow (f-4) ----
#include <opengv2/event_camera_calib/EventCalibIni.hpp>
#include <opengv2/system/SystemBase.hpp>

namespace {
struct IniSession {
    std::shared_ptr<opengv2::EventContainer> container;
    std::shared_ptr<opengv2::PinholeCamera> camera;
    std::shared_ptr<opengv2::MapBase> map;
    opengv2::SystemBase system;
    CalibrationSetting::Ptr setting;
    std::shared_ptr<opengv2::EventCalibIni> ini;
    double prm[10];
    int W, H;
};
IniSession *ini_new(int W, int H, double step, const double *prm10, int n_use) {
    auto *s = new IniSession();
    s->W = W, s->H = H;
    for (int i = 0; i < 10; ++i) s->prm[i] = prm10[i];
    s->container = std::make_shared<opengv2::EventContainer>();
    s->camera = std::make_shared<opengv2::PinholeCamera>(Eigen::Vector2d(W, H));
    s->container->camera = s->camera;
    s->map = std::make_shared<opengv2::MapBase>();
    s->system.map = s->map;
    s->system.viewer = std::make_shared<opengv2::ViewerBase>();  // cvCalibration() calls it unconditionally (EventCalibIni.cpp:320)
    cv::FileStorage fs;  // parameters.hpp:32-46 with the values of example.yaml
    fs.kv = {{"BoardSize_Cols", prm10[0]}, {"BoardSize_Rows", prm10[1]}, {"Square_Size", prm10[2]}, {"Is_Pattern_Asymmetric", prm10[3]},
             {"Circles_Radius", prm10[4]}, {"Calibrate_FixAspectRatio", 1}, {"Calibrate_AssumeZeroTangentialDistortion", 1},
             {"Calibrate_FixPrincipalPointAtTheCenter", 1}, {"Calibrate_UseFisheyeModel", 0}, {"Fix_K1", 0}, {"Fix_K2", 0}, {"Fix_K3", 0},
             {"Fix_K4", 1}, {"Fix_K5", 1}, {"Calibrate_NrOfFrameToUse", (double) n_use}};
    s->setting = std::make_shared<CalibrationSetting>(fs);
    s->ini = std::make_shared<opengv2::EventCalibIni>(s->map, s->setting, step);
    s->ini->setSystem(&s->system);
    return s;
}
}  // namespace

extern "C" {
void *ref_ini_new(int W, int H, double step, const double *prm10, int n_use) { return ini_new(W, H, step, prm10, n_use); }
void ref_ini_free(void *h) { delete (IniSession *) h; }
void ref_ini_add_events(void *h, const double *t, const double *x, const double *y, const unsigned char *pol, long long n) {
    auto *s = (IniSession *) h;
    for (long long i = 0; i < n; ++i)
        s->container->container.emplace(t[i], opengv2::Event_loc_pol(Eigen::Vector2d(x[i], y[i]), pol[i] != 0));
}
// tracking->process(bf) (eventCameraCalib.cpp:60) for a frame whose features are given (board order, centres only): the first
// frame initialises the map, later ones go through EventCalibIni::track.  Returns 1 accepted / 0 rejected.
int ref_ini_gate(void *h, double stamp, const double *xy, int n_feat) {
    auto *s = (IniSession *) h;
    cv::FileStorage fs;
    fs.kv = {{"BoardSize_Cols", s->prm[0]}, {"BoardSize_Rows", s->prm[1]}, {"Square_Size", s->prm[2]}, {"Is_Pattern_Asymmetric", s->prm[3]},
             {"Circles_Radius", s->prm[4]}};
    auto frame = std::make_shared<KeyFrameProbe>(s->container, std::make_pair(-2.0, -1.0), std::make_shared<CirclePatternParameters>(fs));
    std::vector<double> circ((size_t) n_feat * 3);
    for (int i = 0; i < n_feat; ++i) circ[3 * i] = xy[2 * i], circ[3 * i + 1] = xy[2 * i + 1], circ[3 * i + 2] = 1.0;
    std::vector<opengv2::LandmarkBase::Ptr> none((size_t) n_feat);
    frame->setCircles(circ.data(), n_feat, none);
    auto bf = std::make_shared<opengv2::Bodyframe>(frame, stamp, Eigen::Vector3d(0, 0, 0), Eigen::Quaterniond(1, 0, 0, 0));
    return s->ini->process(bf) ? 1 : 0;
}
// EventCalibIni::checkPose(cur) against a map whose last key frame is `ref` (poses: q = x y z w of Qwb, t = twb)
int ref_check_pose(double ref_stamp, const double *rq, const double *rt, double cur_stamp, const double *cq, const double *ct, double step) {
    const double prm[10] = {4, 9, 5.5, 1, 1.75, 4, 2, 5, 3, 0};
    std::unique_ptr<IniSession> s(ini_new(346, 260, step, prm, 200));
    s->map->addFrame(std::make_shared<opengv2::Bodyframe>(nullptr, ref_stamp, Eigen::Vector3d(rt[0], rt[1], rt[2]),
                                                           Eigen::Quaterniond(rq[3], rq[0], rq[1], rq[2])));
    auto cur = std::make_shared<opengv2::Bodyframe>(nullptr, cur_stamp, Eigen::Vector3d(ct[0], ct[1], ct[2]),
                                                    Eigen::Quaterniond(cq[3], cq[0], cq[1], cq[2]));
    return s->ini->checkPose(cur) ? 1 : 0;
}
// The reference's front-to-back flow on raw events for given windows: per window CirclesEventFrame + extractFeatures(), the
// tracking gate, then EventCalibIni::cvCalibration().  Outputs: cam9 (fx fy cx cy k1 k2 p1 p2 k3), per input window
// status[w] = 0 no features / 1 gate rejected / 2 in the map before cvCalibration but dropped by it / 3 kept, pose[w][7] = twb,
// Qwb (x y z w) and feat[w][n_feat][3] = rectified circles (r < 0: deleted) for kept frames.  Returns cvCalibration()'s bool.
int ref_ini_run(void *h, const double *windows, int n_win, int fit_circle, double *cam9, int *status, double *pose, double *feat,
                int *counts) {
    auto *s = (IniSession *) h;
    double prm[10];
    for (int i = 0; i < 10; ++i) prm[i] = s->prm[i];
    prm[9] = fit_circle;
    cv::FileStorage fs;
    fs.kv = {{"BoardSize_Cols", prm[0]}, {"BoardSize_Rows", prm[1]}, {"Square_Size", prm[2]}, {"Is_Pattern_Asymmetric", prm[3]},
             {"Circles_Radius", prm[4]}, {"dbscan_eps", prm[5]}, {"dbscan_startMinSample", prm[6]}, {"clusterMinSample", prm[7]},
             {"knn_num", prm[8]}, {"fitCircle", prm[9]}};
    auto pattern = std::make_shared<CirclePatternParameters>(fs);
    opengv2::CirclesEventFrame::Params params(fs);
    const int n_feat = (int) (prm[0] * prm[1]);
    std::map<double, int> stamp2win;
    hook::image_points = nullptr;
    for (int w = 0; w < n_win; ++w) {
        status[w] = 0;
        auto frame = std::make_shared<CircleProbe>(s->container, std::make_pair(windows[2 * w], windows[2 * w + 1]), pattern, params);
        if (!frame->extractFeatures()) continue;
        const double ts = (windows[2 * w] + windows[2 * w + 1]) / 2;  // Bodyframe time stamp (eventCameraCalib.cpp:57)
        auto bf = std::make_shared<opengv2::Bodyframe>(frame, ts, Eigen::Vector3d(0, 0, 0), Eigen::Quaterniond(1, 0, 0, 0));
        status[w] = s->ini->process(bf) ? 2 : 1;
        if (status[w] == 2) stamp2win[ts] = w;
    }
    counts[0] = (int) s->map->frameNum();
    hook::calibrate_calls = 0;
    const bool ok = s->ini->cvCalibration();
    counts[1] = (int) s->map->frameNum();
    counts[2] = hook::calibrate_views;
    counts[3] = hook::calibrate_flags;
    const Eigen::Matrix3d &K = s->camera->K();
    cam9[0] = K(0, 0), cam9[1] = K(1, 1), cam9[2] = K(0, 2), cam9[3] = K(1, 2);
    for (int k = 0; k < 5; ++k) cam9[4 + k] = s->camera->distCoeffs().size() > k ? s->camera->distCoeffs()[k] : 0.0;
    for (const auto &kv : s->map->keyframes()) {
        const int w = stamp2win[kv.first];
        status[w] = 3;
        const auto &bf = kv.second;
        double *o = pose + 7 * w;
        for (int a = 0; a < 3; ++a) o[a] = bf->twb()[a];
        for (int a = 0; a < 4; ++a) o[3 + a] = bf->unitQwb().coeffs()[a];
        double *f = feat + (size_t) w * n_feat * 3;
        for (int i = 0; i < n_feat; ++i) f[3 * i] = f[3 * i + 1] = 0.0, f[3 * i + 2] = -1.0;
        for (auto &fb : bf->frame(0)->features()) {
            auto c = dynamic_cast<opengv2::CalibCircle *>(fb.get());
            const int id = c->landmark()->id();
            f[3 * id] = c->location()[0], f[3 * id + 1] = c->location()[1], f[3 * id + 2] = c->radius;
        }
    }
    return ok ? 1 : 0;
}
}
