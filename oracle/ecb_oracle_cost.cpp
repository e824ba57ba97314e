// TEST INFRASTRUCTURE — CPU restatement of EventCalib's cost evaluation (the parity oracle of the CUDA
// residual / Jacobian / normal-equation kernels).  Not product code; see oracle/ecb_oracle_frontend.cpp.
//
// What is restated (paths relative to /root/reference/modules/):
//   findSpan / dersBasisFuns          core/spline/include/opengv2/spline/BsplineReal.hpp:208-231,107-145
//   knot placement (eq. 9.68)         core/spline/include/opengv2/spline/BsplineReal.hpp:87-100
//   unDistort                         camera_calibration/event_camera_calib/include/opengv2/event_camera_calib/EventCalibSpline.hpp:36-63
//   CalibReprojectionError::operator() .../EventCalibSpline.hpp:168-229, evaluated on forward-mode dual numbers with 37
//                                     partials — the arithmetic ceres::Jet<double,37> performs under AutoDiffCostFunction
//   findCenter + association loop     .../CirclesEventFrame.hpp:50-65, camera_calibration/event_camera_calib/src/EventCalibSpline.cpp:157-192
//   inverseRadialDistortion           core/sensor/src/PinholeCamera.cpp:70-95
// [external — Ceres 1.x / Eigen, NOT in /root/reference, restated from their published behaviour]:
//   HuberLoss + Corrector (rho'' <= 0 => residual and Jacobian scaled by sqrt(rho')), cost = 1/2 rho(r^2)
//   EigenQuaternionParameterization::ComputeJacobian / Plus (storage x,y,z,w)
//   Eigen: Vector4::normalize(), Quaternion * Vector3 (uv = 2 q.vec x v; v + w uv + q.vec x uv), norm(), dot()
// Pinning: (1) the reference's only known-answer on this path, unit_test_inverseDistortion, checked in
// tests/test_oracle_cost.py; (2) the reference's OWN functor and B-spline compiled where they lie against stand-in Eigen
// headers (oracle/ref_functor_capi.cpp -> oracle/_ref/libref_functor.so): the restated residual() below equals it bit for bit
// on Jet<37> (value + 37 partials), knots / findSpan / dersBasisFuns are bit-identical (tests/test_oracle_reference_source.py);
// (3) the association (orc_associate) returns the same ordered (event, landmark) list, spans and basis values as the
// reference's own EventCalibSpline.cpp constructor compiled in place with a recording ceres::Problem; (4) an independent
// 40-digit mpmath evaluation (tests/test_oracle_cost.py).  Still PARITY UNPINNED: what Ceres and Sophus do
// (corrector, parameterisations, SO(3) exp / log) — external, restated from their published behaviour.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

// ---------------------------------------------------------------- B-spline basis (degree 3) ----
size_t find_span(const double *knots, size_t nk, double u) {  // BsplineReal.hpp:208-231
    const size_t degree = 3;
    size_t n = nk - 2 - degree;
    if (u == knots[n + 1]) return n;
    size_t low = degree, high = n + 1, mid = (low + high) / 2;
    while (u < knots[mid] || u >= knots[mid + 1]) {
        if (u < knots[mid])
            high = mid;
        else
            low = mid;
        mid = (low + high) / 2;
    }
    return mid;
}

void basis_funs(const double *knots, size_t span, double u, double N[4]) {  // BsplineReal.hpp:107-145 (derivativeLimit 0)
    const int degree = 3;
    double ndu[4][4], left[4], right[4];
    ndu[0][0] = 1;
    for (int j = 1; j <= degree; j++) {
        left[j] = u - knots[span + 1 - j];
        right[j] = knots[span + j] - u;
        double saved = 0.0;
        for (int r = 0; r < j; ++r) {
            ndu[j][r] = right[r + 1] + left[j - r];
            double temp = ndu[r][j - 1] / ndu[j][r];
            ndu[r][j] = saved + right[r + 1] * temp;
            saved = left[j - r] * temp;
        }
        ndu[j][j] = saved;
    }
    for (int j = 0; j <= degree; j++) N[j] = ndu[j][degree];
}

// ---------------------------------------------------------------- dual numbers ----
template <int NP>
struct Jet {
    double a;
    double v[NP];
    Jet() : a(0) { std::memset(v, 0, sizeof v); }
    Jet(double s) : a(s) { std::memset(v, 0, sizeof v); }  // NOLINT
};
#define JET_BIN(op, expr_a, expr_v)                                      \
    template <int NP> Jet<NP> operator op(const Jet<NP> &f, const Jet<NP> &g) { \
        Jet<NP> h;                                                        \
        h.a = expr_a;                                                     \
        for (int i = 0; i < NP; ++i) h.v[i] = expr_v;                     \
        return h;                                                         \
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
template <int NP> Jet<NP> operator-(const Jet<NP> &f) {
    Jet<NP> h;
    h.a = -f.a;
    for (int i = 0; i < NP; ++i) h.v[i] = -f.v[i];
    return h;
}
template <int NP> Jet<NP> operator-(double s, const Jet<NP> &f) { return Jet<NP>(s) - f; }
template <int NP> Jet<NP> operator-(const Jet<NP> &f, double s) { return f - Jet<NP>(s); }
template <int NP> Jet<NP> operator+(double s, const Jet<NP> &f) { return Jet<NP>(s) + f; }
template <int NP> Jet<NP> jsqrt(const Jet<NP> &f) {
    Jet<NP> h;
    h.a = std::sqrt(f.a);
    const double t = 1.0 / (2.0 * h.a);
    for (int i = 0; i < NP; ++i) h.v[i] = f.v[i] * t;
    return h;
}
inline double jsqrt(double f) { return std::sqrt(f); }
// ceres/jet.h: sin, cos, atan
template <int NP> Jet<NP> jsin(const Jet<NP> &f) {
    Jet<NP> h;
    h.a = std::sin(f.a);
    const double c = std::cos(f.a);
    for (int i = 0; i < NP; ++i) h.v[i] = c * f.v[i];
    return h;
}
template <int NP> Jet<NP> jcos(const Jet<NP> &f) {
    Jet<NP> h;
    h.a = std::cos(f.a);
    const double sn = -std::sin(f.a);
    for (int i = 0; i < NP; ++i) h.v[i] = sn * f.v[i];
    return h;
}
template <int NP> Jet<NP> jatan(const Jet<NP> &f) {
    Jet<NP> h;
    h.a = std::atan(f.a);
    const double t = 1.0 / (1.0 + f.a * f.a);
    for (int i = 0; i < NP; ++i) h.v[i] = t * f.v[i];
    return h;
}
inline double jsin(double f) { return std::sin(f); }
inline double jcos(double f) { return std::cos(f); }
inline double jatan(double f) { return std::atan(f); }
template <int NP> double jval(const Jet<NP> &f) { return f.a; }
inline double jval(double f) { return f; }

// ---------------------------------------------------------------- the residual ----
// EventCalibSpline.hpp:36-63
template <class T>
void undistort(const T &fx, const T &fy, const T &cx, const T &cy, const T &k1, const T &k2, const T &k3, const T &k4,
               const T &k5, const double obs[2], T Xc[3]) {
    Xc[0] = (obs[0] - cx) / fx;
    Xc[1] = (obs[1] - cy) / fy;
    Xc[2] = T(1.0);
    T xx = Xc[0] * Xc[0];
    T yy = Xc[1] * Xc[1];
    T r2 = xx + yy;
    T r4 = r2 * r2;
    T r6 = r4 * r2;
    T r8 = r6 * r2;
    T r10 = r8 * r2;
    T r_coeff = 1.0 + k1 * r2 + k2 * r4 + k3 * r6 + k4 * r8 + k5 * r10;
    Xc[0] = Xc[0] * r_coeff;
    Xc[1] = Xc[1] * r_coeff;
}

// EventCalibSpline.hpp:168-229.  r_cp: 4 rotation control points (x,y,z,w), t_cp: 4 translation control points.
template <class T>
T residual(const T *intr, const T *const r_cp[4], const T *const t_cp[4], const double obs[2], const double lm[3],
           double radius, const double rb[4], const double tb[4]) {
    T q[4], tw[3];
    for (int c = 0; c < 4; ++c) q[c] = rb[0] * r_cp[0][c] + rb[1] * r_cp[1][c] + rb[2] * r_cp[2][c] + rb[3] * r_cp[3][c];
    // Eigen normalize(): z = squaredNorm(); if (z > 0) v /= sqrt(z)
    T z = (q[0] * q[0] + q[1] * q[1]) + (q[2] * q[2] + q[3] * q[3]);
    T nz = jsqrt(z);
    for (int c = 0; c < 4; ++c) q[c] = q[c] / nz;
    for (int c = 0; c < 3; ++c) tw[c] = tb[0] * t_cp[0][c] + tb[1] * t_cp[1][c] + tb[2] * t_cp[2][c] + tb[3] * t_cp[3][c];
    T Xc[3];
    undistort(intr[0], intr[1], intr[2], intr[3], intr[4], intr[5], intr[6], intr[7], intr[8], obs, Xc);
    const T &qx = q[0], &qy = q[1], &qz = q[2], &qw = q[3];
    T tx = 2. * qx, ty = 2. * qy, tz = 2. * qz;
    T twx = tx * qw, twy = ty * qw, txx = tx * qx, txz = tz * qx, tyy = ty * qy, tyz = tz * qy;
    T R2[3] = {txz - twy, tyz + twx, 1.0 - (txx + tyy)};
    T depth = -tw[2] / (R2[0] * Xc[0] + (R2[1] * Xc[1] + R2[2] * Xc[2]));
    for (int c = 0; c < 3; ++c) Xc[c] = Xc[c] * depth;
    // Eigen Quaternion * Vector3: uv = q.vec x v; uv += uv; v + w*uv + q.vec x uv
    T uv[3] = {qy * Xc[2] - qz * Xc[1], qz * Xc[0] - qx * Xc[2], qx * Xc[1] - qy * Xc[0]};
    for (int c = 0; c < 3; ++c) uv[c] = uv[c] + uv[c];
    T cr[3] = {qy * uv[2] - qz * uv[1], qz * uv[0] - qx * uv[2], qx * uv[1] - qy * uv[0]};
    T d[3];
    for (int c = 0; c < 3; ++c) d[c] = ((Xc[c] + qw * uv[c]) + cr[c] + tw[c]) - lm[c];
    return jsqrt(d[0] * d[0] + (d[1] * d[1] + d[2] * d[2])) - radius;
}

// ---- Sophus::SO3 pieces used by CalibReprojectionError_SO3  [external: Sophus 1.0 so3.hpp, NOT in /root/reference] ----
// storage = Eigen quaternion coefficients x y z w.  Every SO3 built from a quaternion is normalised by Sophus' constructor
// (coeffs /= norm, Eigen's reduction order (x^2 + y^2) + (z^2 + w^2)); that includes the results of operator* and inverse().
// The same semantics are written out in the stand-in oracle/shim_functor/sophus/so3.hpp the reference's own SO(3) functor is
// compiled against (oracle/_ref/libref_functor.so) — tests/test_oracle_reference_source.py holds the two bit-identical.
template <class T> void so3_normalize(T r[4]) {
    const T n = jsqrt((r[0] * r[0] + r[1] * r[1]) + (r[2] * r[2] + r[3] * r[3]));
    for (int c = 0; c < 4; ++c) r[c] = r[c] / n;
}
template <class T> void so3_mul(const T a[4], const T b[4], T r[4]) {  // SO3::operator*
    r[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
    r[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    r[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
    r[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
    so3_normalize(r);
}
template <class T> void so3_inverse(const T a[4], T r[4]) {  // SO3(conjugate)
    r[0] = -a[0];
    r[1] = -a[1];
    r[2] = -a[2];
    r[3] = a[3];
    so3_normalize(r);
}
template <class T> void so3_log(const T q[4], T t[3]) {  // SO3::logAndTheta
    const double eps = 1e-10;  // Sophus::Constants<double>::epsilon()
    T squared_n = q[0] * q[0] + (q[1] * q[1] + q[2] * q[2]);  // vec().squaredNorm(), Eigen's reduction order
    T w = q[3];
    T two_atan_nbyw_by_n;
    if (jval(squared_n) < eps * eps) {
        T squared_w = w * w;
        two_atan_nbyw_by_n = T(2.0) / w - T(2.0 / 3.0) * (squared_n) / (w * squared_w);
    } else {
        T n = jsqrt(squared_n);
        if (std::abs(jval(w)) < eps) {
            two_atan_nbyw_by_n = T(jval(w) > 0 ? M_PI : -M_PI) / n;
        } else {
            two_atan_nbyw_by_n = T(2.0) * jatan(n / w) / n;
        }
    }
    for (int c = 0; c < 3; ++c) t[c] = two_atan_nbyw_by_n * q[c];
}
template <class T> void so3_exp(const T o[3], T q[4]) {  // SO3::expAndTheta
    const double eps = 1e-10;
    T theta_sq = o[0] * o[0] + (o[1] * o[1] + o[2] * o[2]);
    T imag, real;
    if (jval(theta_sq) < eps * eps) {
        T theta_po4 = theta_sq * theta_sq;
        imag = T(0.5) - T(1.0 / 48.0) * theta_sq + T(1.0 / 3840.0) * theta_po4;
        real = T(1.0) - T(1.0 / 8.0) * theta_sq + T(1.0 / 384.0) * theta_po4;
    } else {
        T theta = jsqrt(theta_sq);
        T half = T(0.5) * theta;
        imag = jsin(half) / theta;
        real = jcos(half);
    }
    for (int c = 0; c < 3; ++c) q[c] = imag * o[c];
    q[3] = real;
}

// CalibReprojectionError_SO3::operator(), EventCalibSpline.hpp:78-140.  rb: cumulative basis beta_1..3, tb: basis N_0..3
template <class T>
T residual_so3(const T *intr, const T *const r_cp[4], const T *const t_cp[4], const double obs[2], const double lm[3],
               double radius, const double rb[3], const double tb[4]) {
    T q[4] = {r_cp[0][0], r_cp[0][1], r_cp[0][2], r_cp[0][3]};  // Qwb = r_cp0            (:100)
    for (int j = 1; j < 4; ++j) {                                 // Qwb *= exp(b_j log(r_cp{j-1}^-1 r_cp{j}))   (:101-103)
        T inv[4], rel[4], lg[3], ex[4], nq[4];
        so3_inverse(r_cp[j - 1], inv);
        so3_mul(inv, r_cp[j], rel);
        so3_log(rel, lg);
        for (int c = 0; c < 3; ++c) lg[c] = rb[j - 1] * lg[c];
        so3_exp(lg, ex);
        so3_mul(q, ex, nq);
        for (int c = 0; c < 4; ++c) q[c] = nq[c];
    }
    T tw[3];
    for (int c = 0; c < 3; ++c) tw[c] = tb[0] * t_cp[0][c] + tb[1] * t_cp[1][c] + tb[2] * t_cp[2][c] + tb[3] * t_cp[3][c];
    T Xc[3];
    undistort(intr[0], intr[1], intr[2], intr[3], intr[4], intr[5], intr[6], intr[7], intr[8], obs, Xc);
    const T &qx = q[0], &qy = q[1], &qz = q[2], &qw = q[3];
    T tx = 2. * qx, ty = 2. * qy, tz = 2. * qz;
    T twx = tx * qw, twy = ty * qw, txx = tx * qx, txz = tz * qx, tyy = ty * qy, tyz = tz * qy;
    T R2[3] = {txz - twy, tyz + twx, 1.0 - (txx + tyy)};
    T depth = -tw[2] / (R2[0] * Xc[0] + (R2[1] * Xc[1] + R2[2] * Xc[2]));
    for (int c = 0; c < 3; ++c) Xc[c] = Xc[c] * depth;
    T uv[3] = {qy * Xc[2] - qz * Xc[1], qz * Xc[0] - qx * Xc[2], qx * Xc[1] - qy * Xc[0]};
    for (int c = 0; c < 3; ++c) uv[c] = uv[c] + uv[c];
    T cr[3] = {qy * uv[2] - qz * uv[1], qz * uv[0] - qx * uv[2], qx * uv[1] - qy * uv[0]};
    T d[3];
    for (int c = 0; c < 3; ++c) d[c] = ((Xc[c] + qw * uv[c]) + cr[c] + tw[c]) - lm[c];
    return jsqrt(d[0] * d[0] + (d[1] * d[1] + d[2] * d[2])) - radius;
}

// Sophus SO3::Dx_this_mul_exp_x_at_0 (4x3, row major; rows x y z w) = Jacobian of LocalParameterizationSO3 (BsplineSO3.hpp:209-217)
void so3_plus_jacobian(const double *x, double J[12]) {
    const double c = 0.5;
    J[0] = c * x[3];  J[1] = -c * x[2]; J[2] = c * x[1];
    J[3] = c * x[2];  J[4] = c * x[3];  J[5] = -c * x[0];
    J[6] = -c * x[1]; J[7] = c * x[0];  J[8] = c * x[3];
    J[9] = -c * x[0]; J[10] = -c * x[1]; J[11] = -c * x[2];
}

typedef Jet<37> J37;

// residual + 1x37 ambient Jacobian: [intrinsics 9 | r_cp0..3 (4 each) | t_cp0..3 (3 each)]
double residual_jac(const double *intr, const double *const rcp[4], const double *const tcp[4], const double obs[2],
                    const double lm[3], double radius, const double rb[4], const double tb[4], double jac[37], bool so3 = false) {
    std::vector<J37> P(37);
    for (int i = 0; i < 9; ++i) P[i].a = intr[i];
    for (int j = 0; j < 4; ++j)
        for (int c = 0; c < 4; ++c) P[9 + 4 * j + c].a = rcp[j][c];
    for (int j = 0; j < 4; ++j)
        for (int c = 0; c < 3; ++c) P[25 + 3 * j + c].a = tcp[j][c];
    for (int i = 0; i < 37; ++i) P[i].v[i] = 1.0;
    const J37 *r4[4] = {&P[9], &P[13], &P[17], &P[21]};
    const J37 *t4[4] = {&P[25], &P[28], &P[31], &P[34]};
    J37 r;
    if (so3) {  // cumulative basis of BsplineSO3::derBasisFuns (BsplineSO3.cpp:88-92)
        double beta[3];
        beta[2] = rb[3];
        beta[1] = beta[2] + rb[2];
        beta[0] = beta[1] + rb[1];
        r = residual_so3<J37>(P.data(), r4, t4, obs, lm, radius, beta, tb);
    } else {
        r = residual<J37>(P.data(), r4, t4, obs, lm, radius, rb, tb);
    }
    std::memcpy(jac, r.v, sizeof(double) * 37);
    return r.a;
}

// EigenQuaternionParameterization::ComputeJacobian (4x3, row major), storage x y z w [external: Ceres]
void quat_plus_jacobian(const double *x, double J[12]) {
    J[0] = x[3];  J[1] = x[2];   J[2] = -x[1];
    J[3] = -x[2]; J[4] = x[3];   J[5] = x[0];
    J[6] = x[1];  J[7] = -x[0];  J[8] = x[3];
    J[9] = -x[0]; J[10] = -x[1]; J[11] = -x[2];
}

struct Problem {
    int n_splines = 0;
    std::vector<int> n_cp;           // control points per spline
    std::vector<int> cp_off;         // offset of each spline's first control point in the flat arrays
    std::vector<int> span_off;       // offset of each spline's first span block (n_cp-3 spans per spline)
    std::vector<std::vector<double>> knots;
    double radius = 1.75, huber = 0.35;
    bool so3 = false;  // useSO3: CalibReprojectionError_SO3 + LocalParameterizationSO3
    // residual records
    std::vector<double> obs, lm, basis;
    std::vector<int> span;    // global span block index
    std::vector<int> cp0;     // global index of the first of the 4 control points
    int total_cp() const { return cp_off.empty() ? 0 : cp_off.back() + n_cp.back(); }
    int total_spans() const { return span_off.empty() ? 0 : span_off.back() + n_cp.back() - 3; }
};

// One residual block as Ceres evaluates it: residual, local-parameterised Jacobian (1x33), Huber corrector.
// Returns the cost contribution 1/2 rho(r^2); r_out / J33 are the corrected residual / Jacobian.
double eval_block(const Problem &p, size_t k, const double *intr, const double *rot, const double *trans, double *r_out,
                  double J33[33]) {
    const int c0 = p.cp0[k];
    const double *rcp[4], *tcp[4];
    for (int j = 0; j < 4; ++j) {
        rcp[j] = rot + 4 * (c0 + j);
        tcp[j] = trans + 3 * (c0 + j);
    }
    double jac[37];
    const double *b = &p.basis[4 * k];
    double r = residual_jac(intr, rcp, tcp, &p.obs[2 * k], &p.lm[3 * k], p.radius, b, b, jac, p.so3);
    for (int i = 0; i < 9; ++i) J33[i] = jac[i];
    for (int j = 0; j < 4; ++j) {  // J_local = J_global(1x4) * PlusJacobian(4x3)
        double PJ[12];
        if (p.so3) so3_plus_jacobian(rcp[j], PJ);
        else quat_plus_jacobian(rcp[j], PJ);
        for (int c = 0; c < 3; ++c) {
            double s = 0;
            for (int a = 0; a < 4; ++a) s += jac[9 + 4 * j + a] * PJ[3 * a + c];
            J33[9 + 3 * j + c] = s;
        }
    }
    for (int i = 0; i < 12; ++i) J33[21 + i] = jac[25 + i];
    // HuberLoss(a): s = r^2; s <= a^2: rho = s, rho' = 1; else rho = 2 a sqrt(s) - a^2, rho' = a / sqrt(s)  [Ceres loss_function.cc]
    const double s = r * r, a2 = p.huber * p.huber;
    double rho, rho1;
    if (s > a2) {
        const double rt = std::sqrt(s);
        rho = 2 * p.huber * rt - a2;
        rho1 = std::max(std::numeric_limits<double>::min(), p.huber / rt);
    } else {
        rho = s;
        rho1 = 1.0;
    }
    // Corrector: rho'' <= 0 for Huber => scale residual and Jacobian by sqrt(rho')  [Ceres corrector.cc]
    const double sr = std::sqrt(rho1);
    *r_out = sr * r;
    for (int i = 0; i < 33; ++i) J33[i] *= sr;
    return 0.5 * rho;
}

}  // namespace

#include <limits>

extern "C" {

void orc_inverse_radial(const double *k4, double *b5) {  // PinholeCamera.cpp:70-95
    const double *k = k4;
    double k00 = k[0] * k[0], k000 = k[0] * k00, k0000 = k[0] * k000, k00000 = k[0] * k0000;
    double k01 = k[0] * k[1], k001 = k[0] * k01, k0001 = k[0] * k001, k11 = k[1] * k[1], k011 = k[0] * k11;
    double k02 = k[0] * k[2], k002 = k[0] * k02, k12 = k[1] * k[2], k03 = k[0] * k[3];
    b5[0] = -k[0];
    b5[1] = 3 * k00 - k[1];
    b5[2] = -12 * k000 + 8 * k01 - k[2];
    b5[3] = 55 * k0000 - 55 * k001 + 5 * k11 + 10 * k02 - k[3];
    b5[4] = -273 * k00000 + 364 * k0001 - 78 * k011 - 78 * k002 + 12 * k12 + 12 * k03;
}

// unit_test_inverseDistortion.cpp: undistort with the inverse polynomial, re-distort, pixel error
double orc_inverse_distortion_roundtrip(void) {
    double radial[4] = {-0.34991902, -0.014698517, 0.59684463, 0}, inv[5];
    orc_inverse_radial(radial, inv);
    const double f = 359.67525, cx = 172.5, cy = 129.5;
    double X[2] = {(50 - cx) / f, (50 - cy) / f};
    double r2 = X[0] * X[0] + X[1] * X[1], r4 = r2 * r2, r6 = r4 * r2, r8 = r6 * r2, r10 = r8 * r2;
    double s = 1 + (r2 * inv[0] + r4 * inv[1] + r6 * inv[2] + r8 * inv[3] + r10 * inv[4]);
    X[0] *= s;
    X[1] *= s;
    r2 = X[0] * X[0] + X[1] * X[1];
    r4 = r2 * r2;
    r6 = r4 * r2;
    double d = 1 + (r2 * radial[0] + r4 * radial[1] + r6 * radial[2]);
    X[0] *= d;
    X[1] *= d;
    double u = f * X[0] + cx - 50, v = f * X[1] + cy - 50;
    return std::sqrt(u * u + v * v);
}

size_t orc_find_span(const double *knots, int nk, double u) { return find_span(knots, (size_t) nk, u); }
void orc_basis(const double *knots, int nk, double u, int *span, double *N) {
    size_t s = find_span(knots, (size_t) nk, u);
    *span = (int) s;
    basis_funs(knots, s, u, N);
}

// knot vector of BsplineReal::approximation (BsplineReal.hpp:87-100): clamped ends, interior by eq. 9.68
void orc_knots(const double *us, int n_data, int n_cp, double *knots) {
    const int degree = 3;
    for (int i = 0; i <= degree; ++i) knots[i] = us[0];
    for (int i = 0; i <= degree; ++i) knots[n_cp + degree - i] = us[n_data - 1];
    double d = n_data / double(n_cp - degree);
    for (int j = 1; j <= n_cp - 1 - degree; j++) {
        int i = (int) std::floor(j * d);
        double alpha = j * d - i;
        knots[degree + j] = (1 - alpha) * us[i - 1] + alpha * us[i];
    }
}

// single residual + ambient Jacobian (for unit parity with the CUDA device function)
double orc_residual_jac(const double *intr, const double *rcp16, const double *tcp12, const double *obs, const double *lm,
                        double radius, const double *basis, double *jac37) {
    const double *r4[4] = {rcp16, rcp16 + 4, rcp16 + 8, rcp16 + 12}, *t4[4] = {tcp12, tcp12 + 3, tcp12 + 6, tcp12 + 9};
    return residual_jac(intr, r4, t4, obs, lm, radius, basis, basis, jac37);
}

// same for CalibReprojectionError_SO3 (basis = the 4 values N; the cumulative rotation basis is derived from them)
double orc_residual_jac_so3(const double *intr, const double *rcp16, const double *tcp12, const double *obs, const double *lm,
                            double radius, const double *basis, double *jac37) {
    const double *r4[4] = {rcp16, rcp16 + 4, rcp16 + 8, rcp16 + 12}, *t4[4] = {tcp12, tcp12 + 3, tcp12 + 6, tcp12 + 9};
    return residual_jac(intr, r4, t4, obs, lm, radius, basis, basis, jac37, true);
}
void orc_so3_plus_jacobian(const double *x, double *J12) { so3_plus_jacobian(x, J12); }

// ---- problem handle ----
void *orc_problem_create(int n_splines, const int *n_cp, const double *knots_concat, double radius, double huber) {
    Problem *p = new Problem();
    p->n_splines = n_splines;
    p->radius = radius;
    p->huber = huber;
    int co = 0, so = 0;
    const double *k = knots_concat;
    for (int s = 0; s < n_splines; ++s) {
        p->n_cp.push_back(n_cp[s]);
        p->cp_off.push_back(co);
        p->span_off.push_back(so);
        p->knots.emplace_back(k, k + n_cp[s] + 4);
        k += n_cp[s] + 4;
        co += n_cp[s];
        so += n_cp[s] - 3;
    }
    return p;
}
void orc_problem_free(void *h) { delete (Problem *) h; }
void orc_problem_set_so3(void *h, int so3) { ((Problem *) h)->so3 = so3 != 0; }
// LocalParameterizationSO3::Plus: T * exp(delta)
void orc_so3_plus(const double *x, const double *d, double *out) {
    double e[4];
    so3_exp<double>(d, e);
    so3_mul<double>(x, e, out);
}

// explicit residual records: obs (2), landmark (3), time, spline index. Basis / span computed like optimize() :173-179
void orc_problem_set_residuals(void *h, const double *obs, const double *lm, const double *t, const int *spline, int64_t n) {
    Problem &p = *(Problem *) h;
    p.obs.assign(obs, obs + 2 * n);
    p.lm.assign(lm, lm + 3 * n);
    p.basis.resize(4 * n);
    p.span.resize(n);
    p.cp0.resize(n);
    for (int64_t k = 0; k < n; ++k) {
        const std::vector<double> &kn = p.knots[spline[k]];
        size_t s = find_span(kn.data(), kn.size(), t[k]);
        basis_funs(kn.data(), s, t[k], &p.basis[4 * k]);
        p.span[k] = p.span_off[spline[k]] + (int) s - 3;
        p.cp0[k] = p.cp_off[spline[k]] + (int) s - 3;
    }
}

// association loop of EventCalibSpline::optimize (EventCalibSpline.cpp:157-192) + findCenter (CirclesEventFrame.hpp:50-65):
// for every raw event inside a spline's time range: nearest keyframe in time, |dt|^2 < (5 step)^2, nearest circle centre of
// that frame, | ||p-c|| - r | < 5 px  ->  residual record.  kf_circ: K x n_circ x 3 (cx, cy, r; r < 0 marks an absent feature),
// lm_xyz: n_circ x 3 board points.  Outputs the selected event indices and circle ids; returns the count.
int64_t orc_associate(void *h, const double *ev_t, const double *ev_x, const double *ev_y, int64_t n_ev, const double *kf_t,
                      const double *kf_circ, int K, int n_circ, const double *lm_xyz, double step, int64_t *out_event,
                      int *out_circle) {
    Problem &p = *(Problem *) h;
    std::vector<double> obs, lm, tt;
    std::vector<int> sp;
    int64_t cnt = 0;
    for (int s = 0; s < p.n_splines; ++s) {
        const std::vector<double> &kn = p.knots[s];
        const double u0 = kn.front(), u1 = kn.back();
        int64_t lo = std::lower_bound(ev_t, ev_t + n_ev, u0) - ev_t, hi = std::upper_bound(ev_t, ev_t + n_ev, u1) - ev_t;
        for (int64_t i = lo; i < hi; ++i) {
            const double u = ev_t[i];
            // nearest keyframe in time (1-NN; ties -> the earlier frame)
            int j = (int) (std::lower_bound(kf_t, kf_t + K, u) - kf_t);
            int best = j < K ? j : K - 1;
            if (j > 0 && (j >= K || (u - kf_t[j - 1]) <= (kf_t[j] - u))) best = j - 1;
            const double dt = u - kf_t[best];
            if (!(dt * dt < (5 * step * 5 * step))) continue;
            const double *c = kf_circ + (size_t) best * n_circ * 3;
            int bi = -1;
            double bd = 0;
            for (int q = 0; q < n_circ; ++q) {
                if (c[3 * q + 2] < 0) continue;
                const double dx = ev_x[i] - c[3 * q], dy = ev_y[i] - c[3 * q + 1], d2 = dx * dx + dy * dy;
                if (bi < 0 || d2 < bd) {
                    bi = q;
                    bd = d2;
                }
            }
            if (bi < 0) continue;
            if (!(std::abs(std::sqrt(bd) - c[3 * bi + 2]) < 5)) continue;
            if (out_event) out_event[cnt] = i;
            if (out_circle) out_circle[cnt] = bi;
            obs.push_back(ev_x[i]);
            obs.push_back(ev_y[i]);
            for (int a = 0; a < 3; ++a) lm.push_back(lm_xyz[3 * bi + a]);
            tt.push_back(u);
            sp.push_back(s);
            ++cnt;
        }
    }
    orc_problem_set_residuals(h, obs.data(), lm.data(), tt.data(), sp.data(), cnt);
    return cnt;
}

int64_t orc_problem_num_residuals(void *h) { return (int64_t) ((Problem *) h)->span.size(); }
int orc_problem_num_spans(void *h) { return ((Problem *) h)->total_spans(); }
void orc_problem_get_records(void *h, double *basis, int *span) {
    Problem &p = *(Problem *) h;
    if (basis) std::memcpy(basis, p.basis.data(), p.basis.size() * 8);
    if (span) std::memcpy(span, p.span.data(), p.span.size() * 4);
}

// cost only: sum of 1/2 rho(r^2)
double orc_cost(void *h, const double *intr, const double *rot, const double *trans) {
    Problem &p = *(Problem *) h;
    double cost = 0, r, J[33];
    for (size_t k = 0; k < p.span.size(); ++k) cost += eval_block(p, k, intr, rot, trans, &r, J);
    return cost;
}

// per-span normal equations: blocks[n_spans][33*33] (full symmetric), grads[n_spans][33] = J^T r, returns cost.
// Local order of a span block: intrinsics 9 | rot tangent of cp0..cp3 (3 each) | trans of cp0..cp3 (3 each).
// residuals_out / jac_out (optional): corrected residual and 1x33 Jacobian of every block.
double orc_normal_eq(void *h, const double *intr, const double *rot, const double *trans, double *blocks, double *grads,
                     double *residuals_out, double *jac_out) {
    Problem &p = *(Problem *) h;
    const int ns = p.total_spans();
    std::memset(blocks, 0, sizeof(double) * (size_t) ns * 1089);
    std::memset(grads, 0, sizeof(double) * (size_t) ns * 33);
    double cost = 0;
    for (size_t k = 0; k < p.span.size(); ++k) {
        double r, J[33];
        cost += eval_block(p, k, intr, rot, trans, &r, J);
        double *B = blocks + (size_t) p.span[k] * 1089, *g = grads + (size_t) p.span[k] * 33;
        for (int i = 0; i < 33; ++i) {
            for (int j = 0; j < 33; ++j) B[33 * i + j] += J[i] * J[j];
            g[i] += J[i] * r;
        }
        if (residuals_out) residuals_out[k] = r;
        if (jac_out) std::memcpy(jac_out + 33 * k, J, sizeof J);
    }
    return cost;
}

// multi-threaded evaluation for the CPU baseline (Ceres: options.num_threads = hardware_concurrency() - 2,
// EventCalibSpline.cpp:242): one Jacobian evaluation (normal equations) + one cost-only evaluation.
}  // extern "C"
#include <thread>
extern "C" double orc_eval_mt(void *h, const double *intr, const double *rot, const double *trans, int threads, double *cost_only) {
    Problem &p = *(Problem *) h;
    const int ns = p.total_spans();
    const size_t n = p.span.size();
    if (threads < 1) threads = 1;
    std::vector<std::vector<double>> H(threads), G(threads);
    std::vector<double> cost(threads, 0.0), cost2(threads, 0.0);
    auto work = [&](int t) {
        H[t].assign((size_t) ns * 1089, 0.0);
        G[t].assign((size_t) ns * 33, 0.0);
        const size_t b = n * t / threads, e = n * (t + 1) / threads;
        for (size_t k = b; k < e; ++k) {
            double r, J[33];
            cost[t] += eval_block(p, k, intr, rot, trans, &r, J);
            double *B = H[t].data() + (size_t) p.span[k] * 1089, *g = G[t].data() + (size_t) p.span[k] * 33;
            for (int i = 0; i < 33; ++i) {
                for (int j = 0; j < 33; ++j) B[33 * i + j] += J[i] * J[j];
                g[i] += J[i] * r;
            }
        }
        // cost-only pass (Evaluate without Jacobians: plain doubles)
        for (size_t k = b; k < e; ++k) {
            const int c0 = p.cp0[k];
            const double *rcp[4], *tcp[4];
            for (int j = 0; j < 4; ++j) {
                rcp[j] = rot + 4 * (c0 + j);
                tcp[j] = trans + 3 * (c0 + j);
            }
            const double *bb = &p.basis[4 * k];
            double r = residual<double>(intr, rcp, tcp, &p.obs[2 * k], &p.lm[3 * k], p.radius, bb, bb);
            const double s = r * r, a2 = p.huber * p.huber;
            cost2[t] += 0.5 * (s > a2 ? 2 * p.huber * std::sqrt(s) - a2 : s);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < threads; ++t) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
    double c = 0, c2 = 0;
    for (int t = 0; t < threads; ++t) {
        c += cost[t];
        c2 += cost2[t];
    }
    if (cost_only) *cost_only = c2;
    return c;
}
extern "C" {

// EigenQuaternionParameterization::Plus (x_plus = [sin|d|/|d| d, cos|d|] (x) x), Eigen product order  [external: Ceres]
void orc_quat_plus(const double *x, const double *d, double *out) {
    const double nd = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    if (nd > 0.0) {
        const double s = std::sin(nd) / nd;
        const double qd[4] = {s * d[0], s * d[1], s * d[2], std::cos(nd)};  // x y z w
        // Eigen quaternion product a*b
        const double ax = qd[0], ay = qd[1], az = qd[2], aw = qd[3], bx = x[0], by = x[1], bz = x[2], bw = x[3];
        out[3] = aw * bw - ax * bx - ay * by - az * bz;
        out[0] = aw * bx + ax * bw + ay * bz - az * by;
        out[1] = aw * by + ay * bw + az * bx - ax * bz;
        out[2] = aw * bz + az * bw + ax * by - ay * bx;
    } else {
        for (int i = 0; i < 4; ++i) out[i] = x[i];
    }
}
}
