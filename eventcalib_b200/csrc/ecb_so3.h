// Cumulative SO(3) B-spline rotation of CalibReprojectionError_SO3 (event_camera_calib/include/opengv2/event_camera_calib/
// EventCalibSpline.hpp:100-105) and its tangent-space derivative:
//      Qwb = R0 * exp(b1 log(R0^-1 R1)) * exp(b2 log(R1^-1 R2)) * exp(b3 log(R2^-1 R3))
// with the cumulative basis b_j of BsplineSO3::derBasisFuns (core/spline/src/BsplineSO3.cpp:73-95) and
// LocalParameterizationSO3 (core/spline/include/opengv2/spline/BsplineSO3.hpp:190-221: Plus = T * exp(delta),
// Jacobian = Dx_this_mul_exp_x_at_0).  Sophus is not in /root/reference; exp / log follow Sophus 1.x so3.hpp
// (expAndTheta / logAndTheta incl. their small-angle Taylor branches, epsilon 1e-10)  [external].
//
// Ceres differentiates the functor with Jets in the 4-coefficient ambient space and multiplies by the 4x3 Plus Jacobian;
// here the same function is differentiated directly along the tangent directions with small forward-mode dual numbers
// (3 partials, one control point at a time) — identical by the chain rule.  Host + device.
#pragma once
#include <math.h>

#ifndef ECB_HD
#ifdef __CUDACC__
#define ECB_HD __host__ __device__ __forceinline__
#else
#define ECB_HD inline
#endif
#endif

namespace ecb_so3 {

constexpr double kEps = 1e-10;  // Sophus::Constants<double>::epsilon()
constexpr double kPi = 3.141592653589793238462643383279502884;

struct D3 {  // value + 3 partials
    double a, d[3];
};
ECB_HD D3 mk(double a) { return D3{a, {0.0, 0.0, 0.0}}; }
ECB_HD D3 operator+(const D3 &f, const D3 &g) { return D3{f.a + g.a, {f.d[0] + g.d[0], f.d[1] + g.d[1], f.d[2] + g.d[2]}}; }
ECB_HD D3 operator-(const D3 &f, const D3 &g) { return D3{f.a - g.a, {f.d[0] - g.d[0], f.d[1] - g.d[1], f.d[2] - g.d[2]}}; }
ECB_HD D3 operator-(const D3 &f) { return D3{-f.a, {-f.d[0], -f.d[1], -f.d[2]}}; }
ECB_HD D3 operator*(const D3 &f, const D3 &g) {
    return D3{f.a * g.a, {f.a * g.d[0] + f.d[0] * g.a, f.a * g.d[1] + f.d[1] * g.a, f.a * g.d[2] + f.d[2] * g.a}};
}
ECB_HD D3 operator*(double s, const D3 &f) { return D3{s * f.a, {s * f.d[0], s * f.d[1], s * f.d[2]}}; }
ECB_HD D3 operator/(const D3 &f, const D3 &g) {
    const double gi = 1.0 / g.a, q = f.a * gi;
    return D3{q, {(f.d[0] - q * g.d[0]) * gi, (f.d[1] - q * g.d[1]) * gi, (f.d[2] - q * g.d[2]) * gi}};
}
ECB_HD D3 operator-(double s, const D3 &f) { return mk(s) - f; }
ECB_HD D3 operator/(double s, const D3 &f) { return mk(s) / f; }
ECB_HD D3 xsqrt(const D3 &f) {
    const double r = sqrt(f.a), t = 0.5 / r;
    return D3{r, {f.d[0] * t, f.d[1] * t, f.d[2] * t}};
}
ECB_HD D3 xsin(const D3 &f) {
    const double c = cos(f.a);
    return D3{sin(f.a), {c * f.d[0], c * f.d[1], c * f.d[2]}};
}
ECB_HD D3 xcos(const D3 &f) {
    const double s = -sin(f.a);
    return D3{cos(f.a), {s * f.d[0], s * f.d[1], s * f.d[2]}};
}
ECB_HD D3 xatan(const D3 &f) {
    const double t = 1.0 / (1.0 + f.a * f.a);
    return D3{atan(f.a), {t * f.d[0], t * f.d[1], t * f.d[2]}};
}
ECB_HD double val(const D3 &f) { return f.a; }
ECB_HD double mk_d(double a) { return a; }
ECB_HD double xsqrt(double f) { return sqrt(f); }
ECB_HD double xsin(double f) { return sin(f); }
ECB_HD double xcos(double f) { return cos(f); }
ECB_HD double xatan(double f) { return atan(f); }
ECB_HD double val(double f) { return f; }

template <class S> struct Lift;
template <> struct Lift<double> { static ECB_HD double of(double a) { return a; } };
template <> struct Lift<D3> { static ECB_HD D3 of(double a) { return mk(a); } };

template <class S>
struct Quat {  // storage order of Eigen::Quaternion::coeffs(): x y z w
    S x, y, z, w;
};

template <class S>
ECB_HD Quat<S> qmul(const Quat<S> &a, const Quat<S> &b) {  // Sophus SO3 operator*
    Quat<S> r;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
    return r;
}
template <class S>
ECB_HD Quat<S> qinv(const Quat<S> &a) {  // SO3::inverse = conjugate
    return Quat<S>{-a.x, -a.y, -a.z, a.w};
}

// Sophus SO3::logAndTheta
template <class S>
ECB_HD void so3_log(const Quat<S> &q, S t[3]) {
    const S squared_n = q.x * q.x + q.y * q.y + q.z * q.z;
    const S w = q.w;
    S two_atan_nbyw_by_n;
    if (val(squared_n) < kEps * kEps) {
        const S squared_w = w * w;
        two_atan_nbyw_by_n = 2.0 / w - (2.0 / 3.0) * (squared_n / (w * squared_w));
    } else {
        const S n = xsqrt(squared_n);
        if (fabs(val(w)) < kEps) {
            two_atan_nbyw_by_n = (val(w) > 0.0 ? kPi : -kPi) / n;
        } else {
            two_atan_nbyw_by_n = 2.0 * xatan(n / w) / n;
        }
    }
    t[0] = two_atan_nbyw_by_n * q.x;
    t[1] = two_atan_nbyw_by_n * q.y;
    t[2] = two_atan_nbyw_by_n * q.z;
}

// Sophus SO3::expAndTheta
template <class S>
ECB_HD Quat<S> so3_exp(const S o[3]) {
    const S theta_sq = o[0] * o[0] + o[1] * o[1] + o[2] * o[2];
    S imag, real;
    if (val(theta_sq) < kEps * kEps) {
        const S theta_po4 = theta_sq * theta_sq;
        imag = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * theta_po4;
        real = 1.0 - (1.0 / 8.0) * theta_sq + (1.0 / 384.0) * theta_po4;
    } else {
        const S theta = xsqrt(theta_sq);
        const S half = 0.5 * theta;
        imag = xsin(half) / theta;
        real = xcos(half);
    }
    return Quat<S>{imag * o[0], imag * o[1], imag * o[2], real};
}

// EventCalibSpline.hpp:100-105; beta = cumulative basis b1..b3
template <class S>
ECB_HD Quat<S> spline_rotation(const Quat<S> R[4], const double beta[3]) {
    Quat<S> q = R[0];
#pragma unroll
    for (int j = 1; j < 4; ++j) {
        S t[3];
        so3_log(qmul(qinv(R[j - 1]), R[j]), t);
        t[0] = beta[j - 1] * t[0];
        t[1] = beta[j - 1] * t[1];
        t[2] = beta[j - 1] * t[2];
        q = qmul(q, so3_exp(t));
    }
    return q;
}

// cumulative basis of BsplineSO3::derBasisFuns (BsplineSO3.cpp:88-92) from the 4 basis values N_{i-3..i}
ECB_HD void cumulative_basis(const double N[4], double beta[3]) {
    beta[2] = N[3];
    beta[1] = beta[2] + N[2];
    beta[0] = beta[1] + N[1];
}

// value (double) of the spline rotation; Q: 4 control points (x y z w each)
ECB_HD void rotation_value(const double *Q, const double beta[3], double q[4]) {
    Quat<double> R[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) R[j] = Quat<double>{Q[4 * j], Q[4 * j + 1], Q[4 * j + 2], Q[4 * j + 3]};
    const Quat<double> r = spline_rotation<double>(R, beta);
    q[0] = r.x;
    q[1] = r.y;
    q[2] = r.z;
    q[3] = r.w;
}

// derivative of the spline rotation's 4 coefficients along the 3 tangent directions of control point j
// (T_j -> T_j * exp(delta), delta = e_k):  dq[c][k]
ECB_HD void rotation_tangent(const double *Q, const double beta[3], int j, double dq[4][3]) {
    Quat<D3> R[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) R[i] = Quat<D3>{mk(Q[4 * i]), mk(Q[4 * i + 1]), mk(Q[4 * i + 2]), mk(Q[4 * i + 3])};
    {  // seed: d/d delta_k of  T_j * (delta/2, 1)  =  Dx_this_mul_exp_x_at_0 (Sophus)
        const double x = Q[4 * j], y = Q[4 * j + 1], z = Q[4 * j + 2], w = Q[4 * j + 3];
        const double sx[3] = {0.5 * w, -0.5 * z, 0.5 * y}, sy[3] = {0.5 * z, 0.5 * w, -0.5 * x};
        const double sz[3] = {-0.5 * y, 0.5 * x, 0.5 * w}, sw[3] = {-0.5 * x, -0.5 * y, -0.5 * z};
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i == j) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    R[i].x.d[k] = sx[k];
                    R[i].y.d[k] = sy[k];
                    R[i].z.d[k] = sz[k];
                    R[i].w.d[k] = sw[k];
                }
            }
    }
    const Quat<D3> r = spline_rotation<D3>(R, beta);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        dq[0][k] = r.x.d[k];
        dq[1][k] = r.y.d[k];
        dq[2][k] = r.z.d[k];
        dq[3][k] = r.w.d[k];
    }
}

// LocalParameterizationSO3::Plus: T * exp(delta)   (BsplineSO3.hpp:196-203)
ECB_HD void plus(const double *T, const double *delta, double *out) {
    const Quat<double> t{T[0], T[1], T[2], T[3]};
    const Quat<double> r = qmul(t, so3_exp<double>(delta));
    out[0] = r.x;
    out[1] = r.y;
    out[2] = r.z;
    out[3] = r.w;
}

}  // namespace ecb_so3
