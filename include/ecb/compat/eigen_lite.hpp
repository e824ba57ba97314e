// The few Eigen types the reference's public signatures on the hot path mention (CirclesEventFrame.hpp:43-65,
// EventCalibSpline.hpp:19): Vector2d / Vector3d / Matrix3d / Quaterniond and Ref<const T>.  With the real Eigen on the include
// path it is used as is; otherwise these stand-ins carry the same spelling (operator[] / operator() / x() y() z() w(),
// Quaterniond(w, x, y, z), column-major Matrix3d) so that reference call sites compile unchanged against the façade.
// No linear algebra lives here — the façade converts to plain arrays at the boundary.
#ifndef ECB_COMPAT_EIGEN_LITE_HPP
#define ECB_COMPAT_EIGEN_LITE_HPP

#if __has_include(<Eigen/Core>) && !defined(ECB_FORCE_EIGEN_LITE)
#include <Eigen/Core>
#include <Eigen/Geometry>
#else
#include <cmath>
namespace Eigen {
template <int N>
struct VectorNd {
    double v[N];
    VectorNd() {
        for (int i = 0; i < N; ++i) v[i] = 0;
    }
    VectorNd(double a, double b) : v{a, b} { static_assert(N == 2, "two coefficients"); }
    VectorNd(double a, double b, double c) : v{a, b, c} { static_assert(N == 3, "three coefficients"); }
    explicit VectorNd(const double *p) {
        for (int i = 0; i < N; ++i) v[i] = p[i];
    }
    double &operator[](int i) { return v[i]; }
    const double &operator[](int i) const { return v[i]; }
    double &operator()(int i) { return v[i]; }
    const double &operator()(int i) const { return v[i]; }
    double *data() { return v; }
    const double *data() const { return v; }
    double norm() const {
        double s = 0;
        for (int i = 0; i < N; ++i) s += v[i] * v[i];
        return std::sqrt(s);
    }
};
typedef VectorNd<2> Vector2d;
typedef VectorNd<3> Vector3d;

struct Matrix3d {  // column-major like Eigen's default
    double m[9];
    Matrix3d() {
        for (double &x : m) x = 0;
    }
    static Matrix3d Identity() {
        Matrix3d r;
        r.m[0] = r.m[4] = r.m[8] = 1;
        return r;
    }
    double &operator()(int r, int c) { return m[c * 3 + r]; }
    const double &operator()(int r, int c) const { return m[c * 3 + r]; }
    const double *data() const { return m; }
};

struct Quaterniond {
    double c[4];  // coeffs(): x y z w
    Quaterniond() : c{0, 0, 0, 1} {}
    Quaterniond(double w, double x, double y, double z) : c{x, y, z, w} {}
    double x() const { return c[0]; }
    double y() const { return c[1]; }
    double z() const { return c[2]; }
    double w() const { return c[3]; }
    const double *coeffs_data() const { return c; }
    Quaterniond conjugate() const { return Quaterniond(c[3], -c[0], -c[1], -c[2]); }
    Matrix3d toRotationMatrix() const {
        Matrix3d R;
        const double X = c[0], Y = c[1], Z = c[2], W = c[3];
        R(0, 0) = 1 - 2 * (Y * Y + Z * Z), R(0, 1) = 2 * (X * Y - Z * W), R(0, 2) = 2 * (X * Z + Y * W);
        R(1, 0) = 2 * (X * Y + Z * W), R(1, 1) = 1 - 2 * (X * X + Z * Z), R(1, 2) = 2 * (Y * Z - X * W);
        R(2, 0) = 2 * (X * Z - Y * W), R(2, 1) = 2 * (Y * Z + X * W), R(2, 2) = 1 - 2 * (X * X + Y * Y);
        return R;
    }
};

template <class T>
struct Ref;
template <class T>
struct Ref<const T> {  // Ref<const Matrix3d>, Ref<const Vector3d>: a view of an existing object
    const T &r;
    Ref(const T &t) : r(t) {}
    operator const T &() const { return r; }
    const double &operator()(int i, int j) const { return r(i, j); }
    const double &operator()(int i) const { return r(i); }
    const double &operator[](int i) const { return r[i]; }
};
}  // namespace Eigen
#endif
#endif  // ECB_COMPAT_EIGEN_LITE_HPP
