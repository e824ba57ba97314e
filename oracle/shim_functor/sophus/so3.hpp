// TEST INFRASTRUCTURE.  Stand-in for <sophus/so3.hpp>.  Sophus is an external dependency of the reference (find_package(Sophus),
// CMakeLists.txt:46, no version) that is absent from /root/reference and from this image.  The reference's SO(3) path —
// CalibReprojectionError_SO3 (EventCalibSpline.hpp:65-156), BsplineSO3 (core/spline/src/BsplineSO3.cpp) and
// LocalParameterizationSO3 (BsplineSO3.hpp:190-221) — is compiled where it lies against THIS header, so what is pinned is the
// reference's own text; the Sophus semantics below are restated from Sophus 1.0's published so3.hpp [external]:
//   * storage: one Eigen::Quaternion<Scalar> (coefficients x y z w); num_parameters 4, DoF 3
//   * SO3(quaternion) normalises (coeffs /= norm); the copy from another SO3 / Map does not
//   * a * b: the quaternion product written out (w, x, y, z rows), result through the normalising constructor;
//     a *= b is a = a * b; inverse() = SO3(conjugate)
//   * exp / log: expAndTheta / logAndTheta with the Taylor branches below Constants<Scalar>::epsilon() = 1e-10
//   * Dx_this_mul_exp_x_at_0: the 4 x 3 Jacobian of q * exp(x) at x = 0
// Nothing else of Sophus exists here.  PARITY UNPINNED at this boundary.
#ifndef ECB_ORACLE_SOPHUS_SO3_SHIM
#define ECB_ORACLE_SOPHUS_SO3_SHIM
#include <Eigen/Eigen>
#include <cmath>

namespace Sophus {
template <class T> using Vector3 = Eigen::Matrix<T, 3, 1>;
template <class T> using Vector4 = Eigen::Matrix<T, 4, 1>;
template <class T> struct Constants {
    static T epsilon() { return T(1e-10); }
    static T pi() { return T(3.141592653589793238462643383279502884); }
};

template <class T>
struct SO3 {
    typedef T Scalar;
    typedef Eigen::Matrix<T, 3, 1> Tangent;
    static constexpr int num_parameters = 4;
    static constexpr int DoF = 3;
    Eigen::Quaternion<T> q;

    SO3() : q(T(1.0), T(0.0), T(0.0), T(0.0)) {}
    struct Raw {};  // copy without normalisation (copy construction from another SO3 / a Map)
    SO3(const Eigen::Quaternion<T> &q_, Raw) : q(q_) {}
    explicit SO3(const Eigen::Quaternion<T> &q_) : q(q_) { normalize(); }
    void normalize() {
        const T length = q.coeffs().norm();
        q.coeffs() /= length;
    }
    const Eigen::Quaternion<T> &unit_quaternion() const { return q; }
    void setQuaternion(const Eigen::Quaternion<T> &quat) {
        q = quat;
        normalize();
    }
    T *data() { return q.coeffs().data(); }
    const T *data() const { return q.coeffs().data(); }
    template <class N> SO3<N> cast() const {
        return SO3<N>(Eigen::Quaternion<N>(N(q.w()), N(q.x()), N(q.y()), N(q.z())), typename SO3<N>::Raw());
    }
    SO3 inverse() const { return SO3(q.conjugate()); }
    SO3 operator*(const SO3 &other) const {
        const Eigen::Quaternion<T> &a = q, &b = other.q;
        return SO3(Eigen::Quaternion<T>(a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z(),
                                        a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(),
                                        a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(),
                                        a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x()));
    }
    SO3 &operator*=(const SO3 &other) {
        *this = *this * other;
        return *this;
    }
    // logAndTheta
    Tangent log() const {
        using std::abs;
        using std::atan;
        using std::sqrt;
        const Tangent vec(q.x(), q.y(), q.z());
        const T squared_n = vec.squaredNorm();
        const T w = q.w();
        T two_atan_nbyw_by_n;
        if (squared_n < Constants<T>::epsilon() * Constants<T>::epsilon()) {
            const T squared_w = w * w;
            two_atan_nbyw_by_n = T(2.0) / w - T(2.0 / 3.0) * (squared_n) / (w * squared_w);
        } else {
            const T n = sqrt(squared_n);
            if (abs(w) < Constants<T>::epsilon()) {
                if (w > T(0.0))
                    two_atan_nbyw_by_n = Constants<T>::pi() / n;
                else
                    two_atan_nbyw_by_n = -Constants<T>::pi() / n;
            } else {
                two_atan_nbyw_by_n = T(2.0) * atan(n / w) / n;
            }
        }
        return two_atan_nbyw_by_n * vec;
    }
    // expAndTheta
    static SO3 exp(const Tangent &omega) {
        using std::cos;
        using std::sin;
        using std::sqrt;
        const T theta_sq = omega.squaredNorm();
        T imag_factor, real_factor;
        if (theta_sq < Constants<T>::epsilon() * Constants<T>::epsilon()) {
            const T theta_po4 = theta_sq * theta_sq;
            imag_factor = T(0.5) - T(1.0 / 48.0) * theta_sq + T(1.0 / 3840.0) * theta_po4;
            real_factor = T(1.0) - T(1.0 / 8.0) * theta_sq + T(1.0 / 384.0) * theta_po4;
        } else {
            const T theta = sqrt(theta_sq);
            const T half_theta = T(0.5) * theta;
            const T sin_half_theta = sin(half_theta);
            imag_factor = sin_half_theta / theta;
            real_factor = cos(half_theta);
        }
        return SO3(Eigen::Quaternion<T>(real_factor, imag_factor * omega[0], imag_factor * omega[1], imag_factor * omega[2]), Raw());
    }
    // rotation matrix (= the adjoint of SO(3))
    Eigen::Matrix<T, 3, 3> matrix() const {
        Eigen::Matrix<T, 3, 3> R;
        const T tx = T(2.0) * q.x(), ty = T(2.0) * q.y(), tz = T(2.0) * q.z();
        const T twx = tx * q.w(), twy = ty * q.w(), twz = tz * q.w();
        const T txx = tx * q.x(), txy = ty * q.x(), txz = tz * q.x();
        const T tyy = ty * q.y(), tyz = tz * q.y(), tzz = tz * q.z();
        R(0, 0) = T(1.0) - (tyy + tzz);
        R(0, 1) = txy - twz;
        R(0, 2) = txz + twy;
        R(1, 0) = txy + twz;
        R(1, 1) = T(1.0) - (txx + tzz);
        R(1, 2) = tyz - twx;
        R(2, 0) = txz - twy;
        R(2, 1) = tyz + twx;
        R(2, 2) = T(1.0) - (txx + tyy);
        return R;
    }
    Eigen::Matrix<T, 3, 3> Adj() const { return matrix(); }
    static Tangent lieBracket(const Tangent &a, const Tangent &b) {
        return Tangent(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
    }
    Eigen::Matrix<T, 4, 3> Dx_this_mul_exp_x_at_0() const {
        Eigen::Matrix<T, 4, 3> J;
        const T c0 = T(0.5) * q.w();
        const T c1 = T(0.5) * q.z();
        const T c2 = -c1;
        const T c3 = T(0.5) * q.y();
        const T c4 = T(0.5) * q.x();
        const T c5 = -c4;
        const T c6 = -c3;
        J(0, 0) = c0; J(0, 1) = c2; J(0, 2) = c3;
        J(1, 0) = c1; J(1, 1) = c0; J(1, 2) = c5;
        J(2, 0) = c6; J(2, 1) = c4; J(2, 2) = c0;
        J(3, 0) = c5; J(3, 1) = c6; J(3, 2) = c2;
        return J;
    }
};
typedef SO3<double> SO3d;
}  // namespace Sophus

namespace Eigen {
// Map<SO3 const>: a view of 4 caller scalars as a rotation (a plain copy here: the memory is not written through it)
template <class T> struct Map<Sophus::SO3<T> const> : Sophus::SO3<T> {
    explicit Map(const T *p) : Sophus::SO3<T>(Quaternion<T>(p), typename Sophus::SO3<T>::Raw()) {}
};
// Map<SO3>: writable view — assignment stores the coefficients back to the caller's memory
template <class T> struct Map<Sophus::SO3<T>> : Sophus::SO3<T> {
    T *raw;
    explicit Map(T *p) : Sophus::SO3<T>(Quaternion<T>(p), typename Sophus::SO3<T>::Raw()), raw(p) {}
    Map &operator=(const Sophus::SO3<T> &o) {
        this->q = o.q;
        for (int i = 0; i < 4; ++i) raw[i] = o.q.coeffs()[i];
        return *this;
    }
};
}  // namespace Eigen
#endif
