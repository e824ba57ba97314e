// Event-to-circle residual of the dynamic-calibration objective and its closed-form tangent-space Jacobian.
//
// Replaces, for one residual block, what Ceres does around CalibReprojectionError::operator()
// (event_camera_calib/include/opengv2/event_camera_calib/EventCalibSpline.hpp:168-229, unDistort :36-63):
//   AutoDiff Jet<double,37> evaluation        -> analytic chain rule below (same function, exact derivative)
//   EigenQuaternionParameterization (4 -> 3)  -> J_local = J_ambient * PlusJacobian(Q_j)          [external: Ceres]
//   HuberLoss(delta) + Corrector              -> residual and Jacobian scaled by sqrt(rho'), cost = rho/2  [external: Ceres]
//
// Local parameter order of a block (33): intrinsics fx fy cx cy k1..k5 (EventCalibSpline.hpp:24-34) |
// rotation tangent of control points i-3..i (3 each) | translation of control points i-3..i (3 each).
//
// The header is plain C++ so that the host build (tests) can check it against the dual-number oracle without a GPU.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define ECB_HD __host__ __device__ __forceinline__
#else
#define ECB_HD inline
#endif

struct EcbResidualOut {
    double res;    // corrected residual  sqrt(rho') * r
    double cost;   // rho(r^2) / 2
    double raw;    // uncorrected residual r
};

// Q: 4 rotation control points (x,y,z,w each), T: 4 translation control points, b: 4 basis values.
// want_jac == false skips the Jacobian (cost-only evaluation).
// JS: stride between consecutive Jacobian entries in J (lets a kernel scatter straight into a shared-memory tile).
template <bool WANT_JAC, int JS = 1>
ECB_HD EcbResidualOut ecb_residual(const double *intr, const double *Q, const double *T, const double *b, double ou,
                                   double ov, double lx, double ly, double lz, double radius, double huber, double *J) {
    // spline evaluation (:195-201)
    double q0 = b[0] * Q[0] + b[1] * Q[4] + b[2] * Q[8] + b[3] * Q[12];
    double q1 = b[0] * Q[1] + b[1] * Q[5] + b[2] * Q[9] + b[3] * Q[13];
    double q2 = b[0] * Q[2] + b[1] * Q[6] + b[2] * Q[10] + b[3] * Q[14];
    double q3 = b[0] * Q[3] + b[1] * Q[7] + b[2] * Q[11] + b[3] * Q[15];
    const double nz = sqrt((q0 * q0 + q1 * q1) + (q2 * q2 + q3 * q3));
    const double inz = 1.0 / nz;  // one reciprocal instead of four divisions (differs from Eigen's /= by <= 1 ulp)
    q0 *= inz; q1 *= inz; q2 *= inz; q3 *= inz;  // x y z w
    const double t0 = b[0] * T[0] + b[1] * T[3] + b[2] * T[6] + b[3] * T[9];
    const double t1 = b[0] * T[1] + b[1] * T[4] + b[2] * T[7] + b[3] * T[10];
    const double t2 = b[0] * T[2] + b[1] * T[5] + b[2] * T[8] + b[3] * T[11];
    // unDistort (:36-63)
    const double fx = intr[0], fy = intr[1], cx = intr[2], cy = intr[3];
    const double ifx = 1.0 / fx, ify = 1.0 / fy;
    const double x = (ou - cx) * ifx, y = (ov - cy) * ify;
    const double r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2, r8 = r6 * r2, r10 = r8 * r2;
    const double s = 1.0 + intr[4] * r2 + intr[5] * r4 + intr[6] * r6 + intr[7] * r8 + intr[8] * r10;
    const double X0 = x * s, X1 = y * s;  // X2 = 1
    // third row of R(q) and the ray-plane depth (:213-223)
    const double R30 = 2.0 * (q0 * q2 - q3 * q1), R31 = 2.0 * (q1 * q2 + q3 * q0), R32 = 1.0 - 2.0 * (q0 * q0 + q1 * q1);
    const double den = R30 * X0 + R31 * X1 + R32;
    const double iden = 1.0 / den;
    const double lam = -t2 * iden;
    const double v0 = lam * X0, v1 = lam * X1, v2 = lam;
    // Xw = q * v + t  (Eigen: uv = 2 q.vec x v; v + w uv + q.vec x uv)  (:224-226)
    const double uv0 = 2.0 * (q1 * v2 - q2 * v1), uv1 = 2.0 * (q2 * v0 - q0 * v2), uv2 = 2.0 * (q0 * v1 - q1 * v0);
    const double d0 = v0 + q3 * uv0 + (q1 * uv2 - q2 * uv1) + t0 - lx;
    const double d1 = v1 + q3 * uv1 + (q2 * uv0 - q0 * uv2) + t1 - ly;
    const double d2 = v2 + q3 * uv2 + (q0 * uv1 - q1 * uv0) + t2 - lz;
    const double nrm = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
    const double r = nrm - radius;  // (:227)
    // Huber + corrector
    const double s2 = r * r, a2 = huber * huber;
    double rho = s2, rho1 = 1.0;
    if (s2 > a2) {
        const double ar = fabs(r);
        rho = 2.0 * huber * ar - a2;
        rho1 = huber / ar;
    }
    const double sr = sqrt(rho1);
    EcbResidualOut o;
    o.res = sr * r;
    o.cost = 0.5 * rho;
    o.raw = r;
    if (!WANT_JAC) return o;

    const double inrm = 1.0 / nrm;
    const double w0 = d0 * inrm, w1 = d1 * inrm, w2 = d2 * inrm;
    // f(q, X) = rotated un-scaled ray; a = w . f(q,X)
    const double ux0 = 2.0 * (q1 - q2 * X1), ux1 = 2.0 * (q2 * X0 - q0), ux2 = 2.0 * (q0 * X1 - q1 * X0);
    const double f0 = X0 + q3 * ux0 + (q1 * ux2 - q2 * ux1);
    const double f1 = X1 + q3 * ux1 + (q2 * ux0 - q0 * ux2);
    const double f2 = 1.0 + q3 * ux2 + (q0 * ux1 - q1 * ux0);
    const double a = w0 * f0 + w1 * f1 + w2 * f2;
    const double c1 = -a * lam * iden;
    // translation gradient
    const double gt0 = w0, gt1 = w1, gt2 = w2 - a * iden;
    // m = R^T w  (rotation by the conjugate): w + 2 qw (w x u) + 2 u x (u x w)
    const double wu0 = w1 * q2 - w2 * q1, wu1 = w2 * q0 - w0 * q2, wu2 = w0 * q1 - w1 * q0;  // w x u
    // u x (u x w) = -(u x (w x u))
    const double m0 = w0 + 2.0 * q3 * wu0 - 2.0 * (q1 * wu2 - q2 * wu1);
    const double m1 = w1 + 2.0 * q3 * wu1 - 2.0 * (q2 * wu0 - q0 * wu2);
    // gradient w.r.t. the undistorted ray X (components 0,1)
    const double gX0 = lam * m0 + c1 * R30, gX1 = lam * m1 + c1 * R31;
    const double sp = intr[4] + 2.0 * intr[5] * r2 + 3.0 * intr[6] * r4 + 4.0 * intr[7] * r6 + 5.0 * intr[8] * r8;
    const double Gx = gX0 * (s + 2.0 * x * x * sp) + gX1 * (2.0 * x * y * sp);
    const double Gy = gX0 * (2.0 * x * y * sp) + gX1 * (s + 2.0 * y * y * sp);
    const double dot = gX0 * x + gX1 * y;
    J[0 * JS] = sr * (-Gx * x * ifx);
    J[1 * JS] = sr * (-Gy * y * ify);
    J[2 * JS] = sr * (-Gx * ifx);
    J[3 * JS] = sr * (-Gy * ify);
    J[4 * JS] = sr * dot * r2;
    J[5 * JS] = sr * dot * r4;
    J[6 * JS] = sr * dot * r6;
    J[7 * JS] = sr * dot * r8;
    J[8 * JS] = sr * dot * r10;
    // gradient w.r.t. the unit quaternion (polynomial forms of R3 and q*v, as coded in the reference)
    const double uxv0 = 0.5 * uv0, uxv1 = 0.5 * uv1, uxv2 = 0.5 * uv2;  // u x v
    // grad_u A = 2 qw (v x w) + 2 ((u x v) x w) + 2 (v x (w x u))
    const double vw0 = v1 * w2 - v2 * w1, vw1 = v2 * w0 - v0 * w2, vw2 = v0 * w1 - v1 * w0;
    const double e0 = uxv1 * w2 - uxv2 * w1, e1 = uxv2 * w0 - uxv0 * w2, e2 = uxv0 * w1 - uxv1 * w0;
    const double h0 = v1 * wu2 - v2 * wu1, h1 = v2 * wu0 - v0 * wu2, h2 = v0 * wu1 - v1 * wu0;
    double g0 = 2.0 * (q3 * vw0 + e0 + h0) + c1 * (2.0 * q2 * X0 + 2.0 * q3 * X1 - 4.0 * q0);
    double g1 = 2.0 * (q3 * vw1 + e1 + h1) + c1 * (-2.0 * q3 * X0 + 2.0 * q2 * X1 - 4.0 * q1);
    double g2 = 2.0 * (q3 * vw2 + e2 + h2) + c1 * (2.0 * q0 * X0 + 2.0 * q1 * X1);
    double g3 = 2.0 * (w0 * uxv0 + w1 * uxv1 + w2 * uxv2) + c1 * (-2.0 * q1 * X0 + 2.0 * q0 * X1);
    // through the normalisation q = qt / |qt|
    const double qg = q0 * g0 + q1 * g1 + q2 * g2 + q3 * g3;
    g0 = (g0 - q0 * qg) * inz * sr;
    g1 = (g1 - q1 * qg) * inz * sr;
    g2 = (g2 - q2 * qg) * inz * sr;
    g3 = (g3 - q3 * qg) * inz * sr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double Qx = Q[4 * j], Qy = Q[4 * j + 1], Qz = Q[4 * j + 2], Qw = Q[4 * j + 3];
        J[(9 + 3 * j + 0) * JS] = b[j] * (g0 * Qw - g1 * Qz + g2 * Qy - g3 * Qx);
        J[(9 + 3 * j + 1) * JS] = b[j] * (g0 * Qz + g1 * Qw - g2 * Qx - g3 * Qy);
        J[(9 + 3 * j + 2) * JS] = b[j] * (-g0 * Qy + g1 * Qx + g2 * Qw - g3 * Qz);
        J[(21 + 3 * j + 0) * JS] = b[j] * sr * gt0;
        J[(21 + 3 * j + 1) * JS] = b[j] * sr * gt1;
        J[(21 + 3 * j + 2) * JS] = b[j] * sr * gt2;
    }
    return o;
}
