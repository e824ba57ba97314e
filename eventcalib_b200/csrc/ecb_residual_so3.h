// SO(3) variant of the event-to-circle residual: CalibReprojectionError_SO3::operator()
// (event_camera_calib/include/opengv2/event_camera_calib/EventCalibSpline.hpp:65-156) as Ceres evaluates it with
// LocalParameterizationSO3 (core/spline/include/opengv2/spline/BsplineSO3.hpp:190-221) and HuberLoss.
// Same camera / ray-plane / distance part as ecb_residual.h (the functor bodies are identical from :107 on); the rotation
// comes from the cumulative SO(3) B-spline (ecb_so3.h) and its tangent derivative from 3-partial dual numbers.
// b: the 4 basis values N_{i-3..i}(u) — the translation uses them directly, the rotation their cumulative sums.
#pragma once
#include "ecb_residual.h"
#include "ecb_so3.h"

template <bool WANT_JAC, int JS = 1>
ECB_HD EcbResidualOut ecb_residual_so3(const double *intr, const double *Q, const double *T, const double *b, double ou,
                                       double ov, double lx, double ly, double lz, double radius, double huber, double *J) {
    double beta[3], qv[4];
    ecb_so3::cumulative_basis(b, beta);
    ecb_so3::rotation_value(Q, beta, qv);
    const double q0 = qv[0], q1 = qv[1], q2 = qv[2], q3 = qv[3];  // x y z w
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
    g0 *= sr; g1 *= sr; g2 *= sr; g3 *= sr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        double dq[4][3];
        ecb_so3::rotation_tangent(Q, beta, j, dq);
#pragma unroll
        for (int k = 0; k < 3; ++k)
            J[(9 + 3 * j + k) * JS] = g0 * dq[0][k] + g1 * dq[1][k] + g2 * dq[2][k] + g3 * dq[3][k];
        J[(21 + 3 * j + 0) * JS] = b[j] * sr * gt0;
        J[(21 + 3 * j + 1) * JS] = b[j] * sr * gt1;
        J[(21 + 3 * j + 2) * JS] = b[j] * sr * gt2;
    }
    return o;
}
