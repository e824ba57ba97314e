// C shim over include/ecb/calib_init.hpp for the CPU test suite (ctypes).  Host only, no libecb.
#include "../../include/ecb/calib_init.hpp"
using namespace ecb;
extern "C" {
// flag bits: 1 fixPrincipalPoint, 2 zeroTangentDist, 4 fixAspectRatio, 8 fixK1, 16 fixK2, 32 fixK3
// obj: n x 3, img: nv x n x 2; cam9 out: fx fy cx cy k1 k2 p1 p2 k3; rvecs / tvecs: nv x 3; per_view: nv floats
double ci_calibrate(const double *obj, int n, const double *img, int nv, int width, int height, int flag_bits, double aspect,
                    double *cam9, double *rvecs, double *tvecs, double *total_err, float *per_view) {
    CalibFlags fl;
    fl.fixPrincipalPoint = flag_bits & 1, fl.zeroTangentDist = flag_bits & 2, fl.fixAspectRatio = flag_bits & 4;
    fl.fixK1 = flag_bits & 8, fl.fixK2 = flag_bits & 16, fl.fixK3 = flag_bits & 32;
    fl.aspectRatio = aspect;
    std::vector<double> o(obj, obj + 3 * n);
    std::vector<std::vector<double>> ip((size_t) nv);
    for (int v = 0; v < nv; ++v) ip[(size_t) v].assign(img + (size_t) v * 2 * n, img + (size_t) (v + 1) * 2 * n);
    CameraModel cam;
    std::vector<std::array<double, 3>> rv, tv;
    const double rms = calibrateCamera(o, ip, width, height, fl, cam, rv, tv);
    if (rms < 0) return rms;
    const double c9[9] = {cam.fx, cam.fy, cam.cx, cam.cy, cam.dist[0], cam.dist[1], cam.dist[2], cam.dist[3], cam.dist[4]};
    for (int k = 0; k < 9; ++k) cam9[k] = c9[k];
    for (int v = 0; v < nv; ++v)
        for (int k = 0; k < 3; ++k) rvecs[3 * v + k] = rv[(size_t) v][(size_t) k], tvecs[3 * v + k] = tv[(size_t) v][(size_t) k];
    std::vector<float> pv;
    *total_err = computeReprojectionErrors(o, ip, rv, tv, cam, pv);
    for (int v = 0; v < nv; ++v) per_view[v] = pv[(size_t) v];
    return rms;
}
static CameraModel cam_of(const double *cam9) {
    CameraModel cam;
    cam.fx = cam9[0], cam.fy = cam9[1], cam.cx = cam9[2], cam.cy = cam9[3];
    for (int k = 0; k < 5; ++k) cam.dist[k] = cam9[4 + k];
    return cam;
}
void ci_project(const double *obj, int n, const double *rvec, const double *tvec, const double *cam9, double *img) {
    projectPoints(obj, n, rvec, tvec, cam_of(cam9), img);
}
void ci_undistort(const double *cam9, const double *uv, int n, double *xy) {
    const CameraModel cam = cam_of(cam9);
    for (int i = 0; i < n; ++i) undistortPoint(cam, uv + 2 * i, xy + 2 * i);
}
int ci_solve_pnp(const double *obj, int n, const double *img, const double *cam9, double thr, double *rvec, double *tvec,
                 int *inliers, int *n_inliers) {
    std::vector<int> in;
    const bool ok = solvePnPPlanar(std::vector<double>(obj, obj + 3 * n), std::vector<double>(img, img + 2 * n), cam_of(cam9), thr,
                                   rvec, tvec, in);
    *n_inliers = (int) in.size();
    for (size_t i = 0; i < in.size(); ++i) inliers[i] = in[i];
    return ok ? 1 : 0;
}
int ci_homography(const double *src, const double *dst, int n, double *H) { return findHomography(src, dst, n, H) ? 1 : 0; }
void ci_rodrigues(const double *r, double *R) { rodrigues<double>(r, R); }
void ci_rodrigues_inv(const double *R, double *r) { rodriguesInverse(R, r); }
void ci_rot2quat(const double *R, double *q) { rotationToQuaternion(R, q); }
void ci_body_pose(const double *rvec, const double *tvec, double *q, double *t) { bodyPoseFromPnP(rvec, tvec, q, t); }
int ci_check_pose(double rs, const double *rq, const double *rt, double cs, const double *cq, const double *ct, double step) {
    return checkPose(rs, rq, rt, cs, cq, ct, step) ? 1 : 0;
}
}
