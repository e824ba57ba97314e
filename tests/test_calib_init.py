"""the OpenCV-free initialisation (include/ecb/calib_init.hpp) against golden results of the cv2 4.13
wheel (tests/golden/calib_init.npz, written by tests/golden/make_calib_init_golden.py) for the calls of
EventCalibIni::cvCalibration (event_camera_calib/src/EventCalibIni.cpp:143-327).  Host only, no GPU.

Tolerances: projectPoints / undistortPoints / Rodrigues are closed forms -> 1e-12; calibrateCamera and the iterative PnP are
optimisers that share one optimum -> 1e-6 relative on intrinsics, distortion, poses and the RMS; the non-iterative IPPE /
RANSAC poses agree at the noise level of the centres (documented deviation in the header)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "calib_init.npz"))
P = lambda a: a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def lib():
    so = os.path.join(ROOT, "tests", "_build", "libcalib_init_host.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tests", "helpers", "calib_init_host.cpp")])
    L = C.CDLL(so)
    L.ci_calibrate.restype = C.c_double
    L.ci_calibrate.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double] + [C.c_void_p] * 5
    L.ci_project.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4
    L.ci_undistort.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.ci_solve_pnp.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double] + [C.c_void_p] * 4
    L.ci_homography.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.ci_check_pose.argtypes = [C.c_double, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_double]
    return L


def _cam9(K, d):
    return np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2], d[0], d[1], d[2], d[3], d[4]], np.float64)


def _calibrate(lib, name):
    obj = np.ascontiguousarray(G["obj"], np.float64)
    img = np.ascontiguousarray(G[name + "_img"], np.float64)
    nv, n = img.shape[:2]
    fpp, ztd, far = (bool(x) for x in G[name + "_flags"])
    bits = (1 if fpp else 0) | (2 if ztd else 0) | (4 if far else 0)
    cam9, rv, tv = np.zeros(9), np.zeros((nv, 3)), np.zeros((nv, 3))
    tot, pv = np.zeros(1), np.zeros(nv, np.float32)
    W, H = (int(x) for x in G["size"])
    rms = lib.ci_calibrate(P(obj), n, P(img), nv, W, H, bits, 1.0, P(cam9), P(rv), P(tv), P(tot), P(pv))
    return rms, cam9, rv, tv, tot[0], pv


@pytest.mark.parametrize("name", ["a", "b", "c", "d"])
def test_calibrate_camera_equals_cv2(lib, name):
    rms, cam9, rv, tv, tot, pv = _calibrate(lib, name)
    ref9 = _cam9(G[name + "_K"], G[name + "_dist"])
    assert rms > 0
    np.testing.assert_allclose(rms, float(G[name + "_rms"]), rtol=1e-6)
    # focal lengths / principal point 1e-6 relative; distortion coefficients are weakly determined individually (their
    # optimum is flat along k1-k2-k3 combinations), so they are compared at 1e-5 absolute + 1e-5 relative
    np.testing.assert_allclose(cam9[:4], ref9[:4], rtol=1e-6)
    np.testing.assert_allclose(cam9[4:], ref9[4:], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(rv, G[name + "_rvecs"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(tv, G[name + "_tvecs"], rtol=1e-5, atol=1e-5)
    # computeReprojectionErrors (EventCalibIni.cpp:115-141) returns the same RMS up to the Point2f rounding of the projections
    assert abs(tot - rms) < 1e-4 and pv.shape[0] == rv.shape[0] and np.all(pv > 0)
    fpp, ztd, far = (bool(x) for x in G[name + "_flags"])
    if fpp:
        assert cam9[2] == (346 - 1) / 2 and cam9[3] == (260 - 1) / 2
    if ztd:
        assert cam9[6] == 0 and cam9[7] == 0
    if far:
        assert cam9[0] == cam9[1]


def test_project_undistort_rodrigues_closed_forms(lib):
    obj = np.ascontiguousarray(G["obj"], np.float64)
    cam9 = _cam9(G["proj_K"], G["proj_dist"])
    img = np.zeros((len(obj), 2))
    lib.ci_project(P(obj), len(obj), P(G["proj_rvec"].copy()), P(G["proj_tvec"].copy()), P(cam9), P(img))
    np.testing.assert_allclose(img, G["proj_img"], rtol=1e-12, atol=1e-10)
    xy = np.zeros_like(img)
    lib.ci_undistort(P(cam9), P(np.ascontiguousarray(G["proj_img"])), len(obj), P(xy))
    np.testing.assert_allclose(xy, G["undist_xy"], rtol=1e-9, atol=1e-11)
    for r, Rref in zip(G["rod_r"], G["rod_R"]):
        R, r2, q = np.zeros(9), np.zeros(3), np.zeros(4)
        lib.ci_rodrigues(P(r.copy()), P(R))
        np.testing.assert_allclose(R.reshape(3, 3), Rref, rtol=1e-12, atol=1e-14)
        lib.ci_rodrigues_inv(P(np.ascontiguousarray(Rref).ravel().copy()), P(r2))
        if np.linalg.norm(r) < np.pi:
            np.testing.assert_allclose(r2, r, rtol=1e-9, atol=1e-12)
        lib.ci_rot2quat(P(np.ascontiguousarray(Rref).ravel().copy()), P(q))
        from scipy.spatial.transform import Rotation as Rot
        qs = Rot.from_matrix(Rref).as_quat()
        assert min(np.abs(q - qs).max(), np.abs(q + qs).max()) < 1e-12


def test_homography_exact_on_noise_free_points(lib):
    rng = np.random.default_rng(0)
    H = np.array([[1.1, 0.2, 30.0], [-0.1, 0.9, 12.0], [1e-3, -2e-3, 1.0]])
    src = rng.uniform(0, 40, (20, 2))
    d = (H @ np.c_[src, np.ones(20)].T).T
    dst = np.ascontiguousarray(d[:, :2] / d[:, 2:])
    out = np.zeros(9)
    assert lib.ci_homography(P(src), P(dst), 20, P(out)) == 1
    np.testing.assert_allclose(out.reshape(3, 3), H, rtol=1e-9, atol=1e-11)
    assert lib.ci_homography(P(src), P(dst), 3, P(out)) == 0


def test_solve_pnp_planar(lib):
    obj = np.ascontiguousarray(G["obj"], np.float64)
    cam9 = _cam9(G["a_K"], G["a_dist"])
    for v in range(len(G["pnp_img"])):
        img = np.ascontiguousarray(G["pnp_img"][v], np.float64)
        r, t, inl, nin = np.zeros(3), np.zeros(3), np.zeros(36, np.int32), np.zeros(1, np.int32)
        assert lib.ci_solve_pnp(P(obj), 36, P(img), P(cam9), 4.0, P(r), P(t), P(inl), P(nin)) == 1
        mask = np.zeros(36, bool)
        mask[inl[:nin[0]]] = True
        # the same inlier set as cv2.solvePnPRansac(..., 4.0, ..., SOLVEPNP_IPPE)
        np.testing.assert_array_equal(mask, G["pnp_ransac_inl"][v])
        if mask.all():
            # without outliers the pose is the reprojection-error minimum = cv2's iterative solvePnP
            np.testing.assert_allclose(r, G["pnp_iter_r"][v], rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(t, G["pnp_iter_t"][v], rtol=1e-5, atol=1e-5)
        # IPPE (non-iterative) on the inliers: agreement at the noise level of the centres (0.15 px here)
        assert np.abs(r - G["pnp_ransac_r"][v]).max() < 2e-2
        assert np.abs(t - G["pnp_ransac_t"][v]).max() < 0.6
        # body pose conversion (EventCalibIni.cpp:260-275): Rwb = Rsw^T, twb = -Rsw^T tsw
        q, tw, R = np.zeros(4), np.zeros(3), np.zeros(9)
        lib.ci_body_pose(P(r), P(t), P(q), P(tw))
        lib.ci_rodrigues(P(r), P(R))
        from scipy.spatial.transform import Rotation as Rot
        np.testing.assert_allclose(Rot.from_quat(q).as_matrix(), R.reshape(3, 3).T, atol=1e-12)
        np.testing.assert_allclose(tw, -R.reshape(3, 3).T @ t, atol=1e-12)


def test_check_pose_thresholds(lib):
    """EventCalibIni::checkPose (:328-346): v_t < 2 * 0.25 cm / step and v_R < 2 * 5e-4 pi / step."""
    from scipy.spatial.transform import Rotation as Rot
    step = 5e-4
    q0 = Rot.from_rotvec([0.1, -0.2, 0.3]).as_quat()
    t0 = np.array([1.0, 2.0, -70.0])
    dt = 0.01
    lim_t, lim_r = 2 * 0.25 / step, 2 * 5e-4 * np.pi / step

    def chk(dq_angle, dtrans):
        q1 = (Rot.from_rotvec(np.array([0, 0, 1.0]) * dq_angle) * Rot.from_quat(q0)).as_quat()
        t1 = t0 + np.array([dtrans, 0, 0])
        return lib.ci_check_pose(1.0, P(q0.copy()), P(t0.copy()), 1.0 + dt, P(q1.copy()), P(t1.copy()), step)

    assert chk(0.0, 0.0) == 1
    assert chk(0.99 * lim_r * dt, 0.99 * lim_t * dt) == 1
    assert chk(1.01 * lim_r * dt, 0.0) == 0
    assert chk(0.0, 1.01 * lim_t * dt) == 0
