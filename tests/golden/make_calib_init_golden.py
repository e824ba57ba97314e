"""Generates tests/golden/calib_init.npz: inputs and cv2 4.13 results for the OpenCV calls of the reference's initialisation
stage (EventCalibIni::cvCalibration, modules/camera_calibration/event_camera_calib/src/EventCalibIni.cpp:143-327):
calibrateCamera with the flags CalibrationSetting::validate builds from example.yaml (| CALIB_USE_LU), projectPoints,
undistortPoints, solvePnP (ITERATIVE and IPPE), solvePnPRansac(..., 50, 4.0, 0.99, inliers, SOLVEPNP_IPPE), Rodrigues.
Image / object points are float32 like the reference's cv::Point2f / cv::Point3f.  Needs cv2 (present in the build
container); run from the repo root:

    python tests/golden/make_calib_init_golden.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from eventcalib_b200 import synth  # noqa: E402

board = synth.Board()
obj = board.centres().astype(np.float32)  # cv::Point3f
W, H = 346, 260
K_true = np.array([[359.67525, 0, (W - 1) / 2], [0, 359.67525, (H - 1) / 2], [0, 0, 1]])
d_true = np.array([-0.34991902, -0.014698517, 0, 0, 0.59684463])
out = {"obj": obj, "size": np.array([W, H])}


def views(seed, n, tilt, sigma):
    rng = np.random.default_rng(seed)
    imgs, rv, tv = [], [], []
    mid = obj.mean(axis=0)
    while len(imgs) < n:
        r = rng.uniform(-tilt, tilt, 3) * np.array([1, 1, 2.0])
        R, _ = cv2.Rodrigues(r)
        t = np.array([rng.uniform(-6, 6), rng.uniform(-4, 4), rng.uniform(55, 85)]) - R @ mid
        p, _ = cv2.projectPoints(obj.astype(np.float64), r, t, K_true, d_true)
        p = p.reshape(-1, 2) + rng.normal(0, sigma, (len(obj), 2))
        if p.min() < 2 or p[:, 0].max() > W - 3 or p[:, 1].max() > H - 3:
            continue
        imgs.append(p.astype(np.float32))
        rv.append(r)
        tv.append(t)
    return np.array(imgs), np.array(rv), np.array(tv)


FLAGS_EXAMPLE = (cv2.CALIB_FIX_PRINCIPAL_POINT | cv2.CALIB_ZERO_TANGENT_DIST | cv2.CALIB_FIX_ASPECT_RATIO | cv2.CALIB_FIX_K4 |
                 cv2.CALIB_FIX_K5 | cv2.CALIB_FIX_K6)  # example.yaml:44-60 through parameters.hpp:49-60
FLAGS_FREE = cv2.CALIB_FIX_K4 | cv2.CALIB_FIX_K5 | cv2.CALIB_FIX_K6
cases = [("a", 11, 40, 0.45, 0.15, FLAGS_EXAMPLE), ("b", 12, 25, 0.30, 0.30, FLAGS_EXAMPLE), ("c", 13, 30, 0.50, 0.10, FLAGS_FREE),
         ("d", 14, 12, 0.12, 0.20, FLAGS_EXAMPLE)]
for name, seed, n, tilt, sigma, flags in cases:
    img, rv, tv = views(seed, n, tilt, sigma)
    ip = [v.reshape(-1, 1, 2) for v in img]
    op = [obj.reshape(-1, 1, 3)] * n
    for tag, crit in (("", None), ("_conv", (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 2000, 1e-16))):
        K0 = np.eye(3)  # aspectRatio = 1 in (0,0) (EventCalibIni.cpp:147-149)
        d0 = np.zeros((8, 1))
        kw = {} if crit is None else {"criteria": crit}
        rms, K, d, rvecs, tvecs = cv2.calibrateCamera(op, ip, (W, H), K0, d0, flags=flags | cv2.CALIB_USE_LU, **kw)
        out[f"{name}_rms{tag}"] = rms
        out[f"{name}_K{tag}"] = K
        out[f"{name}_dist{tag}"] = d.ravel()
        out[f"{name}_rvecs{tag}"] = np.array(rvecs).reshape(n, 3)
        out[f"{name}_tvecs{tag}"] = np.array(tvecs).reshape(n, 3)
    out[f"{name}_img"] = img
    out[f"{name}_flags"] = np.array([bool(flags & cv2.CALIB_FIX_PRINCIPAL_POINT), bool(flags & cv2.CALIB_ZERO_TANGENT_DIST),
                                     bool(flags & cv2.CALIB_FIX_ASPECT_RATIO)])
    print(name, "rms", out[f"{name}_rms"], out[f"{name}_rms_conv"], "f", out[f"{name}_K"][0, 0], out[f"{name}_K_conv"][0, 0],
          "dist", out[f"{name}_dist"][:5])

# projectPoints / undistortPoints / Rodrigues known answers
rng = np.random.default_rng(5)
d_full = np.array([-0.31, 0.12, 0.002, -0.001, 0.05])
r = np.array([0.21, -0.33, 1.4])
t = np.array([-12.0, -20.0, 70.0])
p, _ = cv2.projectPoints(obj.astype(np.float64), r, t, K_true, d_full)
out["proj_rvec"], out["proj_tvec"], out["proj_dist"], out["proj_K"], out["proj_img"] = r, t, d_full, K_true, p.reshape(-1, 2)
u = cv2.undistortPointsIter(p.reshape(-1, 1, 2), K_true, d_full, None, None, criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 100, 1e-14))
out["undist_xy"] = u.reshape(-1, 2)
rs = rng.normal(0, 1.0, (20, 3))
rs[0] = 0
rs[1] *= 1e-9
out["rod_r"] = rs
out["rod_R"] = np.array([cv2.Rodrigues(x)[0] for x in rs])

# PnP on the views of case a with its calibrated camera; some views get gross outliers
img, _, _ = views(11, 40, 0.45, 0.15)
K, d = out["a_K"], out["a_dist"][:5]
pnp_img, it_r, it_t, ip_r, ip_t, rs_r, rs_t, rs_in = [], [], [], [], [], [], [], []
for v in range(12):
    q = img[v].copy()
    if v % 3 == 2:
        bad = rng.choice(36, 4, replace=False)
        q[bad] += rng.uniform(8, 20, (4, 2)).astype(np.float32)
    ok, r1, t1 = cv2.solvePnP(obj, q, K, d, flags=cv2.SOLVEPNP_ITERATIVE)
    ok, r2, t2 = cv2.solvePnP(obj, q, K, d, flags=cv2.SOLVEPNP_IPPE)
    cv2.setRNGSeed(0)
    ok, r3, t3, inl = cv2.solvePnPRansac(obj, q, K, d, None, None, False, 50, 4.0, 0.99, None, cv2.SOLVEPNP_IPPE)
    m = np.zeros(36, bool)
    m[inl.ravel()] = True
    pnp_img.append(q); it_r.append(r1.ravel()); it_t.append(t1.ravel()); ip_r.append(r2.ravel()); ip_t.append(t2.ravel())
    rs_r.append(r3.ravel()); rs_t.append(t3.ravel()); rs_in.append(m)
out.update(pnp_img=np.array(pnp_img), pnp_iter_r=np.array(it_r), pnp_iter_t=np.array(it_t), pnp_ippe_r=np.array(ip_r),
           pnp_ippe_t=np.array(ip_t), pnp_ransac_r=np.array(rs_r), pnp_ransac_t=np.array(rs_t), pnp_ransac_inl=np.array(rs_in))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "calib_init.npz"), **out)
print("written", len(out), "arrays")
