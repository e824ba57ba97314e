"""Generates tests/golden/reference_source.npz: inputs and outputs of the reference's OWN sources compiled in place
(oracle/_ref/libref_functor.so, see oracle/Makefile and oracle/ref_functor_capi.cpp) for every row of the hot path, so that
the restatement / the product can be checked against them where /root/reference does not exist:

  functor_*   CalibReprojectionError::operator() on Jet<37> (EventCalibSpline.hpp:168-229): value + 1x37 Jacobian
  basis_*     BsplineReal findSpan / dersBasisFuns (BsplineReal.hpp:107-145,208-231), knot vector (eq. 9.68), constructor fit
  frame_*     EventFrame constructor (EventFrame.cpp:10-36): per-polarity pixel lists in hash-set iteration order
  extract_*   CirclesEventFrame::extractFeatures (CirclesEventFrame.cpp:61-359): candidate centres (cv::Point2f), features
  fit_*       CirclesEventFrame::fitCircle (:361-415)
  rectify_*   rectifyFeatures (:417-638): rectified features, verdict; findCenter (CirclesEventFrame.hpp:50-65)
  spline_*    EventCalibSpline constructor (EventCalibSpline.cpp:14-251): segments, intrinsics, residual list, assembly
  so3_*       CalibReprojectionError_SO3::operator() on Jet<37> (EventCalibSpline.hpp:65-156) for the residual blocks of a small
              calibration problem, BsplineSO3::derBasisFuns (BsplineSO3.cpp:73-109), LocalParameterizationSO3 (BsplineSO3.hpp:190-221)
  gate_*      TrackingBase::process + EventCalibIni::track (EventCalibIni.cpp:18-97) on frame sequences arriving out of order
  pose_*      EventCalibIni::checkPose (:328-346)

Needs /root/reference (build container).  Run from the repo root:  python tests/golden/make_reference_source_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from eventcalib_b200 import synth, calib_problem  # noqa: E402

oracle.build(force=True)
assert oracle.have_ref_functor(), "oracle/_ref/libref_functor.so missing (needs /root/reference)"
out = {}
rng = np.random.default_rng(20240601)
cam = synth.Camera()

# ---- functor ----
F = dict(intr=[], rcp=[], tcp=[], obs=[], lm=[], b=[], r=[], jac=[], rd=[])
for _ in range(64):
    intr = cam.intrinsics() * (1 + rng.normal(0, 0.01, 9))
    q = rng.normal(0, 1, 4)
    q /= np.linalg.norm(q)
    rcp = q[None, :] + rng.normal(0, 0.05, (4, 4))
    tcp = np.array([20, 20, -75.0])[None, :] + rng.normal(0, 2, (4, 3))
    obs = np.array([rng.integers(0, 346), rng.integers(0, 260)], float)
    lm = np.array([rng.uniform(0, 40), rng.uniform(0, 44), 0.0])
    b = rng.uniform(0, 1, 4)
    b /= b.sum()
    r, jac, rd = oracle.ref_residual_jac(intr, rcp, tcp, obs, lm, 1.75, b)
    for k, v in zip(F, (intr, rcp, tcp, obs, lm, b, r, jac, rd)):
        F[k].append(v)
for k, v in F.items():
    out["functor_" + k] = np.array(v)

# ---- spline ----
board = synth.Board()
traj = synth.Trajectory(3, board, 78.0)
us = np.sort(rng.uniform(1.0, 1.5, 60))
us[0], us[-1] = 1.0, 1.5
q, tw = traj.quat_xyzw(us)
out["basis_us"], out["basis_q"], out["basis_tw"] = us, q, tw
for n_cp in (4, 9, 20):
    kn, cp3, _ = oracle.ref_spline_fit(us, tw, n_cp)
    _, cp4, _ = oracle.ref_spline_fit(us, q, n_cp)
    out[f"basis_knots_{n_cp}"], out[f"basis_cp3_{n_cp}"], out[f"basis_cp4_{n_cp}"] = kn, cp3, cp4
kn = out["basis_knots_20"]
uu = np.r_[rng.uniform(1.0, 1.5, 200), us, kn]
sp, NN = zip(*(oracle.ref_basis(kn, float(u)) for u in uu))
out["basis_u"], out["basis_span"], out["basis_N"] = uu, np.array(sp), np.array(NN)

# ---- event frame, extract, rectify ----
ev = synth.make_stream(40000, 346, 260, t0=5.0, duration=0.02, seed=1001, return_truth=True)
t, x, y, p = ev["t"], ev["x"], ev["y"], ev["p"]
out["ev_t"], out["ev_x"], out["ev_y"], out["ev_p"] = t, x.astype(np.int16), y.astype(np.int16), p.astype(np.uint8)
wins = synth.tiling_windows(5.0, 5.02, 1.5e-3)[::3]
out["windows"] = np.array(wins)
cen, sk = board.centres(), board.radius / np.sqrt(2)
for i, w in enumerate(wins):
    a, b_ = float(w[0]), float(w[1])
    P1, N1 = oracle.ref_event_frame(t, x, y, p, a, b_)
    out[f"frame_pos_{i}"], out[f"frame_neg_{i}"] = P1.astype(np.int16), N1.astype(np.int16)
    for fit in (0, 1):
        r1 = oracle.ref_extract(t, x, y, p, a, b_, 346, 260, fit)
        out[f"extract_found_{i}_{fit}"] = np.array(r1["found"])
        out[f"extract_cand_{i}_{fit}"] = r1["cand_f32"] if r1["cand_f32"] is not None else np.zeros((0, 2), np.float32)
        out[f"extract_reached_{i}_{fit}"] = np.array(r1["cand_f32"] is not None)
        out[f"extract_features_{i}_{fit}"] = r1["features"]
        R, tw_ = ev["trajectory"].pose(np.array([(a + b_) / 2]))
        img = np.zeros((36, 5, 2))
        for k in range(36):
            o5 = np.array([cen[k], cen[k] + [sk, sk, 0], cen[k] + [sk, -sk, 0], cen[k] + [-sk, -sk, 0], cen[k] + [-sk, sk, 0]])
            u, v = synth.project(ev["camera"], np.repeat(R, 5, 0), np.repeat(tw_, 5, 0), o5)
            img[k, :, 0], img[k, :, 1] = u, v
        if i % 2:
            img += rng.normal(0, 3.0, img.shape)
        img = img.astype(np.float32).astype(np.float64)
        fxy = np.c_[rng.integers(0, 346, 100), rng.integers(0, 260, 100)].astype(float)
        rc, o, fid = oracle.ref_rectify(t, x, y, p, a, b_, 346, 260, fit, img, fxy)
        out[f"rectify_img_{i}_{fit}"], out[f"rectify_rc_{i}_{fit}"], out[f"rectify_out_{i}_{fit}"] = img, np.array(rc), o
        out[f"rectify_fxy_{i}_{fit}"], out[f"rectify_fid_{i}_{fit}"] = fxy, fid
out["rthr"] = np.array(oracle.ref_extract(t, x, y, p, 5.0, 5.0015, 346, 260, 0)["rthr"])

# ---- fitCircle ----
fp, fn, fo = [], [], []
for _ in range(32):
    c, r = rng.uniform(50, 200, 2), rng.uniform(4, 12)
    th = rng.uniform(0, 2 * np.pi, 40)
    pts = np.rint(np.c_[c[0] + r * np.cos(th), c[1] + r * np.sin(th)] + rng.normal(0, 0.5, (40, 2)))
    fp.append(pts[:20]); fn.append(pts[20:]); fo.append(oracle.ref_fit_circle(pts[:20], pts[20:]))
out["fit_p"], out["fit_n"], out["fit_out"] = np.array(fp), np.array(fn), np.array(fo)

# ---- EventCalibSpline constructor ----
ev2 = synth.make_stream(75000, 346, 260, t0=5.0, duration=0.25, seed=11, return_truth=True)
step = 5e-4
pb = calib_problem.build_from_truth(ev2["camera"], ev2["trajectory"], board, 5.0, 5.25, step=step)
kf_t, circ = pb["kf_t"], pb["circles"].copy()
keep = ~(((kf_t > 5.10) & (kf_t < 5.13)) | ((kf_t >= 5.142) & (kf_t < 5.17)))   # two segments + a 3-frame island
kf_t, circ = kf_t[keep], circ[keep]
circ[rng.uniform(size=circ.shape[:2]) < 0.05, 2] = -1.0
kq, ktw = ev2["trajectory"].quat_xyzw(kf_t)
cam9 = np.array([cam.f * 1.01, cam.f * 0.99, cam.cx + 0.5, cam.cy - 0.5, -0.33, -0.02, 0, 0, 0.5])
r = oracle.ref_calib_spline(ev2["t"], ev2["x"], ev2["y"], ev2["p"], kf_t, kq, ktw, circ, board.centres(), cam9, 346, 260, step, board.radius)
out.update(spline_ev_t=ev2["t"], spline_ev_x=ev2["x"].astype(np.int16), spline_ev_y=ev2["y"].astype(np.int16),
           spline_ev_p=ev2["p"].astype(np.uint8), spline_kf_t=kf_t, spline_kf_q=kq, spline_kf_tw=ktw, spline_circ=circ, spline_cam9=cam9,
           spline_n_cp=r["n_cp"], spline_ranges=r["ranges"], spline_intrinsics=r["intrinsics"], spline_frames_left=np.array(r["frames_left"]),
           spline_kf_pose=r["kf_pose"], spline_span=r["span"].astype(np.int16), spline_spline=r["spline"].astype(np.int8),
           spline_first_cp=r["first_cp"].astype(np.int16), spline_lm_idx=np.array([int(np.argmin(((board.centres() - l) ** 2).sum(1))) for l in r["lm"]], np.int8),
           spline_obs=r["obs"].astype(np.int16), spline_basis_sample=r["basis"][::101],
           spline_assembly=np.array([r["param_blocks"], r["quaternion_blocks"], r["linear_solver"], r["solve_calls"]]),
           spline_huber_tol=np.array([r["huber"], r["gradient_tolerance"], r["function_tolerance"]]))
for s in range(r["n_splines"]):
    out[f"spline_knots_{s}"], out[f"spline_rot_{s}"], out[f"spline_trans_{s}"] = r["knots"][s], r["rot_cp"][s], r["trans_cp"][s]
# ---- a11: the SO(3) variant — CalibReprojectionError_SO3 on Jet<37>, BsplineSO3::derBasisFuns, LocalParameterizationSO3 ----
ev3 = synth.make_stream(1500, 346, 260, t0=5.0, duration=0.1, seed=21, return_truth=True)
pb3 = calib_problem.build(ev3, seed=3)
P3 = oracle.CostProblem([pb3["n_cp"]], [pb3["knots"]], pb3["radius"], pb3["huber"], so3=True)
oe3, oc3 = P3.associate(ev3["t"], ev3["x"], ev3["y"], pb3["kf_t"], pb3["circles"], pb3["landmarks"], pb3["step"])
S = dict(span=[], beta=[], N=[], r=[], jac=[], rd=[])
for e_, c_ in zip(oe3, oc3):
    u = float(ev3["t"][e_])
    sp_r, N4 = oracle.ref_basis(pb3["knots"], u)
    sp_s, beta = oracle.ref_so3_basis(pb3["knots"], u)
    assert sp_r == sp_s
    r_, jac_, rd_ = oracle.ref_residual_jac_so3(pb3["intrinsics"], pb3["rot_cp"][sp_s - 3:sp_s + 1], pb3["trans_cp"][sp_s - 3:sp_s + 1],
                                                np.array([ev3["x"][e_], ev3["y"][e_]]), pb3["landmarks"][c_], pb3["radius"], beta, N4)
    for k, v in zip(S, (sp_s, beta, N4, r_, jac_, rd_)):
        S[k].append(v)
out.update(so3_ev_t=ev3["t"], so3_ev_x=ev3["x"].astype(np.int16), so3_ev_y=ev3["y"].astype(np.int16), so3_ev_p=ev3["p"].astype(np.uint8),
           so3_kf_t=pb3["kf_t"], so3_circles=pb3["circles"], so3_landmarks=pb3["landmarks"], so3_knots=pb3["knots"],
           so3_rot_cp=pb3["rot_cp"], so3_trans_cp=pb3["trans_cp"], so3_intrinsics=pb3["intrinsics"], so3_step=np.array(pb3["step"]),
           so3_event=oe3, so3_circle=oc3.astype(np.int8))
for k, v in S.items():
    out["so3_" + k] = np.array(v)
out["so3_plus_jac"] = np.array([oracle.ref_so3_plus_jacobian(q_) for q_ in pb3["rot_cp"]])
px = rng.normal(size=(40, 4))
px /= np.linalg.norm(px, axis=1, keepdims=True)
pd = rng.normal(size=(40, 3)) * 10.0 ** rng.uniform(-12, 0, (40, 1))
out["so3_plus_x"], out["so3_plus_d"] = px, pd
out["so3_plus_out"] = np.array([oracle.ref_so3_plus(a_, b_) for a_, b_ in zip(px, pd)])

# ---- tracking gate and checkPose ----
from scipy.spatial.transform import Rotation as Rot  # noqa: E402
for trial, amp in enumerate((1.0, 3.0)):
    trj = synth.Trajectory(5 + trial, board, 78.0, rot_amp=(0.1, 0.1, amp))
    ini = oracle.RefIni(346, 260, 5e-4)
    ts = np.sort(rng.uniform(5.0, 5.6, 80))
    rng.shuffle(ts[20:])
    xs, acc = [], []
    for tt in ts:
        R, tw_ = trj.pose(np.array([tt]))
        u, v = synth.project(cam, np.repeat(R, 36, 0), np.repeat(tw_, 36, 0), board.centres())
        xy = np.ascontiguousarray(np.c_[u, v] + rng.normal(0, 0.2, (36, 2)))
        xs.append(xy)
        acc.append(ini.gate(float(tt), xy))
    out[f"gate_ts_{trial}"], out[f"gate_xy_{trial}"], out[f"gate_accept_{trial}"] = ts, np.array(xs), np.array(acc, np.int8)
pc = []
for it in range(400):
    q0, t0 = Rot.random(random_state=it).as_quat(), rng.normal(0, 30, 3)
    dt = rng.uniform(1e-3, 2e-2)
    ang = rng.uniform(0, 2.2) * 2 * 5e-4 * np.pi / 5e-4 * dt
    dtr = rng.uniform(0, 2.2) * 2 * 0.25 / 5e-4 * dt
    q1 = (Rot.from_rotvec(Rot.random(random_state=it + 7).apply([0, 0, 1]) * ang) * Rot.from_quat(q0)).as_quat()
    t1 = t0 + Rot.random(random_state=it + 9).apply([1, 0, 0]) * dtr
    pc.append(np.r_[q0, t0, dt, q1, t1, oracle.ref_check_pose(1.0, q0, t0, 1.0 + dt, q1, t1, 5e-4)])
out["pose_cases"] = np.array(pc)   # q0(4) t0(3) dt q1(4) t1(3) verdict

path = os.path.join(ROOT, "tests", "golden", "reference_source.npz")
np.savez_compressed(path, **out)
print("written", len(out), "arrays,", os.path.getsize(path) // 1024, "KiB; residuals", r["n_residuals"])
