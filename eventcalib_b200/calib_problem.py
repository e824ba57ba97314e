"""Synthetic dynamic-calibration problems (BASELINE.json config C4) for tests and bench.py.

From a synthetic stream with ground truth (synth.make_stream(return_truth=True)) this builds what
EventCalibSpline's constructor has when optimize() starts (src/EventCalibSpline.cpp:14-113): key frames with
their 36 circle centres / radii, one spline segment with knots every 50 x MotionTimeStep, control points fitted
to the (noisy) key-frame poses, and perturbed initial intrinsics.
"""
import numpy as np

from . import spline, synth


def build(ev, **kw):
    """Problem for a stream generated with return_truth=True."""
    return build_from_truth(ev["camera"], ev["trajectory"], ev["board"], float(ev["t"][0]), float(ev["t"][-1]), **kw)


def build_from_truth(cam, traj, board, t0, t1, step=5e-4, kf_every=None, seed=0, intr_noise=0.02, pose_noise=(0.002, 0.05),
                     max_cp=None):
    kf_every = kf_every or 8 * step          # len 3 step + frameGap 5 step (eventCameraCalib.cpp:168-169)
    kf_t = np.arange(t0 + 2 * step, t1 - 2 * step, kf_every)
    K = len(kf_t)
    R, tw = traj.pose(kf_t)
    centres = board.centres()
    n_c = len(centres)
    circ = np.zeros((K, n_c, 3))
    th = np.linspace(0, 2 * np.pi, 16, endpoint=False)
    for q in range(n_c):
        C = np.repeat(centres[q][None, :], K, 0)
        cu, cv = synth.project(cam, R, tw, C)
        rad = np.zeros(K)
        for a in th:
            rim = C + board.radius * np.array([np.cos(a), np.sin(a), 0.0])
            u, v = synth.project(cam, R, tw, rim)
            rad += np.hypot(u - cu, v - cv)
        circ[:, q, 0], circ[:, q, 1], circ[:, q, 2] = cu, cv, rad / len(th)
    # spline segment: time bounds extended by 3 steps, cpNum = floor(T / (50 step)) clamped (EventCalibSpline.cpp:63-91)
    us = kf_t.copy()
    us[0] -= 3 * step
    us[-1] += 3 * step
    n_cp = int(np.floor((us[-1] - us[0]) / (50 * step)))
    if n_cp > len(us):
        n_cp = len(us) - 1
    n_cp = max(n_cp, 4)
    if max_cp:
        n_cp = min(n_cp, max_cp)
    kn = spline.knot_vector(us, n_cp)
    rng = np.random.default_rng(seed)
    q, tw = traj.quat_xyzw(kf_t)
    qn = q + rng.normal(0, pose_noise[0], q.shape)
    qn /= np.linalg.norm(qn, axis=1, keepdims=True)
    twn = tw + rng.normal(0, pose_noise[1], tw.shape)
    rot_cp = spline.fit_control_points(kn, us, qn, n_cp)
    rot_cp /= np.linalg.norm(rot_cp, axis=1, keepdims=True)   # optimize() treats them as quaternions (:129-133)
    trans_cp = spline.fit_control_points(kn, us, twn, n_cp)
    intr = cam.intrinsics() * (1 + intr_noise * np.array([1, 1, 0.5, -0.5, 1, -1, 1, -1, 1.0]))
    return dict(kf_t=kf_t, circles=circ, landmarks=centres.copy(), n_cp=n_cp, knots=kn, rot_cp=rot_cp, trans_cp=trans_cp,
                intrinsics=intr, step=step, radius=board.radius, huber=0.2 * board.radius, truth_intrinsics=cam.intrinsics())
