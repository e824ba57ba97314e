"""Host-only parts of the C++ façade (include/ecb/event_calib.hpp), no GPU: EventStream::txt2bin (EventStream.cpp:25-67),
EventCalibSpline::evaluate for both rotation models (EventCalibSpline.hpp:305-330) and the TUM trajectory writer
(SystemBase.cpp:122-150)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    import eventcalib_b200.build as b
    b.build()
    so = os.path.join(ROOT, "tests", "_build", "libfacade_host.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tests", "helpers", "facade_host.cpp"),
                           "-L" + os.path.join(ROOT, "eventcalib_b200"), "-lecb",
                           "-Wl,-rpath," + os.path.join(ROOT, "eventcalib_b200")])
    return C.CDLL(so)


def test_txt2bin_round_trip(tmp_path):
    from eventcalib_b200 import synth
    lib = _lib()
    lib.fh_txt2bin.restype = C.c_longlong
    lib.fh_txt2bin.argtypes = [C.c_char_p, C.c_double]
    rng = np.random.default_rng(0)
    n = 1000
    stamps = np.sort(rng.integers(1_000_000, 9_000_000, n)) + 1_600_000_000_000_000
    x, y, p = rng.integers(0, 346, n), rng.integers(0, 260, n), rng.integers(0, 2, n)
    txt = tmp_path / "ev.txt"
    txt.write_text("\n".join("%d %d %d %d" % v for v in zip(stamps, x, y, p)) + "\n")
    cnt = lib.fh_txt2bin(str(txt).encode(), 1e-6)
    ev = synth.read_bin(str(tmp_path / "ev.bin"))
    # the reference's reader loop re-emits the last line once when the stream ends with a newline (is.good() idiom)
    assert cnt in (n, n + 1) and len(ev["t"]) == cnt
    np.testing.assert_array_equal(ev["t"][:n], (stamps - stamps[0]) * 1e-6)
    np.testing.assert_array_equal(ev["x"][:n], x)
    np.testing.assert_array_equal(ev["y"][:n], y)
    np.testing.assert_array_equal(ev["p"][:n], p)


def test_spline_evaluate_and_tum(tmp_path):
    from scipy.spatial.transform import Rotation as Rot
    from eventcalib_b200 import synth, spline
    lib = _lib()
    board = synth.Board()
    traj = synth.Trajectory(3, board, 78.0)
    us = np.linspace(1.0, 1.5, 60)
    n_cp = 9
    kn = spline.knot_vector(us, n_cp)
    q, tw = traj.quat_xyzw(us)
    rot = np.ascontiguousarray(spline.fit_control_points(kn, us, q, n_cp))
    rot /= np.linalg.norm(rot, axis=1, keepdims=True)
    trans = np.ascontiguousarray(spline.fit_control_points(kn, us, tw, n_cp))
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.fh_eval.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
    for t in (1.0, 1.1234, 1.3, 1.5):
        sp = spline.find_span(kn, t)
        N = spline.basis(kn, sp, t)
        want_t = N @ trans[sp - 3: sp + 1]
        qo, to = np.zeros(4), np.zeros(3)
        assert lib.fh_eval(P(kn), n_cp, P(rot), P(trans), 0, t, P(qo), P(to)) == 1
        wq = N @ rot[sp - 3: sp + 1]
        np.testing.assert_allclose(qo, wq / np.linalg.norm(wq), rtol=0, atol=1e-15)
        np.testing.assert_allclose(to, want_t, rtol=0, atol=1e-13)
        assert lib.fh_eval(P(kn), n_cp, P(rot), P(trans), 1, t, P(qo), P(to)) == 1
        beta = [N[1] + N[2] + N[3], N[2] + N[3], N[3]]
        R = [Rot.from_quat(v) for v in rot[sp - 3: sp + 1]]
        Rw = R[0]
        for j in range(1, 4):
            Rw = Rw * Rot.from_rotvec(beta[j - 1] * (R[j - 1].inv() * R[j]).as_rotvec())
        wq = Rw.as_quat()
        assert min(np.abs(qo - wq).max(), np.abs(qo + wq).max()) < 1e-13
    assert lib.fh_eval(P(kn), n_cp, P(rot), P(trans), 0, 2.0, P(qo), P(to)) == 0   # outside every segment
    ts = np.array([0.5, 1.05, 1.25, 1.45, 3.0])
    out = tmp_path / "TrajectoryByEvent.txt"
    lib.fh_tum.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_char_p, C.c_void_p, C.c_int]
    assert lib.fh_tum(P(kn), n_cp, P(rot), P(trans), 0, str(out).encode(), P(ts), len(ts)) == 3
    lines = out.read_text().splitlines()
    assert len(lines) == 3
    for line, t in zip(lines, ts[1:4]):
        f = line.split()
        assert len(f) == 8 and all(len(v.split(".")[1]) == 10 for v in f)     # fixed, precision 10
        assert abs(float(f[0]) - t) < 1e-9
        v = np.array(list(map(float, f[1:])))
        assert abs(np.linalg.norm(v[3:]) - 1) < 1e-8


def test_tracking_gate_matches_numpy_restatement():
    """TrackingGate (EventCalibIni::track) against tests/gate_py.py on a sequence with slow and fast rotations"""
    import gate_py
    from eventcalib_b200 import synth
    lib = _lib()
    lib.fh_gate_new.restype = C.c_void_p
    lib.fh_gate_new.argtypes = [C.c_int, C.c_int, C.c_double]
    lib.fh_gate_process.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_int]
    lib.fh_gate_free.argtypes = [C.c_void_p]
    board, cam = synth.Board(), synth.Camera()
    c = board.centres()
    ctr = c.mean(0)
    rng = np.random.default_rng(1)
    g = C.c_void_p(lib.fh_gate_new(9, 4, 5e-4))
    ref = gate_py.Gate(9, 4, 5e-4)
    th, t = 0.3, 5.0
    verdicts = []
    order = []
    for k in range(60):
        t += 4e-3
        th += (0.002 if k % 5 else 0.06) * rng.uniform(0.5, 1.5)      # every fifth frame jumps: > pi rad/s
        order.append((t, th))
    order[10], order[11] = order[11], order[10]                       # out-of-order arrival: lower_bound finds a LATER key frame
    for t_, th_ in order:
        cz, sz = np.cos(th_), np.sin(th_)
        R = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]) @ np.array([[1, 0, 0], [0, np.cos(0.2), -np.sin(0.2)], [0, np.sin(0.2), np.cos(0.2)]])
        tw = ctr + R @ np.array([0, 0, -95.0])
        u, v = synth.project(cam, np.repeat(R[None], 36, 0), np.repeat(tw[None], 36, 0), c)
        f = np.ascontiguousarray(np.stack([u, v], 1) + rng.normal(0, 0.1, (36, 2)))
        a = lib.fh_gate_process(g, t_, f.ctypes.data_as(C.c_void_p), 36)
        b = ref.process(t_, f)
        verdicts.append((a, int(b)))
    lib.fh_gate_free(g)
    assert all(a == b for a, b in verdicts)
    assert 0 < sum(a for a, _ in verdicts) < len(verdicts)            # both accepted and rejected frames occurred


def test_inverse_radial_distortion_known_answer():
    """PinholeCamera::inverseRadialDistortion on the values of the reference's unit_test_inverseDistortion.cpp:10-16
    (known answer: SURVEY.md section 4) and against the oracle restatement, bit for bit"""
    import oracle
    oracle.build()
    lib = _lib()
    k = np.array([-0.34991902, -0.014698517, 0.59684463, 0.0])
    b = np.zeros(5)
    lib.fh_inverse_radial(k.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p))
    want = [0.34991902, 0.38202847867328127, -0.041555343865844696, -1.1638270394205459, -4.138165444396021]
    np.testing.assert_allclose(b, want, rtol=1e-15)
    ob = np.zeros(5)
    oracle.port().orc_inverse_radial(k.ctypes.data_as(C.c_void_p), ob.ctypes.data_as(C.c_void_p))
    assert np.array_equal(b, ob)


def _fit_fixed_ends(us, data, n_cp):
    """numpy restatement of BsplineReal::approximation + optimization (BsplineReal.hpp:87-100,329-449, no derivative
    samples): eq. 9.68 knots, end control points = end samples, interior ones by least squares over the interior samples"""
    from eventcalib_b200 import spline
    kn = spline.knot_vector(us, n_cp)
    N = np.zeros((len(us), n_cp))
    for k, u in enumerate(us):
        sp = spline.find_span(kn, u)
        N[k, sp - 3: sp + 1] = spline.basis(kn, sp, u)
    cp = np.zeros((n_cp, data.shape[1]))
    cp[0], cp[-1] = data[0], data[-1]
    R = data[1:-1] - np.outer(N[1:-1, 0], data[0]) - np.outer(N[1:-1, -1], data[-1])
    Nc = N[1:-1, 1:-1]
    cp[1:-1] = np.linalg.solve(Nc.T @ Nc, Nc.T @ R)
    return kn, cp


def test_spline_setup_from_keyframes():
    """EventCalibSpline constructor set-up (EventCalibSpline.cpp:25-91): frame count check, gap segmentation (reduceMap
    :319-348), extended bounds, cpNum rule and the BsplineReal fits"""
    from eventcalib_b200 import synth
    lib = _lib()
    board = synth.Board()
    traj = synth.Trajectory(5, board, 78.0)
    step = 5e-4
    ts = np.concatenate([np.arange(1.0, 1.4, 4e-3), np.arange(1.5, 1.506, 4e-3), np.arange(1.6, 2.3, 4e-3)])  # middle: 2 frames only
    q, tw = traj.quat_xyzw(ts)
    P = lambda a: np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)
    lib.fh_segments.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_void_p] + [C.c_void_p] * 3
    ncp = C.c_int(0)
    kn, rot, tr = np.zeros(2000), np.zeros(8000), np.zeros(6000)
    assert lib.fh_segments(P(ts[:10]), P(q[:10]), P(tw[:10]), 10, step, 0, C.byref(ncp), P(kn), P(rot), P(tr)) == -1   # <= 10 frames
    segs = [np.arange(0, 100), np.arange(102, len(ts))]
    for which, idx in enumerate(segs):
        assert lib.fh_segments(P(ts), P(q), P(tw), len(ts), step, which, C.byref(ncp), P(kn), P(rot), P(tr)) == 2
        us = ts[idx].copy()
        us[0] -= 3 * step
        us[-1] += 3 * step
        want_cp = int(np.floor((us[-1] - us[0]) / (50 * step)))
        want_cp = max(4, len(us) - 1 if want_cp > len(us) else want_cp)
        assert ncp.value == want_cp
        k_ref, t_ref = _fit_fixed_ends(us, tw[idx], want_cp)
        _, q_ref = _fit_fixed_ends(us, q[idx], want_cp)
        np.testing.assert_array_equal(kn[:want_cp + 4], k_ref)
        np.testing.assert_allclose(tr[:3 * want_cp].reshape(-1, 3), t_ref, rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(rot[:4 * want_cp].reshape(-1, 4), q_ref, rtol=1e-9, atol=1e-9)
