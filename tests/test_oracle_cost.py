"""CPU checks of the cost-evaluation oracle and of the product's closed-form Jacobian (host build of ecb_residual.h)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_inverse_distortion_known_answer(oracle_mod):
    # the reference's only executable check on this path (unit_test_inverseDistortion.cpp; values in SURVEY.md §4)
    b = oracle_mod.inverse_radial([-0.34991902, -0.014698517, 0.59684463, 0.0])
    np.testing.assert_allclose(b, [0.34991902, 0.38202847867328127, -0.041555343865844696, -1.1638270394205459,
                                   -4.138165444396021], rtol=1e-15)
    assert abs(oracle_mod.inverse_distortion_roundtrip() - 0.0269245980857) < 1e-12
    from eventcalib_b200 import synth
    np.testing.assert_allclose(synth.inverse_radial_distortion([-0.34991902, -0.014698517, 0.59684463, 0.0]), b, rtol=1e-15)


def test_basis_and_knots(oracle_mod):
    from scipy.interpolate import BSpline
    from eventcalib_b200 import spline
    us = np.sort(np.random.default_rng(0).uniform(2.0, 3.0, 60))
    for n_cp in (4, 5, 9, 17):
        kn = oracle_mod.knots(us, n_cp)
        np.testing.assert_array_equal(kn, spline.knot_vector(us, n_cp))
        assert np.all(np.diff(kn) >= 0) and kn[0] == us[0] and kn[-1] == us[-1]
        for u in list(np.random.default_rng(1).uniform(us[0], us[-1], 50)) + [us[0], us[-1]]:
            sp, N = oracle_mod.basis(kn, u)
            assert abs(N.sum() - 1) < 1e-12 and N.min() >= -1e-15
            assert sp == spline.find_span(kn, u)
            np.testing.assert_allclose(N, spline.basis(kn, sp, u), rtol=1e-14, atol=1e-16)
            if us[0] < u < us[-1]:
                for j in range(4):
                    c = np.zeros(n_cp)
                    c[sp - 3 + j] = 1
                    assert abs(BSpline(kn, c, 3)(u) - N[j]) < 1e-12


def _case(rng):
    from eventcalib_b200 import synth
    cam, board = synth.Camera(), synth.Board()
    traj = synth.Trajectory(int(rng.integers(1, 100)), board, 78.0, rot_amp=(0.3, 0.3, 0.3))
    t = rng.uniform(0, 3)
    q, tw = traj.quat_xyzw(np.array([t, t + 0.01, t + 0.02, t + 0.03]))
    Q = q + rng.normal(0, 0.01, q.shape)
    T = tw + rng.normal(0, 0.3, tw.shape)
    b = rng.dirichlet([2, 2, 2, 2])
    lm = board.centres()[rng.integers(36)]
    obs = np.array([rng.uniform(20, 320), rng.uniform(20, 240)])
    return cam.intrinsics() * (1 + rng.normal(0, 0.01, 9)), Q, T, b, obs, lm


def test_residual_against_mpmath_and_finite_differences(oracle_mod):
    import mpmath as mp
    mp.mp.dps = 40
    rng = np.random.default_rng(5)

    M = lambda v: v if isinstance(v, mp.mpf) else mp.mpf(float(v))

    def f_mp(intr, Q, T, b, obs, lm):
        intr = [M(v) for v in intr]
        q = [sum(M(b[j]) * M(Q[j][c]) for j in range(4)) for c in range(4)]
        nz = mp.sqrt(sum(v * v for v in q))
        qx, qy, qz, qw = [v / nz for v in q]
        t = [sum(M(b[j]) * M(T[j][c]) for j in range(4)) for c in range(3)]
        x = (mp.mpf(float(obs[0])) - intr[2]) / intr[0]
        y = (mp.mpf(float(obs[1])) - intr[3]) / intr[1]
        r2 = x * x + y * y
        s = 1 + sum(intr[4 + i] * r2 ** (i + 1) for i in range(5))
        X = [x * s, y * s, mp.mpf(1)]
        R3 = [2 * (qx * qz - qw * qy), 2 * (qy * qz + qw * qx), 1 - 2 * (qx * qx + qy * qy)]
        lam = -t[2] / sum(a * c for a, c in zip(R3, X))
        v = [lam * c for c in X]
        u = [qx, qy, qz]
        cr = lambda a, c: [a[1] * c[2] - a[2] * c[1], a[2] * c[0] - a[0] * c[2], a[0] * c[1] - a[1] * c[0]]
        uv = [2 * c for c in cr(u, v)]
        w = cr(u, uv)
        Xw = [v[i] + qw * uv[i] + w[i] + t[i] for i in range(3)]
        return mp.sqrt(sum((Xw[i] - mp.mpf(float(lm[i]))) ** 2 for i in range(3))) - mp.mpf("1.75")

    for _ in range(20):
        intr, Q, T, b, obs, lm = _case(rng)
        r, jac = oracle_mod.residual_jac(intr, Q, T, obs, lm, 1.75, b)
        assert abs(r - float(f_mp(intr, Q, T, b, obs, lm))) < 1e-11
        # central differences in 40-digit arithmetic for a few ambient partials
        flat = np.concatenate([intr, Q.ravel(), T.ravel()])
        for k in rng.choice(37, 6, replace=False):
            h = 1e-12

            def at(d):
                z = [mp.mpf(float(v)) for v in flat]
                z[k] += d
                zi = z[:9]
                zq = [z[9 + 4 * j: 13 + 4 * j] for j in range(4)]
                zt = [z[25 + 3 * j: 28 + 3 * j] for j in range(4)]
                return f_mp(zi, zq, zt, b, obs, lm)
            fd = float((at(mp.mpf(h)) - at(-mp.mpf(h))) / (2 * mp.mpf(h)))
            assert abs(fd - jac[k]) <= 1e-7 * max(1.0, abs(jac[k]))


def test_closed_form_jacobian_matches_dual_numbers(oracle_mod):
    so = os.path.join(ROOT, "tests", "_build", "libresid_host.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", so,
                           os.path.join(ROOT, "tests", "helpers", "residual_host.cpp")])
    h = C.CDLL(so)
    h.host_residual.restype = C.c_double
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    rng = np.random.default_rng(0)
    worst = 0.0
    for it in range(500):
        intr, Q, T, b, obs, lm = _case(rng)
        if it % 3 == 0:
            obs = obs + rng.normal(0, 25, 2)   # far off the rim: exercises the Huber branch
        Q, T = np.ascontiguousarray(Q), np.ascontiguousarray(T)
        r, jac = oracle_mod.residual_jac(intr, Q, T, obs, lm, 1.75, b)
        J = np.zeros(33)
        cost, raw = C.c_double(), C.c_double()
        res = h.host_residual(P(intr), P(Q), P(T), P(b), P(obs), P(lm), C.c_double(1.75), C.c_double(0.35), P(J),
                              C.byref(cost), C.byref(raw))
        Jr = np.zeros(33)
        Jr[:9] = jac[:9]
        for j in range(4):
            x, y, z, w = Q[j]
            Jr[9 + 3 * j: 12 + 3 * j] = jac[9 + 4 * j: 13 + 4 * j] @ np.array([[w, z, -y], [-z, w, x], [y, -x, w], [-x, -y, -z]])
        Jr[21:] = jac[25:]
        s = r * r
        rho1 = 1.0 if s <= 0.35 ** 2 else 0.35 / np.sqrt(s)
        rho = s if s <= 0.35 ** 2 else 2 * 0.35 * np.sqrt(s) - 0.35 ** 2
        Jr *= np.sqrt(rho1)
        worst = max(worst, np.abs(J - Jr).max() / np.abs(Jr).max(), abs(raw.value - r) / max(abs(r), 1e-9))
        assert abs(cost.value - 0.5 * rho) <= 1e-12 * max(rho, 1e-12)
        assert abs(res - np.sqrt(rho1) * r) <= 1e-12 * max(abs(r), 1e-9)
    assert worst < 1e-11


def test_association_and_normal_equations_consistency(oracle_mod):
    from eventcalib_b200 import synth, calib_problem
    ev = synth.make_stream(20000, 346, 260, t0=5.0, duration=0.2, seed=4, return_truth=True)
    pb = calib_problem.build(ev, seed=0)
    P = oracle_mod.CostProblem([pb["n_cp"]], [pb["knots"]], pb["radius"], pb["huber"])
    oe, oc = P.associate(ev["t"], ev["x"], ev["y"], pb["kf_t"], pb["circles"], pb["landmarks"], pb["step"])
    assert 0.8 * len(ev["t"]) < len(oe) <= len(ev["t"]) and np.all(np.diff(oe) > 0) and oc.min() >= 0 and oc.max() < 36
    x = (pb["intrinsics"], pb["rot_cp"], pb["trans_cp"])
    c, H, g, r, J = P.normal_eq(*x, want_rows=True)
    assert abs(c - P.cost(*x)) <= 1e-12 * c
    c_mt, c2 = P.eval_mt(*x, 3)
    assert abs(c_mt - c) <= 1e-12 * c and abs(c2 - c) <= 1e-12 * c
    sp = np.zeros(len(oe), np.int32)
    P.lib.orc_problem_get_records(P.h, None, sp.ctypes.data_as(C.c_void_p))
    for s in range(P.n_spans):
        m = sp == s
        np.testing.assert_allclose(H[s], J[m].T @ J[m], rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(g[s], J[m].T @ r[m], rtol=1e-12, atol=1e-9)


# ---------------------------------------------------------------------------------------------------- SO(3) variant ----
def _so3_case(rng):
    intr, Q, T, b, obs, lm = _case(rng)
    Q = Q / np.linalg.norm(Q, axis=1, keepdims=True)      # Sophus::SO3d control points are unit quaternions
    return intr, np.ascontiguousarray(Q), np.ascontiguousarray(T), b, obs, lm


def _so3_value_scipy(intr, Q, T, b, obs, lm):
    """independent evaluation of CalibReprojectionError_SO3 (EventCalibSpline.hpp:78-140) with scipy rotations"""
    from scipy.spatial.transform import Rotation as Rot
    beta = [b[1] + b[2] + b[3], b[2] + b[3], b[3]]
    R = [Rot.from_quat(q) for q in Q]
    Rw = R[0]
    for j in range(1, 4):
        Rw = Rw * Rot.from_rotvec(beta[j - 1] * (R[j - 1].inv() * R[j]).as_rotvec())
    tw = sum(b[j] * T[j] for j in range(4))
    fx, fy, cx, cy = intr[:4]
    x, y = (obs[0] - cx) / fx, (obs[1] - cy) / fy
    r2 = x * x + y * y
    s = 1 + sum(intr[4 + i] * r2 ** (i + 1) for i in range(5))
    Xc = np.array([x * s, y * s, 1.0])
    M = Rw.as_matrix()
    depth = -tw[2] / (M[2] @ Xc)
    Xw = M @ (depth * Xc) + tw
    return np.linalg.norm(Xw - lm) - 1.75


def test_so3_residual_oracle_against_scipy_and_finite_differences(oracle_mod):
    from scipy.spatial.transform import Rotation as Rot
    rng = np.random.default_rng(17)
    for it in range(40):
        intr, Q, T, b, obs, lm = _so3_case(rng)
        r, jac = oracle_mod.residual_jac_so3(intr, Q, T, obs, lm, 1.75, b)
        ref = _so3_value_scipy(intr, Q, T, b, obs, lm)
        assert abs(r - ref) <= 1e-10 * max(1.0, abs(ref))
        # tangent-space derivative along T_j * exp(delta): ambient Jacobian x Dx_this_mul_exp_x_at_0 vs central differences
        for j in range(4):
            Jl = jac[9 + 4 * j: 13 + 4 * j] @ oracle_mod.so3_plus_jacobian(Q[j])
            for k in range(3):
                h = 1e-6
                vals = []
                for sgn in (1, -1):
                    d = np.zeros(3)
                    d[k] = sgn * h
                    Q2 = Q.copy()
                    Q2[j] = (Rot.from_quat(Q[j]) * Rot.from_rotvec(d)).as_quat()
                    if np.dot(Q2[j], Q[j]) < 0:
                        Q2[j] = -Q2[j]
                    vals.append(_so3_value_scipy(intr, Q2, T, b, obs, lm))
                fd = (vals[0] - vals[1]) / (2 * h)
                assert abs(fd - Jl[k]) <= 2e-6 * max(1.0, abs(Jl[k]))
        # LocalParameterizationSO3::Plus
        d = rng.normal(0, 0.2, 3)
        want = (Rot.from_quat(Q[0]) * Rot.from_rotvec(d)).as_quat()
        got = oracle_mod.so3_plus(Q[0], d)
        assert min(np.abs(got - want).max(), np.abs(got + want).max()) < 1e-14


def test_so3_tangent_jacobian_matches_dual_numbers(oracle_mod):
    """product header (ecb_residual_so3.h: 3-partial duals along the tangent) vs the oracle (Jet<37> in the ambient space
    times Sophus' Dx_this_mul_exp_x_at_0, what Ceres does with LocalParameterizationSO3) — incl. small relative rotations
    that take the Taylor branches of Sophus exp / log"""
    so = os.path.join(ROOT, "tests", "_build", "libresid_host.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", so,
                           os.path.join(ROOT, "tests", "helpers", "residual_host.cpp")])
    h = C.CDLL(so)
    h.host_residual_so3.restype = C.c_double
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    rng = np.random.default_rng(3)
    worst = 0.0
    for it in range(400):
        intr, Q, T, b, obs, lm = _so3_case(rng)
        if it % 3 == 0:
            obs = obs + rng.normal(0, 25, 2)
        if it % 7 == 0:
            Q[2] = Q[1]                       # identical neighbours: log / exp small-angle branches
        if it % 11 == 0:
            Q[1] = -Q[1]                      # the double cover: same rotation, opposite sign
        r, jac = oracle_mod.residual_jac_so3(intr, Q, T, obs, lm, 1.75, b)
        J = np.zeros(33)
        cost, raw = C.c_double(), C.c_double()
        res = h.host_residual_so3(P(intr), P(Q), P(T), P(b), P(obs), P(lm), C.c_double(1.75), C.c_double(0.35), P(J),
                                  C.byref(cost), C.byref(raw))
        Jr = np.zeros(33)
        Jr[:9] = jac[:9]
        for j in range(4):
            Jr[9 + 3 * j: 12 + 3 * j] = jac[9 + 4 * j: 13 + 4 * j] @ oracle_mod.so3_plus_jacobian(Q[j])
        Jr[21:] = jac[25:]
        s = r * r
        rho1 = 1.0 if s <= 0.35 ** 2 else 0.35 / np.sqrt(s)
        rho = s if s <= 0.35 ** 2 else 2 * 0.35 * np.sqrt(s) - 0.35 ** 2
        Jr *= np.sqrt(rho1)
        worst = max(worst, np.abs(J - Jr).max() / np.abs(Jr).max(), abs(raw.value - r) / max(abs(r), 1e-9))
        assert abs(cost.value - 0.5 * rho) <= 1e-12 * max(rho, 1e-12)
    assert worst < 1e-10
    # Plus
    h.host_so3_plus.restype = None
    for it in range(20):
        x = rng.normal(size=4)
        x /= np.linalg.norm(x)
        d = rng.normal(0, 10.0 ** rng.uniform(-12, 0), 3)
        out = np.zeros(4)
        h.host_so3_plus(P(x), P(d), P(out))
        assert np.abs(out - oracle_mod.so3_plus(x, d)).max() < 1e-15
