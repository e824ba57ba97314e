"""The restatement (oracle/, always buildable) and the host side of the product against tests/golden/reference_source.npz —
outputs of the reference's OWN sources compiled in place (written by tests/golden/make_reference_source_golden.py in the
build container).  Unlike tests/test_oracle_reference_source.py these checks need neither /root/reference nor the prebuilt
oracle/_ref library.  Bit-exact unless a tolerance is written."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "reference_source.npz"))
P = lambda a: a.ctypes.data_as(C.c_void_p)


def _events():
    return G["ev_t"], G["ev_x"].astype(np.float64), G["ev_y"].astype(np.float64), G["ev_p"]


def test_functor_golden(oracle_mod):
    """a9 / a10: restated functor on Jet<37> == the reference functor's value and 37 partials."""
    for i in range(len(G["functor_r"])):
        r, jac = oracle_mod.residual_jac(G["functor_intr"][i], G["functor_rcp"][i], G["functor_tcp"][i], G["functor_obs"][i],
                                         G["functor_lm"][i], 1.75, G["functor_b"][i])
        assert r == G["functor_r"][i]
        np.testing.assert_array_equal(jac, G["functor_jac"][i])


def test_product_residual_header_golden():
    """csrc/ecb_residual.h (host build) vs the reference functor: 1e-9 relative (north star)."""
    so = os.path.join(ROOT, "tests", "_build", "libresid_host.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-o", so,
                           os.path.join(ROOT, "tests", "helpers", "residual_host.cpp")])
    h = C.CDLL(so)
    h.host_residual.restype = C.c_double
    for i in range(len(G["functor_r"])):
        Q, T = np.ascontiguousarray(G["functor_rcp"][i]), np.ascontiguousarray(G["functor_tcp"][i])
        jac, r1 = G["functor_jac"][i], float(G["functor_r"][i])
        J = np.zeros(33)
        cost, raw = C.c_double(), C.c_double()
        res = h.host_residual(P(G["functor_intr"][i].copy()), P(Q), P(T), P(G["functor_b"][i].copy()), P(G["functor_obs"][i].copy()),
                              P(G["functor_lm"][i].copy()), C.c_double(1.75), C.c_double(1e30), P(J), C.byref(cost), C.byref(raw))
        Jr = np.zeros(33)
        Jr[:9] = jac[:9]
        for k in range(4):
            x, y, z, w = Q[k]
            Jr[9 + 3 * k: 12 + 3 * k] = jac[9 + 4 * k: 13 + 4 * k] @ np.array([[w, z, -y], [-z, w, x], [y, -x, w], [-x, -y, -z]])
        Jr[21:] = jac[25:]
        assert abs(res - r1) <= 1e-9 * max(1.0, abs(r1))
        assert np.abs(J - Jr).max() <= 1e-9 * np.abs(Jr).max()


def _so3_golden_rows():
    """per residual of the golden SO(3) problem: (span, 1 x 33 tangent Jacobian row, residual) from the reference functor's
    Jet<37> output and the reference's LocalParameterizationSO3 Jacobians"""
    span, jac, r = G["so3_span"], G["so3_jac"], G["so3_r"]
    PJ = G["so3_plus_jac"]
    rows = np.zeros((len(r), 33))
    rows[:, :9] = jac[:, :9]
    for k in range(4):
        rows[:, 9 + 3 * k: 12 + 3 * k] = np.einsum("na,nab->nb", jac[:, 9 + 4 * k: 13 + 4 * k], PJ[span - 3 + k])
    rows[:, 21:] = jac[:, 25:]
    return span, rows, r


def so3_golden_normal_equations(huber=0.35):
    """J^T J / J^T r per span and the cost as Ceres builds them from the reference functor's rows (Huber corrector: residual and
    row scaled by sqrt(rho'), rho'' <= 0 for Huber)"""
    span, rows, r = _so3_golden_rows()
    n_spans = len(G["so3_knots"]) - 4 - 3
    H, g, cost = np.zeros((n_spans, 33, 33)), np.zeros((n_spans, 33)), 0.0
    for k in range(len(r)):
        s2 = r[k] * r[k]
        rho1 = 1.0 if s2 <= huber ** 2 else huber / np.sqrt(s2)
        cost += 0.5 * (s2 if s2 <= huber ** 2 else 2 * huber * np.sqrt(s2) - huber ** 2)
        row, rr = rows[k] * np.sqrt(rho1), r[k] * np.sqrt(rho1)
        H[span[k] - 3] += np.outer(row, row)
        g[span[k] - 3] += row * rr
    return H, g, cost


def test_so3_functor_golden(oracle_mod):
    """a11: restated SO(3) functor on Jet<37> == the reference's CalibReprojectionError_SO3 (value and 37 partials), the
    cumulative basis and LocalParameterizationSO3, bit for bit; the oracle's normal equations of the whole golden problem
    equal the ones assembled from the reference functor's rows (1e-12)."""
    kn, Q, T, intr = G["so3_knots"], G["so3_rot_cp"], G["so3_trans_cp"], G["so3_intrinsics"]
    ev_t, ev_x, ev_y = G["so3_ev_t"], G["so3_ev_x"].astype(np.float64), G["so3_ev_y"].astype(np.float64)
    for k in range(0, len(G["so3_r"]), 3):
        e, sp = int(G["so3_event"][k]), int(G["so3_span"][k])
        r, jac = oracle_mod.residual_jac_so3(intr, Q[sp - 3:sp + 1], T[sp - 3:sp + 1], np.array([ev_x[e], ev_y[e]]),
                                             G["so3_landmarks"][G["so3_circle"][k]], 1.75, G["so3_N"][k])
        assert r == G["so3_r"][k]
        np.testing.assert_array_equal(jac, G["so3_jac"][k])
        N = G["so3_N"][k]
        np.testing.assert_array_equal(G["so3_beta"][k], [(N[3] + N[2]) + N[1], N[3] + N[2], N[3]])
    for x, d, o in zip(G["so3_plus_x"], G["so3_plus_d"], G["so3_plus_out"]):
        np.testing.assert_array_equal(oracle_mod.so3_plus(x, d), o)
    for q, J in zip(Q, G["so3_plus_jac"]):
        np.testing.assert_array_equal(oracle_mod.so3_plus_jacobian(q), J)
    Pc = oracle_mod.CostProblem([len(Q)], [kn], 1.75, 0.35, so3=True)
    oe, oc = Pc.associate(ev_t, ev_x, ev_y, G["so3_kf_t"], G["so3_circles"], G["so3_landmarks"], float(G["so3_step"]))
    np.testing.assert_array_equal(oe, G["so3_event"])
    np.testing.assert_array_equal(oc, G["so3_circle"].astype(oc.dtype))
    c, H, g = Pc.normal_eq(intr, Q, T)
    Hg, gg, cg = so3_golden_normal_equations()
    assert abs(c - cg) <= 1e-12 * cg
    assert np.abs(H - Hg).max() <= 1e-12 * np.abs(Hg).max() and np.abs(g - gg).max() <= 1e-12 * np.abs(gg).max()


def test_spline_golden(oracle_mod):
    """a8 + spline set-up: knots, spans, basis bit-exact; façade fit 1e-12."""
    from eventcalib_b200 import spline
    import eventcalib_b200.build as b
    b.build()
    so = os.path.join(ROOT, "tests", "_build", "libfacade_host.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tests", "helpers", "facade_host.cpp"),
                           "-L" + os.path.join(ROOT, "eventcalib_b200"), "-lecb",
                           "-Wl,-rpath," + os.path.join(ROOT, "eventcalib_b200")])
    F = C.CDLL(so)
    us = G["basis_us"]
    for n_cp in (4, 9, 20):
        kn = G[f"basis_knots_{n_cp}"]
        np.testing.assert_array_equal(oracle_mod.knots(us, n_cp), kn)
        np.testing.assert_array_equal(spline.knot_vector(us, n_cp), kn)
        for data, key in ((G["basis_tw"], "cp3"), (G["basis_q"], "cp4")):
            data = np.ascontiguousarray(data)
            dim = data.shape[1]
            kn2, cp2 = np.zeros(n_cp + 4), np.zeros((n_cp, dim))
            F.fh_fit_spline(P(us.copy()), P(data), len(us), dim, n_cp, P(kn2), P(cp2))
            np.testing.assert_array_equal(kn2, kn)
            ref_cp = G[f"basis_{key}_{n_cp}"]
            assert np.abs(cp2 - ref_cp).max() <= 1e-12 * np.abs(ref_cp).max()
    kn = G["basis_knots_20"]
    for u, sp, N in zip(G["basis_u"], G["basis_span"], G["basis_N"]):
        s2, N2 = oracle_mod.basis(kn, float(u))
        assert s2 == sp == spline.find_span(kn, float(u))
        np.testing.assert_array_equal(N2, N)


def test_event_frame_extract_rectify_golden(oracle_mod):
    """a2 - a7 on raw events."""
    from test_circles_grid import _lib as grid_lib, _order as grid_order
    glib = grid_lib()
    t, x, y, p = _events()
    rthr = float(G["rthr"])
    assert rthr == oracle_mod.radius_threshold(346, 260, 9, 4, 1, 5.5, 1.75)
    found = verdicts = 0
    for i, w in enumerate(G["windows"]):
        P0, N0, _, _ = oracle_mod.event_frame(t, x, y, p, float(w[0]), float(w[1]))
        np.testing.assert_array_equal(P0, G[f"frame_pos_{i}"].astype(np.float64))
        np.testing.assert_array_equal(N0, G[f"frame_neg_{i}"].astype(np.float64))
        for fit in (0, 1):
            r0 = oracle_mod.extract(P0, N0, fitCircle=fit, Rthr=rthr)
            if not bool(G[f"extract_reached_{i}_{fit}"]):
                assert not r0["enough"]
                continue
            c0 = r0["cand"]
            np.testing.assert_array_equal(c0[:, 2:4].astype(np.float32), G[f"extract_cand_{i}_{fit}"])
            ok, order = grid_order(glib, c0[:, 2:4].astype(np.float32).astype(np.float64))
            assert ok == bool(G[f"extract_found_{i}_{fit}"])
            if ok:
                found += 1
                np.testing.assert_array_equal(c0[order][:, 2:5], G[f"extract_features_{i}_{fit}"])
            rc = int(G[f"rectify_rc_{i}_{fit}"])
            if rc < 0:
                continue
            out0, ok0 = oracle_mod.rectify(P0, N0, G[f"rectify_img_{i}_{fit}"], 346, 260, fitCircle=fit)
            out1 = G[f"rectify_out_{i}_{fit}"]
            assert rc == int(ok0)
            keep = out1[:, 2] >= 0
            np.testing.assert_array_equal(out0[:, 2] < 0, ~keep)
            np.testing.assert_array_equal(out0[keep], out1[keep])
            verdicts += 1
            if rc == 1:   # findCenter on the rectified frame
                fxy, ids = G[f"rectify_fxy_{i}_{fit}"], np.nonzero(keep)[0]
                d2 = ((fxy[:, None, :] - out1[ids][None, :, :2]) ** 2).sum(-1)
                best = d2.argmin(1)
                acc = np.abs(np.sqrt(d2[np.arange(len(fxy)), best]) - out1[ids][best, 2]) < 5
                np.testing.assert_array_equal(G[f"rectify_fid_{i}_{fit}"], np.where(acc, ids[best], -1))
    assert found >= 4 and verdicts >= 4
    for pxy, nxy, o in zip(G["fit_p"], G["fit_n"], G["fit_out"]):
        np.testing.assert_array_equal(oracle_mod.fit_circle(pxy, nxy), o)


def test_calib_spline_golden(oracle_mod):
    """a7 + a12 assembly + spline set-up of the reference's EventCalibSpline constructor."""
    from eventcalib_b200 import synth
    import eventcalib_b200.build as b
    b.build()
    so = os.path.join(ROOT, "tests", "_build", "libfacade_host.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tests", "helpers", "facade_host.cpp"),
                           "-L" + os.path.join(ROOT, "eventcalib_b200"), "-lecb",
                           "-Wl,-rpath," + os.path.join(ROOT, "eventcalib_b200")])
    F = C.CDLL(so)
    board = synth.Board()
    step = 5e-4
    kf_t, q, tw, circ = G["spline_kf_t"], G["spline_kf_q"], G["spline_kf_tw"], G["spline_circ"]
    n_cp, K = G["spline_n_cp"], len(G["spline_kf_t"])
    knots = []
    for w in range(len(n_cp)):
        ncp = C.c_int()
        kn, rot, tr = np.zeros(K + 8), np.zeros((K, 4)), np.zeros((K, 3))
        ns = F.fh_segments(P(kf_t.copy()), P(np.ascontiguousarray(q)), P(np.ascontiguousarray(tw)), K, C.c_double(step), w,
                           C.byref(ncp), P(kn), P(rot), P(tr))
        n = ncp.value
        assert ns == len(n_cp) and n == n_cp[w]
        np.testing.assert_array_equal(kn[:n + 4], G[f"spline_knots_{w}"])
        np.testing.assert_allclose(rot[:n], G[f"spline_rot_{w}"], rtol=0, atol=1e-13)
        np.testing.assert_allclose(tr[:n], G[f"spline_trans_{w}"], rtol=1e-13, atol=1e-13)
        knots.append(G[f"spline_knots_{w}"])
    cam9 = G["spline_cam9"]
    np.testing.assert_array_equal(G["spline_intrinsics"], np.r_[cam9[:4], oracle_mod.inverse_radial([cam9[4], cam9[5], cam9[8], 0.0])])
    left = ~np.isnan(G["spline_kf_pose"][:, 0])
    assert left.sum() == int(G["spline_frames_left"]) == K - 3
    t, x, y = G["spline_ev_t"], G["spline_ev_x"].astype(np.float64), G["spline_ev_y"].astype(np.float64)
    cp = oracle_mod.CostProblem(n_cp, knots, radius=board.radius, huber=0.2 * board.radius)
    oe, oc = cp.associate(t, x, y, kf_t[left], circ[left], board.centres(), step)
    assert len(oe) == len(G["spline_span"]) > 10000
    np.testing.assert_array_equal(np.c_[x[oe], y[oe]], G["spline_obs"].astype(np.float64))
    np.testing.assert_array_equal(oc, G["spline_lm_idx"].astype(np.int32))
    np.testing.assert_array_equal(np.searchsorted(G["spline_ranges"][:, 1], t[oe], side="left"), G["spline_spline"])
    for j, i in enumerate(range(0, len(oe), 101)):
        sp, N = oracle_mod.basis(knots[int(G["spline_spline"][i])], float(t[oe[i]]))
        assert sp == G["spline_span"][i]
        np.testing.assert_array_equal(N, G["spline_basis_sample"][j])
    np.testing.assert_array_equal(G["spline_first_cp"][:, 0], G["spline_span"] - 3)
    np.testing.assert_array_equal(G["spline_first_cp"][:, 1], G["spline_span"] - 3)
    pb, qb, solver, calls = (int(v) for v in G["spline_assembly"])
    assert pb == qb == int(n_cp.sum()) and solver == 1 and calls == 1
    assert abs(G["spline_huber_tol"][0] - 0.2 * board.radius) < 1e-15 and G["spline_huber_tol"][1] == 1e-10 and G["spline_huber_tol"][2] == 1e-10


def test_gate_and_check_pose_golden():
    """f-1 / f-4: the façade's TrackingGate and ecb::checkPose against the decisions of the reference's EventCalibIni."""
    import eventcalib_b200.build as b
    b.build()
    os.makedirs(os.path.join(ROOT, "tests", "_build"), exist_ok=True)
    so = os.path.join(ROOT, "tests", "_build", "libfacade_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tests", "helpers", "facade_host.cpp"),
                           "-L" + os.path.join(ROOT, "eventcalib_b200"), "-lecb",
                           "-Wl,-rpath," + os.path.join(ROOT, "eventcalib_b200")])
    so2 = os.path.join(ROOT, "tests", "_build", "libcalib_init_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so2,
                           os.path.join(ROOT, "tests", "helpers", "calib_init_host.cpp")])
    F, CI = C.CDLL(so), C.CDLL(so2)
    F.fh_gate_new.restype = C.c_void_p
    CI.ci_check_pose.argtypes = [C.c_double, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_double]
    for trial in (0, 1):
        g = C.c_void_p(F.fh_gate_new(9, 4, C.c_double(5e-4)))
        acc = [int(F.fh_gate_process(g, C.c_double(float(tt)), P(np.ascontiguousarray(xy)), 36))
               for tt, xy in zip(G["gate_ts_%d" % trial], G["gate_xy_%d" % trial])]
        F.fh_gate_free(g)
        np.testing.assert_array_equal(acc, G["gate_accept_%d" % trial])
        assert 0 < sum(acc) < len(acc)
    for row in G["pose_cases"]:
        q0, t0, dt, q1, t1, verdict = row[:4].copy(), row[4:7].copy(), float(row[7]), row[8:12].copy(), row[12:15].copy(), int(row[15])
        assert CI.ci_check_pose(1.0, P(q0), P(t0), 1.0 + dt, P(q1), P(t1), 5e-4) == verdict
