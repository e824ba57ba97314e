"""Pins the cost-evaluation oracle to the reference's OWN source text: oracle/_ref/libref_functor.so is the UNMODIFIED
`CalibReprojectionError::operator()` / `unDistort` (event_camera_calib/include/opengv2/event_camera_calib/EventCalibSpline.hpp:
36-63,158-250) and `BsplineReal` (core/spline/include/opengv2/spline/BsplineReal.hpp) compiled where they lie against the
stand-in Eigen / Ceres / Sophus headers of oracle/shim_functor/ (oracle/Makefile, oracle/ref_functor_capi.cpp).

  * restated functor (oracle/ecb_oracle_cost.cpp) == reference functor on Jet<37>: value and all 37 partials BIT-EXACT
  * reference functor on double vs on Jet: 1e-12 relative (ceres' Jet divides by multiplying with the reciprocal)
  * the product's closed-form residual / Jacobian (csrc/ecb_residual.h, host build) vs the reference functor: 1e-9 relative
  * knot vector (NURBS-book eq. 9.68), findSpan, dersBasisFuns: bit-exact for the oracle, the Python builder and the façade
  * control points of the reference's constructor fit vs the façade's EventCalibSpline::fitSpline: 1e-12 relative (the
    stand-in LDLT factorises in another order than Eigen's SimplicialLDLT; same normal equations)

  * event window (a2): pixel lists of both polarities after dedupe / cancellation, in hash-set iteration order, element for
    element equal between the restatement (oracle/ecb_oracle_frontend.cpp) and the reference's EventFrame.cpp compiled in
    place with its own EigenMatrixHash (utility.hpp:38-51); record reader (a1): Event.hpp's operator>> reads our .bin files

  * extractFeatures / fitCircle / rectifyFeatures / findCenter (a3-a7): the reference's CirclesEventFrame.cpp with its own
    DBSCAN (dbscan.h + kdtree.cpp) and the real std::nth_element, compiled in place; hooks stand in for the three OpenCV
    calls (findCirclesGrid = the product's grid finder on the candidate centres the reference hands over, projectPoints =
    the caller's 5 image points per circle), exhaustive searches for nanoflann.  The restatement (oracle.extract /
    fit_circle / rectify, what the GPU tests compare with) gives the same candidate lists, and BIT-IDENTICAL feature centres,
    radii, rectified features and frame verdicts

  * spline set-up, association and problem assembly (a7, a12): the reference's EventCalibSpline.cpp constructor compiled in
    place (reduceMap segmentation, 3-step time extension, cpNum rule, BsplineReal fits, intrinsics with the inverse radial
    polynomial, association loop, AddParameterBlock / AddResidualBlock calls into a RECORDING ceres::Problem, no-op Solve,
    updateMap): segments / knots / control points equal the façade's (1e-15), the (event, landmark) residual list, spans and
    basis values equal the oracle's association exactly, every residual touches control points span-3 .. span, rotation
    blocks carry EigenQuaternionParameterization, HuberLoss(0.2 r), tolerances 1e-10, SPARSE_NORMAL_CHOLESKY

  * tracking gate, checkPose and the cvCalibration flow (rows f-1, f-4): the reference's EventCalibIni.cpp compiled in place
    with the product's host header (include/ecb/calib_init.hpp) behind the calibrateCamera / solvePnPRansac / projectPoints /
    Rodrigues hooks: the façade's TrackingGate and ecb::checkPose take the same decisions as the reference's own code, and the
    reference's sequential loop (frame selection, pose conversion, checkPose chain, rectifyFeatures) keeps the same frames
    with the same features as the batched-then-replayed logic of the façade / CLI

The library is built in the build container (where /root/reference exists) and travels as a prebuilt file; without it the
tests skip."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = lambda a: a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def ref(oracle_mod):
    oracle_mod.build()
    if not oracle_mod.have_ref_functor():
        pytest.skip("oracle/_ref/libref_functor.so not built (no /root/reference here)")
    return oracle_mod


def _random_case(rng, cam):
    intr = cam.intrinsics() * (1 + rng.normal(0, 0.01, 9))
    q = rng.normal(0, 1, 4)
    q /= np.linalg.norm(q)
    rcp = np.ascontiguousarray(q[None, :] + rng.normal(0, 0.05, (4, 4)))
    tcp = np.ascontiguousarray(np.array([20, 20, -75.0])[None, :] + rng.normal(0, 2, (4, 3)))
    obs = np.array([rng.integers(0, 346), rng.integers(0, 260)], float)
    lm = np.array([rng.uniform(0, 40), rng.uniform(0, 44), 0.0])
    b = rng.uniform(0, 1, 4)
    b /= b.sum()
    return intr, rcp, tcp, obs, lm, b


def test_restated_functor_is_bit_identical_to_the_reference_source(ref):
    from eventcalib_b200 import synth
    rng = np.random.default_rng(2024)
    cam = synth.Camera()
    for _ in range(1000):
        intr, rcp, tcp, obs, lm, b = _random_case(rng, cam)
        r0, j0 = ref.residual_jac(intr, rcp, tcp, obs, lm, 1.75, b)
        r1, j1, rd = ref.ref_residual_jac(intr, rcp, tcp, obs, lm, 1.75, b)
        assert r0 == r1
        np.testing.assert_array_equal(j0, j1)
        assert abs(rd - r1) <= 1e-12 * max(1.0, abs(r1))


def _so3_random_case(rng, cam, it):
    intr, rcp, tcp, obs, lm, b = _random_case(rng, cam)
    rcp = np.ascontiguousarray(rcp / np.linalg.norm(rcp, axis=1, keepdims=True))   # Sophus::SO3d coefficients are unit quaternions
    if it % 7 == 0:
        rcp[2] = rcp[1]           # identical neighbours: the Taylor branches of Sophus' exp / log
    if it % 11 == 0:
        rcp[1] = -rcp[1]          # the double cover
    if it % 13 == 0:
        rcp[3] = rcp[3] * (1 + 1e-7)   # slightly off the unit sphere, like a control point after many Plus steps
    beta = np.zeros(3)            # BsplineSO3::derBasisFuns (BsplineSO3.cpp:88-92)
    beta[2] = b[3]
    beta[1] = beta[2] + b[2]
    beta[0] = beta[1] + b[1]
    return intr, rcp, tcp, obs, lm, b, beta


def test_restated_so3_functor_is_bit_identical_to_the_reference_source(ref):
    """a11: the reference's CalibReprojectionError_SO3::operator() (EventCalibSpline.hpp:65-156), compiled where it lies
    against the stand-in Sophus, on Jet<37>: the restated functor of oracle/ecb_oracle_cost.cpp gives the same value and the
    same 37 partials, bit for bit."""
    from eventcalib_b200 import synth
    rng = np.random.default_rng(2025)
    cam = synth.Camera()
    for it in range(1000):
        intr, rcp, tcp, obs, lm, b, beta = _so3_random_case(rng, cam, it)
        r0, j0 = ref.residual_jac_so3(intr, rcp, tcp, obs, lm, 1.75, b)
        r1, j1, rd = ref.ref_residual_jac_so3(intr, rcp, tcp, obs, lm, 1.75, beta, b)
        assert r0 == r1
        np.testing.assert_array_equal(j0, j1)
        assert abs(rd - r1) <= 1e-12 * max(1.0, abs(r1))


def test_so3_basis_and_local_parameterization_vs_reference_source(ref):
    """BsplineSO3::findSpan / derBasisFuns (core/spline/src/BsplineSO3.cpp:73-109, compiled where it lies): the cumulative basis
    the product derives from the four B-spline values is the reference's, bit for bit; LocalParameterizationSO3::Plus and
    ::ComputeJacobian (BsplineSO3.hpp:190-221) equal the restatement exactly."""
    from eventcalib_b200 import spline
    rng = np.random.default_rng(7)
    us = np.sort(rng.uniform(2.0, 2.6, 80))
    for n_cp in (4, 7, 23):
        kn = spline.knot_vector(us, n_cp)
        for u in np.r_[rng.uniform(us[0], us[-1], 300), us, kn]:
            sp, N = ref.ref_basis(kn, float(u))
            sp2, beta = ref.ref_so3_basis(kn, float(u))
            assert sp == sp2
            b2 = N[3]
            b1 = b2 + N[2]
            b0 = b1 + N[1]
            np.testing.assert_array_equal(beta, [b0, b1, b2])
    for it in range(200):
        x = rng.normal(size=4)
        x /= np.linalg.norm(x)
        d = rng.normal(size=3) * 10.0 ** rng.uniform(-13, 0.3)
        np.testing.assert_array_equal(ref.ref_so3_plus(x, d), ref.so3_plus(x, d))
        np.testing.assert_array_equal(ref.ref_so3_plus_jacobian(x), ref.so3_plus_jacobian(x))


def test_product_so3_residual_header_vs_reference_source(ref):
    """csrc/ecb_residual_so3.h (what k_normal_eq<SO3> / k_cost<SO3> evaluate; host build) against the reference's SO(3) functor:
    residual and the 1 x 33 tangent Jacobian (ambient Jet partials x the reference's LocalParameterizationSO3 Jacobian, what
    Ceres multiplies) within 1e-9 relative (north star; measured ~1e-12)."""
    from eventcalib_b200 import synth
    so = os.path.join(ROOT, "tests", "_build", "libresid_host.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-o", so,
                           os.path.join(ROOT, "tests", "helpers", "residual_host.cpp")])
    h = C.CDLL(so)
    h.host_residual_so3.restype = C.c_double
    rng = np.random.default_rng(11)
    cam = synth.Camera()
    worst = 0.0
    for it in range(500):
        intr, rcp, tcp, obs, lm, b, beta = _so3_random_case(rng, cam, it)
        if it % 13 == 0:
            rcp = np.ascontiguousarray(rcp / np.linalg.norm(rcp, axis=1, keepdims=True))
        r1, jac, _ = ref.ref_residual_jac_so3(intr, rcp, tcp, obs, lm, 1.75, beta, b)
        J = np.zeros(33)
        cost, raw = C.c_double(), C.c_double()
        h.host_residual_so3(P(intr), P(rcp), P(tcp), P(b), P(obs), P(lm), C.c_double(1.75), C.c_double(1e30), P(J), C.byref(cost),
                            C.byref(raw))
        Jr = np.zeros(33)
        Jr[:9] = jac[:9]
        for k in range(4):
            Jr[9 + 3 * k: 12 + 3 * k] = jac[9 + 4 * k: 13 + 4 * k] @ ref.ref_so3_plus_jacobian(rcp[k])
        Jr[21:] = jac[25:]
        worst = max(worst, abs(raw.value - r1) / max(1.0, abs(r1)), np.abs(J - Jr).max() / np.abs(Jr).max())
    assert worst <= 1e-9, worst


def test_undistort_reference_source(ref):
    from eventcalib_b200 import synth
    rng = np.random.default_rng(5)
    cam = synth.Camera()
    intr = cam.intrinsics()
    for _ in range(200):
        obs = np.array([rng.integers(0, 346), rng.integers(0, 260)], float)
        X = ref.ref_undistort(intr, obs)
        x, y = (obs[0] - intr[2]) / intr[0], (obs[1] - intr[3]) / intr[1]
        r2 = x * x + y * y
        s = 1.0 + intr[4] * r2 + intr[5] * r2**2 + intr[6] * r2**3 + intr[7] * r2**4 + intr[8] * r2**5
        np.testing.assert_allclose(X, [x * s, y * s, 1.0], rtol=1e-14)


def test_product_residual_header_vs_reference_source(ref):
    """csrc/ecb_residual.h (the closed-form residual + Jacobian the kernels use), built for the host like in
    test_oracle_cost.py, against the reference functor: 1e-9 relative (north star), measured ~1e-12.  The 1x37 ambient
    Jacobian of the functor goes to the 1x33 tangent one through EigenQuaternionParameterization's Plus Jacobian, like in
    Ceres; the Huber corrector is switched off (huge threshold) so that the raw functor is what is compared."""
    so = os.path.join(ROOT, "tests", "_build", "libresid_host.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-o", so,
                           os.path.join(ROOT, "tests", "helpers", "residual_host.cpp")])
    h = C.CDLL(so)
    h.host_residual.restype = C.c_double
    from eventcalib_b200 import synth
    rng = np.random.default_rng(77)
    cam = synth.Camera()
    worst = 0.0
    for _ in range(300):
        intr, Q, T, obs, lm, b = _random_case(rng, cam)
        r1, jac, _ = ref.ref_residual_jac(intr, Q, T, obs, lm, 1.75, b)
        J = np.zeros(33)
        cost, raw = C.c_double(), C.c_double()
        res = h.host_residual(P(intr), P(Q), P(T), P(b), P(obs), P(lm), C.c_double(1.75), C.c_double(1e30), P(J),
                              C.byref(cost), C.byref(raw))
        Jr = np.zeros(33)
        Jr[:9] = jac[:9]
        for k in range(4):
            x, y, z, w = Q[k]
            Jr[9 + 3 * k: 12 + 3 * k] = jac[9 + 4 * k: 13 + 4 * k] @ np.array([[w, z, -y], [-z, w, x], [y, -x, w], [-x, -y, -z]])
        Jr[21:] = jac[25:]
        assert abs(res - r1) <= 1e-9 * max(1.0, abs(r1)) and abs(raw.value - r1) <= 1e-9 * max(1.0, abs(r1))
        worst = max(worst, np.abs(J - Jr).max() / np.abs(Jr).max())
    assert worst < 1e-9, worst


def test_spline_knots_basis_and_fit_vs_reference_source(ref):
    from eventcalib_b200 import synth, spline
    lib = None
    so = os.path.join(ROOT, "tests", "_build", "libfacade_host.so")
    import eventcalib_b200.build as b
    b.build()
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tests", "helpers", "facade_host.cpp"),
                           "-L" + os.path.join(ROOT, "eventcalib_b200"), "-lecb",
                           "-Wl,-rpath," + os.path.join(ROOT, "eventcalib_b200")])
    lib = C.CDLL(so)
    board = synth.Board()
    traj = synth.Trajectory(3, board, 78.0)
    rng = np.random.default_rng(0)
    us = np.sort(rng.uniform(1.0, 1.5, 80))
    us[0], us[-1] = 1.0, 1.5
    q, tw = traj.quat_xyzw(us)
    for data in (tw, q):
        data = np.ascontiguousarray(data)
        dim = data.shape[1]
        for n_cp in (4, 5, 9, 20, 40):
            kn, cp, n = ref.ref_spline_fit(us, data, n_cp)
            assert n == n_cp
            np.testing.assert_array_equal(kn, ref.knots(us, n_cp))            # oracle restatement
            np.testing.assert_array_equal(kn, spline.knot_vector(us, n_cp))   # Python builder
            kn2, cp2 = np.zeros(n_cp + 4), np.zeros((n_cp, dim))
            lib.fh_fit_spline(P(us), P(data), len(us), dim, n_cp, P(kn2), P(cp2))   # façade: EventCalibSpline::fitSpline
            np.testing.assert_array_equal(kn, kn2)
            assert np.abs(cp - cp2).max() <= 1e-12 * np.abs(cp).max()
            np.testing.assert_array_equal(cp[0], data[0])
            np.testing.assert_array_equal(cp[-1], data[-1])
    kn = ref.knots(us, 20)
    for u in np.r_[rng.uniform(1.0, 1.5, 3000), us, kn]:
        s1, N1 = ref.ref_basis(kn, float(u))
        s2, N2 = ref.basis(kn, float(u))
        assert s1 == s2 == spline.find_span(kn, float(u))
        np.testing.assert_array_equal(N1, N2)


def test_event_frame_and_record_reader_vs_reference_source(ref, tmp_path):
    """a1 / a2: the reference's EventFrame constructor (event/src/EventFrame.cpp:10-36, closed window, per-polarity sets of
    distinct pixels, +/- cancellation, std::unordered_set iteration order with the reference's own hash) and its record
    reader (Event.hpp:41-47), compiled in place, against the restatement the GPU tests use."""
    from eventcalib_b200 import synth
    ev = synth.make_stream(60000, 346, 260, t0=5.0, duration=0.03, seed=5)
    t, x, y, p = ev["t"], ev["x"], ev["y"], ev["p"]
    windows = [(5.0, 5.0015), (5.001, 5.004), (5.0101, 5.0116), (5.02, 5.03), (4.0, 4.5), (5.0, 5.03),
               (float(t[100]), float(t[4000])),      # bounds exactly on time stamps: the window is closed on both ends
               (float(t[7]), float(t[7]))]
    for a, b in windows:
        P0, N0, lo, hi = ref.event_frame(t, x, y, p, a, b)
        P1, N1 = ref.ref_event_frame(t, x, y, p, a, b)
        np.testing.assert_array_equal(P0, P1)
        np.testing.assert_array_equal(N0, N1)
    # every rehash boundary of the hash sets: growing prefixes of one window
    sel = np.nonzero((t >= 5.0) & (t <= 5.02))[0]
    for m in (1, 2, 12, 13, 14, 28, 29, 30, 58, 59, 60, 126, 127, 128, 256, 257, 258, 540, 541, 542, 1108, 1109, 1110, 2357, 5000):
        s = sel[:m]
        P0, N0, _, _ = ref.event_frame(t[s], x[s], y[s], p[s], 0.0, 10.0)
        P1, N1 = ref.ref_event_frame(t[s], x[s], y[s], p[s], 0.0, 10.0)
        np.testing.assert_array_equal(P0, P1)
        np.testing.assert_array_equal(N0, N1)
    path = str(tmp_path / "e.bin")
    synth.write_bin(path, ev)
    t1, x1, y1, p1 = ref.ref_read_bin(path, len(t) + 8)
    assert len(t1) == len(t)
    np.testing.assert_array_equal(t1, t)
    np.testing.assert_array_equal(x1, x)
    np.testing.assert_array_equal(y1, y)
    np.testing.assert_array_equal(p1, p)


def _grid():
    from test_circles_grid import _lib as grid_lib, _order as grid_order
    return grid_lib(), grid_order


def test_extract_features_vs_reference_source(ref):
    """a3 + a4 + a5 end to end on raw events: reference constructor (window, sets), DBSCAN, cluster filter, nth_element
    medians, k-NN pairing, memoised Kasa fits, fit-error gates, mutual-best check, grid order -> features_."""
    from eventcalib_b200 import synth
    glib, grid_order = _grid()
    found = total = 0
    for seed, orbit, amp in ((1001, False, None), (7, True, (0.35, 0.35, 0.3))):
        ev = synth.make_stream(60000, 346, 260, t0=5.0, duration=0.03, seed=seed, orbit=orbit, rot_amp=amp)
        t, x, y, p = ev["t"], ev["x"], ev["y"], ev["p"]
        for fit in (0, 1):
            for w in synth.tiling_windows(5.0, 5.03, 1.5e-3)[::2]:
                a, b = float(w[0]), float(w[1])
                m = (t >= a - 1e-3) & (t <= b + 1e-3)
                r1 = ref.ref_extract(t[m], x[m], y[m], p[m], a, b, 346, 260, fit)
                P0, N0, _, _ = ref.event_frame(t, x, y, p, a, b)
                assert abs(r1["rthr"] - ref.radius_threshold(346, 260, 9, 4, 1, 5.5, 1.75)) == 0
                r0 = ref.extract(P0, N0, fitCircle=fit, Rthr=r1["rthr"])
                total += 1
                if r1["cand_f32"] is None:          # fewer than rows*cols clusters in a polarity (:127-129)
                    assert not r0["enough"]
                    continue
                c0 = r0["cand"]
                np.testing.assert_array_equal(c0[:, 2:4].astype(np.float32), r1["cand_f32"])
                ok, order = grid_order(glib, c0[:, 2:4].astype(np.float32).astype(np.float64))
                assert ok == r1["found"]
                if ok:
                    found += 1
                    np.testing.assert_array_equal(c0[order][:, 2:5], r1["features"])   # centres and radii, bit for bit
    assert found >= 10 and total >= 30


def test_fit_circle_vs_reference_source(ref):
    rng = np.random.default_rng(0)
    for _ in range(300):
        c, r = rng.uniform(50, 200, 2), rng.uniform(4, 12)
        th = rng.uniform(0, 2 * np.pi, 60)
        pts = np.rint(np.c_[c[0] + r * np.cos(th), c[1] + r * np.sin(th)] + rng.normal(0, 0.5, (60, 2)))
        np.testing.assert_array_equal(ref.fit_circle(pts[:30], pts[30:]), ref.ref_fit_circle(pts[:30], pts[30:]))


def test_rectify_and_find_center_vs_reference_source(ref):
    """a6 + a7: rectifyFeatures (radius search, quadrant band, cluster expansion, refit, gates, edge scores, 20 % rule) and
    findCenter on the rectified frame."""
    from eventcalib_b200 import synth
    board = synth.Board()
    cen, sk = board.centres(), board.radius / np.sqrt(2)
    rng = np.random.default_rng(3)
    verdicts = {}
    for seed, fit in ((1001, 0), (1001, 1), (7, 1)):
        ev = synth.make_stream(60000, 346, 260, t0=5.0, duration=0.03, seed=seed, return_truth=True)
        cam, traj = ev["camera"], ev["trajectory"]
        t, x, y, p = ev["t"], ev["x"], ev["y"], ev["p"]
        for w in synth.tiling_windows(5.0, 5.03, 1.5e-3)[::2]:
            a, b = float(w[0]), float(w[1])
            R, tw = traj.pose(np.array([(a + b) / 2]))
            img = np.zeros((36, 5, 2))
            for k in range(36):
                o5 = np.array([cen[k], cen[k] + [sk, sk, 0], cen[k] + [sk, -sk, 0], cen[k] + [-sk, -sk, 0], cen[k] + [-sk, sk, 0]])
                u, v = synth.project(cam, np.repeat(R, 5, 0), np.repeat(tw, 5, 0), o5)
                img[k, :, 0], img[k, :, 1] = u, v
            if rng.uniform() < 0.35:
                img += rng.normal(0, 3.0, img.shape)          # bad projections: deleted features, rejected frames
            img = img.astype(np.float32).astype(np.float64)  # vector<cv::Point2f>
            m = (t >= a - 1e-3) & (t <= b + 1e-3)
            fxy = np.c_[rng.integers(0, 346, 300), rng.integers(0, 260, 300)].astype(float)
            rc, out1, fid = ref.ref_rectify(t[m], x[m], y[m], p[m], a, b, 346, 260, fit, img, fxy)
            if rc < 0:
                continue
            P0, N0, _, _ = ref.event_frame(t, x, y, p, a, b)
            out0, ok0 = ref.rectify(P0, N0, img, 346, 260, fitCircle=fit)
            assert rc == int(ok0)
            np.testing.assert_array_equal(out0[:, 2] < 0, out1[:, 2] < 0)
            keep = out1[:, 2] >= 0
            np.testing.assert_array_equal(out0[keep], out1[keep])
            verdicts[rc] = verdicts.get(rc, 0) + 1
            if rc == 1:   # findCenter: nearest kept circle, accepted iff | ||p - c|| - r | < 5 px -> its landmark (board index)
                ids = np.nonzero(keep)[0]
                d2 = ((fxy[:, None, :] - out1[ids][None, :, :2]) ** 2).sum(-1)
                best = d2.argmin(1)
                acc = np.abs(np.sqrt(d2[np.arange(len(fxy)), best]) - out1[ids][best, 2]) < 5
                np.testing.assert_array_equal(fid, np.where(acc, ids[best], -1))
    assert verdicts.get(1, 0) >= 5 and verdicts.get(0, 0) >= 1


@pytest.mark.parametrize("with_gaps", [False, True])
def test_calib_spline_setup_association_and_assembly_vs_reference_source(ref, with_gaps):
    from eventcalib_b200 import synth, calib_problem
    import eventcalib_b200.build as b
    b.build()
    so = os.path.join(ROOT, "tests", "_build", "libfacade_host.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tests", "helpers", "facade_host.cpp"),
                           "-L" + os.path.join(ROOT, "eventcalib_b200"), "-lecb",
                           "-Wl,-rpath," + os.path.join(ROOT, "eventcalib_b200")])
    F = C.CDLL(so)
    dur = 0.3 if with_gaps else 0.2
    ev = synth.make_stream(int(1.2e6 * dur), 346, 260, t0=5.0, duration=dur, seed=11, return_truth=True)
    cam, traj, board = ev["camera"], ev["trajectory"], ev["board"]
    step = 5e-4
    pb = calib_problem.build_from_truth(cam, traj, board, 5.0, 5.0 + dur, step=step)
    kf_t, circ = pb["kf_t"], pb["circles"].copy()
    if with_gaps:   # a gap > 50 steps -> two segments; a 3-frame island between two gaps -> dropped (reduceMap :319-348)
        keep = ~(((kf_t > 5.12) & (kf_t < 5.15)) | ((kf_t >= 5.162) & (kf_t < 5.19)))
        kf_t, circ = kf_t[keep], circ[keep]
    rng = np.random.default_rng(0)
    circ[rng.uniform(size=circ.shape[:2]) < 0.05, 2] = -1.0        # features deleted by rectifyFeatures
    q, tw = traj.quat_xyzw(kf_t)
    cam9 = np.array([cam.f * 1.01, cam.f * 0.99, cam.cx + 0.5, cam.cy - 0.5, -0.33, -0.02, 0, 0, 0.5])
    r = ref.ref_calib_spline(ev["t"], ev["x"], ev["y"], ev["p"], kf_t, q, tw, circ, board.centres(), cam9, 346, 260, step, board.radius)
    assert r["n_splines"] == (2 if with_gaps else 1) and r["solve_calls"] == 1
    left = ~np.isnan(r["kf_pose"][:, 0])
    assert left.sum() == r["frames_left"] == (len(kf_t) - 3 if with_gaps else len(kf_t))
    # ---- spline set-up vs the façade (EventCalibSpline::segmentsFromKeyframes) ----
    K = len(kf_t)
    for w in range(r["n_splines"]):
        ncp = C.c_int()
        kn, rot, tr = np.zeros(K + 8), np.zeros((K, 4)), np.zeros((K, 3))
        ns = F.fh_segments(P(kf_t), P(np.ascontiguousarray(q)), P(np.ascontiguousarray(tw)), K, C.c_double(step), w, C.byref(ncp),
                           P(kn), P(rot), P(tr))
        n = ncp.value
        assert ns == r["n_splines"] and n == r["n_cp"][w]
        np.testing.assert_array_equal(kn[:n + 4], r["knots"][w])
        np.testing.assert_allclose(rot[:n], r["rot_cp"][w], rtol=0, atol=1e-13)
        np.testing.assert_allclose(tr[:n], r["trans_cp"][w], rtol=1e-13, atol=1e-13)
        assert r["ranges"][w, 0] == r["knots"][w][0] and r["ranges"][w, 1] == r["knots"][w][-1]
    if not with_gaps:   # cpNum = floor(T / (50 step)) with T extended by 3 steps on both sides (:64-85)
        assert r["n_cp"][0] == int(np.floor((kf_t[-1] - kf_t[0] + 6 * step) / (50 * step)))
    # ---- intrinsics_ (:94-105): K and the inverse of the radial part (k1, k2, k3 = distCoeffs(4), 0) ----
    np.testing.assert_array_equal(r["intrinsics"], np.r_[cam9[:4], ref.inverse_radial([cam9[4], cam9[5], cam9[8], 0.0])])
    # ---- association loop (:157-192) vs the oracle's, on the key frames left in the map ----
    cp = ref.CostProblem(r["n_cp"], r["knots"], radius=board.radius, huber=0.2 * board.radius)
    oe, oc = cp.associate(ev["t"], ev["x"], ev["y"], kf_t[left], circ[left], board.centres(), step)
    assert len(oe) == r["n_residuals"] > 10000
    np.testing.assert_array_equal(r["obs"], np.c_[ev["x"][oe], ev["y"][oe]])
    np.testing.assert_array_equal(r["lm"], board.centres()[oc])
    spl = np.searchsorted(r["ranges"][:, 1], ev["t"][oe], side="left")
    np.testing.assert_array_equal(r["spline"], spl)
    for i in range(0, len(oe), 53):
        sp, N = ref.basis(r["knots"][r["spline"][i]], float(ev["t"][oe[i]]))
        assert sp == r["span"][i]
        np.testing.assert_array_equal(N, r["basis"][i])
    # ---- problem assembly (:116-247) ----
    np.testing.assert_array_equal(r["first_cp"][:, 0], r["span"] - 3)     # rotation blocks span-3 .. span
    np.testing.assert_array_equal(r["first_cp"][:, 1], r["span"] - 3)     # translation blocks span-3 .. span
    assert r["param_blocks"] == r["quaternion_blocks"] == int(r["n_cp"].sum())   # every rotation control point: 4 -> 3
    assert abs(r["huber"] - 0.2 * board.radius) < 1e-15 and r["gradient_tolerance"] == 1e-10 and r["function_tolerance"] == 1e-10
    assert r["linear_solver"] == 1    # SPARSE_NORMAL_CHOLESKY
    # ---- updateMap (:253-317): key-frame poses = the splines at the stamps (no-op solve: the initial fit) ----
    F.fh_eval.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
    for k in np.nonzero(left)[0][::5]:
        w = int(np.searchsorted(r["ranges"][:, 1], kf_t[k], side="left"))
        q4, t3 = np.zeros(4), np.zeros(3)
        assert F.fh_eval(P(r["knots"][w]), int(r["n_cp"][w]), P(np.ascontiguousarray(r["rot_cp"][w])),
                         P(np.ascontiguousarray(r["trans_cp"][w])), 0, float(kf_t[k]), P(q4), P(t3)) == 1
        np.testing.assert_allclose(r["kf_pose"][k, 1:4], t3, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(r["kf_pose"][k, 4:8], q4, rtol=1e-12, atol=1e-12)


def _host_libs():
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
    CI.ci_calibrate.restype = C.c_double
    CI.ci_calibrate.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double] + [C.c_void_p] * 5
    CI.ci_solve_pnp.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double] + [C.c_void_p] * 4
    CI.ci_project.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4
    CI.ci_check_pose.argtypes = [C.c_double, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_double]
    return F, CI


def test_tracking_gate_and_check_pose_vs_reference_source(ref):
    """f-1: TrackingBase::process + EventCalibIni::track (row directions, median angle / duration < 5e-4 pi / step, reference
    frame by lower_bound of the time stamp) against the façade's TrackingGate, frames arriving out of order like from the
    reference's worker threads; f-4: EventCalibIni::checkPose against ecb::checkPose."""
    from eventcalib_b200 import synth
    from scipy.spatial.transform import Rotation as Rot
    F, CI = _host_libs()
    board, cam, step = synth.Board(), synth.Camera(), 5e-4
    rng = np.random.default_rng(0)
    mixed = 0
    for trial, amp in enumerate((1.0, 2.0, 3.0, 6.0)):
        traj = synth.Trajectory(5 + trial, board, 78.0, rot_amp=(0.1, 0.1, amp))
        ini = ref.RefIni(346, 260, step)
        g = C.c_void_p(F.fh_gate_new(9, 4, C.c_double(step)))
        ts = np.sort(rng.uniform(5.0, 5.6, 160))
        rng.shuffle(ts[40:])
        a1, a2 = [], []
        for tt in ts:
            R, tw = traj.pose(np.array([tt]))
            u, v = synth.project(cam, np.repeat(R, 36, 0), np.repeat(tw, 36, 0), board.centres())
            xy = np.ascontiguousarray(np.c_[u, v] + rng.normal(0, 0.2, (36, 2)))
            a1.append(ini.gate(float(tt), xy))
            a2.append(int(F.fh_gate_process(g, C.c_double(tt), P(xy), 36)))
        F.fh_gate_free(g)
        assert a1 == a2
        mixed += 0 < sum(a1) < len(a1)
    assert mixed >= 3
    accepted = 0
    for it in range(1500):
        q0, t0 = Rot.random(random_state=it).as_quat(), rng.normal(0, 30, 3)
        dt = rng.uniform(1e-3, 2e-2)
        ang = rng.uniform(0, 2.2) * 2 * 5e-4 * np.pi / step * dt
        dtr = rng.uniform(0, 2.2) * 2 * 0.25 / step * dt
        q1 = (Rot.from_rotvec(Rot.random(random_state=it + 7).apply([0, 0, 1]) * ang) * Rot.from_quat(q0)).as_quat()
        t1 = t0 + Rot.random(random_state=it + 9).apply([1, 0, 0]) * dtr
        a = ref.ref_check_pose(1.0, q0, t0, 1.0 + dt, q1, t1, step)
        b = CI.ci_check_pose(1.0, P(q0.copy()), P(t0.copy()), 1.0 + dt, P(q1.copy()), P(t1.copy()), step)
        assert a == b
        accepted += a
    assert 100 < accepted < 1400


@pytest.mark.parametrize("fit,n_use", [(0, 200), (1, 10)])
def test_cv_calibration_flow_vs_reference_source(ref, fit, n_use):
    """f-4: the reference's own loop (eventCameraCalib.cpp:49-62 per window, then EventCalibIni::cvCalibration :143-327) on raw
    events, with the product's calibrateCamera / PnP / projectPoints as the OpenCV hooks, against the batched-then-replayed
    logic the façade (opengv2::EventCalibIni::cvCalibration in include/ecb/event_calib.hpp) and the CLI use: frames per
    status, camera (exact), poses (1e-12: the reference goes through Eigen::Quaterniond(Rsw), the façade through R^T),
    rectified features (exact)."""
    from eventcalib_b200 import synth
    glib, grid_order = _grid()
    F, CI = _host_libs()
    step, W, H = 5e-4, 346, 260
    ev = synth.make_stream(400000, W, H, t0=5.0, duration=0.2, seed=1001, rot_amp=(0.35, 0.35, 0.3), orbit=True)
    t, x, y, p = ev["t"], ev["x"], ev["y"], ev["p"]
    wins = np.array([[a, a + 1.5e-3] for a in np.arange(5.0, 5.198, 4e-3)])
    ini = ref.RefIni(W, H, step, n_use=n_use, fitCircle=fit)
    ini.add_events(t, x, y, p)
    r = ini.run(wins)
    assert r["ok"] and r["calibrate_views"] == min(n_use, r["frames_before"])
    assert r["calibrate_flags"] == (1 << 17) | 2 | 4 | 8 | 2048 | 4096 | 8192   # CALIB_USE_LU | validate()'s flags (parameters.hpp:49-60)
    # ---- replay ----
    board = synth.Board()
    obj = np.ascontiguousarray(board.centres().astype(np.float32).astype(np.float64))   # cv::Point3f
    rthr = ref.radius_threshold(W, H, 9, 4, 1, 5.5, 1.75)
    g = C.c_void_p(F.fh_gate_new(9, 4, C.c_double(step)))
    st = np.zeros(len(wins), np.int32)
    frames = {}
    for w, (a, b) in enumerate(wins):
        P0, N0, _, _ = ref.event_frame(t, x, y, p, float(a), float(b))
        c = ref.extract(P0, N0, fitCircle=fit, Rthr=rthr)["cand"]
        if len(c) < 36:
            continue
        okg, order = grid_order(glib, c[:, 2:4].astype(np.float32).astype(np.float64))
        if not okg:
            continue
        f = np.ascontiguousarray(c[order][:, 2:5])
        ts = (a + b) / 2
        acc = F.fh_gate_process(g, C.c_double(ts), P(np.ascontiguousarray(f[:, :2])), 36)
        st[w] = 2 if acc else 1
        if acc:
            frames[ts] = (w, f, P0, N0)
    F.fh_gate_free(g)
    stamps = sorted(frames)
    use, stp = n_use, len(stamps) // n_use
    if stp == 0:
        use, stp = len(stamps), 1
    img = np.ascontiguousarray(np.array([frames[stamps[i * stp]][1][:, :2] for i in range(use)]).astype(np.float32).astype(np.float64))
    c9, rv, tv = np.zeros(9), np.zeros((use, 3)), np.zeros((use, 3))
    tot, pv = np.zeros(1), np.zeros(use, np.float32)
    CI.ci_calibrate(P(obj), 36, P(img), use, W, H, 7, 1.0, P(c9), P(rv), P(tv), P(tot), P(pv))
    np.testing.assert_array_equal(c9, r["cam9"])
    last, sk = None, 1.75 / np.sqrt(2)
    for s in stamps:
        w, f, P0, N0 = frames[s]
        im = np.ascontiguousarray(f[:, :2].astype(np.float32).astype(np.float64))
        r3, t3, inl, nin = np.zeros(3), np.zeros(3), np.zeros(36, np.int32), np.zeros(1, np.int32)
        CI.ci_solve_pnp(P(obj), 36, P(im), P(c9), 4.0, P(r3), P(t3), P(inl), P(nin))
        q, tw = np.zeros(4), np.zeros(3)
        CI.ci_body_pose(P(r3), P(t3), P(q), P(tw))
        if last is not None and not CI.ci_check_pose(last[0], P(last[1]), P(last[2]), s, P(q), P(tw), step):
            continue
        ip = np.zeros((36, 5, 2))
        for k in range(36):
            o5 = np.array([obj[k], obj[k] + [sk, sk, 0], obj[k] + [sk, -sk, 0], obj[k] + [-sk, -sk, 0], obj[k] + [-sk, sk, 0]])
            o5 = np.ascontiguousarray(o5.astype(np.float32).astype(np.float64))
            out = np.zeros((5, 2))
            CI.ci_project(P(o5), 5, P(r3), P(t3), P(c9), P(out))
            ip[k] = out
        o2, ok2 = ref.rectify(P0, N0, ip.astype(np.float32).astype(np.float64), W, H, fitCircle=fit)
        if not ok2:
            continue
        st[w] = 3
        np.testing.assert_allclose(r["pose"][w], np.r_[tw, q], rtol=0, atol=1e-11)
        keep = o2[:, 2] >= 0
        np.testing.assert_array_equal(r["feat"][w][:, 2] < 0, ~keep)
        np.testing.assert_array_equal(r["feat"][w][keep], o2[keep])
        last = (s, q.copy(), tw.copy())
    np.testing.assert_array_equal(st, r["status"])
    assert (st == 3).sum() == r["frames_after"] >= 10 and (st >= 2).sum() == r["frames_before"]
    # ---- the façade's own replay (opengv2::EventCalibIni::replayKeep): poses and rectify verdicts of ALL frames first (what
    # the batched GPU call delivers), then the loop — must keep exactly the frames the reference's sequential loop keeps ----
    n = len(stamps)
    qs, tws, verdict = np.zeros((n, 4)), np.zeros((n, 3)), np.zeros(n, np.int32)
    for i, s_ in enumerate(stamps):
        w, f, P0, N0 = frames[s_]
        im = np.ascontiguousarray(f[:, :2].astype(np.float32).astype(np.float64))
        r3, t3, inl, nin = np.zeros(3), np.zeros(3), np.zeros(36, np.int32), np.zeros(1, np.int32)
        CI.ci_solve_pnp(P(obj), 36, P(im), P(c9), 4.0, P(r3), P(t3), P(inl), P(nin))
        CI.ci_body_pose(P(r3), P(t3), P(qs[i]), P(tws[i]))
        ip = np.zeros((36, 5, 2))
        for k in range(36):
            o5 = np.array([obj[k], obj[k] + [sk, sk, 0], obj[k] + [sk, -sk, 0], obj[k] + [-sk, -sk, 0], obj[k] + [-sk, sk, 0]])
            o5 = np.ascontiguousarray(o5.astype(np.float32).astype(np.float64))
            out = np.zeros((5, 2))
            CI.ci_project(P(o5), 5, P(r3), P(t3), P(c9), P(out))
            ip[k] = out
        verdict[i] = int(ref.rectify(P0, N0, ip.astype(np.float32).astype(np.float64), W, H, fitCircle=fit)[1])
    keep, counters = np.zeros(n, np.int8), np.zeros(2, np.int32)
    F.fh_replay_keep(P(np.ascontiguousarray(np.array(stamps))), P(qs), P(tws), P(verdict), n, C.c_double(step), P(keep), P(counters))
    np.testing.assert_array_equal(keep.astype(bool), np.array([r["status"][frames[s_][0]] == 3 for s_ in stamps]))
    assert counters[0] + counters[1] == n - keep.sum()
