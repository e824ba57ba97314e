"""The CUDA path (through the C ABI) against tests/golden/reference_source.npz — outputs of the reference's OWN sources compiled
in place (tests/golden/make_reference_source_golden.py): pixel lists in the reference's set order, candidate centres as the
reference hands them to findCirclesGrid (cv::Point2f, exact), feature centres / radii (1e-9 relative, north star), and the
residual list of the reference's association loop (exact)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "reference_source.npz"))


@pytest.mark.parametrize("fit_circle", [0, 1])
def test_frontend_vs_reference_source_golden(ctx, fit_circle):
    import eventcalib_b200 as ecb
    from eventcalib_b200 import synth
    from test_circles_grid import _lib as grid_lib, _order as grid_order
    glib = grid_lib()
    ev = dict(t=G["ev_t"], x=G["ev_x"].astype(np.float64), y=G["ev_y"].astype(np.float64), p=G["ev_p"])
    ctx.set_sensor(346, 260)
    assert ctx.load_events(synth.to_records(ev)) == len(ev["t"])
    rthr = ecb.radius_threshold(346, 260, 9, 4, True, 5.5, 1.75)
    assert rthr == float(G["rthr"])
    prm = ecb.default_params(fit_circle=fit_circle, radius_threshold=rthr, order_mode=1, median_mode=1)
    win = np.ascontiguousarray(G["windows"], np.float64)
    ctx.frontend_run(win, prm)
    summ = ctx.summary()
    pts = [ctx.points(0), ctx.points(1)]
    cand = ctx.candidates(64)
    found = 0
    for i in range(len(win)):
        s = summ[i]
        for pol, key in ((0, "frame_neg_%d" % i), (1, "frame_pos_%d" % i)):
            o, k = int(s["point_offset"][pol]), int(s["n_points"][pol])
            np.testing.assert_array_equal(pts[pol][0][o:o + k], G[key].astype(np.float64))   # EventFrame's own order
        n = int(s["n_candidates"])
        if not bool(G["extract_reached_%d_%d" % (i, fit_circle)]):
            assert n == 0
            continue
        g = cand[i, :n]
        np.testing.assert_array_equal(g[:, 2:4].astype(np.float32), G["extract_cand_%d_%d" % (i, fit_circle)])
        ok, order = grid_order(glib, g[:, 2:4].astype(np.float32).astype(np.float64))
        assert ok == bool(G["extract_found_%d_%d" % (i, fit_circle)])
        if ok:
            found += 1
            np.testing.assert_allclose(g[order][:, 2:5], G["extract_features_%d_%d" % (i, fit_circle)], rtol=1e-9, atol=0)
    assert found >= 1


def test_association_vs_reference_source_golden(ctx):
    from eventcalib_b200 import synth
    board = synth.Board()
    ev = dict(t=G["spline_ev_t"], x=G["spline_ev_x"].astype(np.float64), y=G["spline_ev_y"].astype(np.float64), p=G["spline_ev_p"])
    ctx.set_sensor(346, 260)
    assert ctx.load_events(synth.to_records(ev)) == len(ev["t"])
    n_cp = G["spline_n_cp"]
    knots = [G["spline_knots_%d" % s] for s in range(len(n_cp))]
    ctx.cost_setup(n_cp, knots, board.radius, 0.2 * board.radius)
    left = ~np.isnan(G["spline_kf_pose"][:, 0])
    n = ctx.cost_associate(G["spline_kf_t"][left], G["spline_circ"][left], board.centres(), 5e-4)
    assert n == len(G["spline_span"])
    oe, oc = ctx.cost_association()
    np.testing.assert_array_equal(np.c_[ev["x"][oe], ev["y"][oe]], G["spline_obs"].astype(np.float64))
    np.testing.assert_array_equal(oc, G["spline_lm_idx"].astype(np.int32))


def test_so3_variant_vs_reference_source_golden(ctx):
    """a11 on the GPU (k_normal_eq<SO3>, k_cost<SO3>, association) against the reference's own CalibReprojectionError_SO3 /
    BsplineSO3::derBasisFuns / LocalParameterizationSO3 compiled in place: residual list exact; J^T J, J^T r and the cost of the
    golden problem within 1e-9 relative of the ones assembled from the reference functor's Jet<37> rows."""
    from eventcalib_b200 import synth
    from test_golden_reference_source import so3_golden_normal_equations
    ev = dict(t=G["so3_ev_t"], x=G["so3_ev_x"].astype(np.float64), y=G["so3_ev_y"].astype(np.float64), p=G["so3_ev_p"])
    Q, T, intr = G["so3_rot_cp"], G["so3_trans_cp"], G["so3_intrinsics"]
    ctx.set_sensor(346, 260)
    ctx.load_events(synth.to_records(ev))
    ctx.cost_setup([len(Q)], [G["so3_knots"]], 1.75, 0.35)
    ctx.cost_set_rotation_model(1)
    try:
        n = ctx.cost_associate(G["so3_kf_t"], G["so3_circles"], G["so3_landmarks"], float(G["so3_step"]))
        assert n == len(G["so3_r"])
        oe, oc = ctx.cost_association()
        np.testing.assert_array_equal(oe, G["so3_event"])
        np.testing.assert_array_equal(oc, G["so3_circle"].astype(np.int32))
        c, H, g = ctx.cost_normal_eq(intr, Q, T)
        c2 = ctx.cost_eval(intr, Q, T)
    finally:
        ctx.cost_set_rotation_model(0)
    Hg, gg, cg = so3_golden_normal_equations()
    assert abs(c - cg) <= 1e-9 * cg and abs(c2 - cg) <= 1e-9 * cg
    assert np.abs(H - Hg).max() <= 1e-9 * np.abs(Hg).max()
    assert np.abs(g - gg).max() <= 1e-9 * np.abs(gg).max()
