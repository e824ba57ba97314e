"""Parity of the CUDA cost evaluation (association, residual/Jacobian, per-span normal equations, cost) with
the dual-number oracle.  Tolerance: 1e-9 relative (north star); measured agreement is ~1e-12."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
RTOL = 1e-9


@pytest.fixture(scope="module")
def problem(ctx, oracle_mod):
    from eventcalib_b200 import synth, calib_problem
    ev = synth.make_stream(300000, 346, 260, t0=5.0, duration=0.3, seed=1004, return_truth=True)
    pb = calib_problem.build(ev, seed=3)
    ctx.set_sensor(346, 260)
    ctx.load_events(synth.to_records(ev))
    ctx.cost_setup([pb["n_cp"]], [pb["knots"]], pb["radius"], pb["huber"])
    n = ctx.cost_associate(pb["kf_t"], pb["circles"], pb["landmarks"], pb["step"])
    P = oracle_mod.CostProblem([pb["n_cp"]], [pb["knots"]], pb["radius"], pb["huber"])
    oe, oc = P.associate(ev["t"], ev["x"], ev["y"], pb["kf_t"], pb["circles"], pb["landmarks"], pb["step"])
    return ev, pb, P, n, oe, oc


def test_association_exact(ctx, problem):
    ev, pb, P, n, oe, oc = problem
    assert n == len(oe) and n > 0.8 * len(ev["t"])
    ge, gc = ctx.cost_association()
    assert np.array_equal(ge, oe)
    assert np.array_equal(gc, oc)


def test_cost_and_normal_equations(ctx, problem):
    ev, pb, P, n, oe, oc = problem
    x = (pb["intrinsics"], pb["rot_cp"], pb["trans_cp"])
    c_ref, H_ref, g_ref = P.normal_eq(*x)
    assert abs(ctx.cost_eval(*x) - c_ref) <= RTOL * c_ref
    c, H, g = ctx.cost_normal_eq(*x)
    assert abs(c - c_ref) <= RTOL * c_ref
    assert H.shape == H_ref.shape
    for s in range(H.shape[0]):
        np.testing.assert_allclose(H[s], H_ref[s], rtol=0, atol=RTOL * np.abs(H_ref[s]).max())
        np.testing.assert_allclose(g[s], g_ref[s], rtol=0, atol=RTOL * np.abs(g_ref[s]).max())
        assert np.array_equal(H[s], H[s].T)
    # measured agreement is far tighter than the bar
    assert np.abs(H - H_ref).max() / np.abs(H_ref).max() < 1e-11


def test_association_from_device_tables(ctx, problem):
    """ecb_cost_associate_device: the key-frame tables already in device memory give the same residual blocks and the same
    normal equations, bit for bit, as the upload from host arrays (test_association_exact pins those to the oracle)"""
    import torch
    ev, pb, P, n, oe, oc = problem
    x = (pb["intrinsics"], pb["rot_cp"], pb["trans_cp"])
    ref = ctx.cost_normal_eq(*x)
    d_t = torch.from_numpy(np.ascontiguousarray(pb["kf_t"], np.float64)).cuda()
    d_c = torch.from_numpy(np.ascontiguousarray(pb["circles"], np.float64)).cuda()
    d_l = torch.from_numpy(np.ascontiguousarray(pb["landmarks"], np.float64)).cuda()
    torch.cuda.synchronize()
    n2 = ctx.cost_associate_device(d_t.data_ptr(), d_c.data_ptr(), len(pb["kf_t"]), pb["circles"].shape[1], d_l.data_ptr(), pb["step"])
    assert n2 == n
    ge, gc = ctx.cost_association()
    assert np.array_equal(ge, oe) and np.array_equal(gc, oc)
    del d_t, d_c   # the tables are read during the call only; the landmark table was copied
    got = ctx.cost_normal_eq(*x)
    assert got[0] == ref[0] and np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2])


def test_deterministic(ctx, problem):
    ev, pb, P, n, oe, oc = problem
    x = (pb["intrinsics"], pb["rot_cp"], pb["trans_cp"])
    a = ctx.cost_normal_eq(*x)
    b = ctx.cost_normal_eq(*x)
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert ctx.cost_eval(*x) == ctx.cost_eval(*x)


def test_explicit_residuals_multi_spline_and_huber(ctx, oracle_mod):
    # two spline segments, non-integer observations, residuals far outside the Huber band
    from eventcalib_b200 import synth, spline
    rng = np.random.default_rng(9)
    cam, board = synth.Camera(), synth.Board()
    traj = synth.Trajectory(3, board, 78.0)
    knots, ncp, rot, trans, obs, lm, tt, sp = [], [], [], [], [], [], [], []
    for s, (a, b, n_cp) in enumerate([(1.0, 1.4, 7), (2.0, 2.25, 5)]):
        us = np.linspace(a, b, 40)
        kn = spline.knot_vector(us, n_cp)
        q, tw = traj.quat_xyzw(us)
        rot.append(spline.fit_control_points(kn, us, q, n_cp))
        trans.append(spline.fit_control_points(kn, us, tw, n_cp))
        knots.append(kn)
        ncp.append(n_cp)
        t = np.sort(rng.uniform(a, b, 5000))
        t[0], t[-1] = a, b  # both clamped ends, incl. the u == last-knot special case
        tt.append(t)
        sp.append(np.full(len(t), s))
        obs.append(np.stack([rng.uniform(10, 330, len(t)), rng.uniform(10, 250, len(t))], 1))
        lm.append(board.centres()[rng.integers(0, 36, len(t))])
    rot, trans = np.concatenate(rot), np.concatenate(trans)
    obs, lm, tt, sp = map(np.concatenate, (obs, lm, tt, sp))
    ctx.cost_setup(ncp, knots, 1.75, 0.35)
    ctx.cost_set_residuals(obs, lm, tt, sp)
    P = oracle_mod.CostProblem(ncp, knots, 1.75, 0.35)
    P.set_residuals(obs, lm, tt, sp)
    intr = cam.intrinsics()
    c_ref, H_ref, g_ref = P.normal_eq(intr, rot, trans)
    c, H, g = ctx.cost_normal_eq(intr, rot, trans)
    assert abs(c - c_ref) <= RTOL * c_ref
    assert np.abs(H - H_ref).max() <= RTOL * np.abs(H_ref).max()
    assert np.abs(g - g_ref).max() <= RTOL * np.abs(g_ref).max()
    assert abs(ctx.cost_eval(intr, rot, trans) - c_ref) <= RTOL * c_ref


def test_rejects_unordered_residuals(ctx):
    import eventcalib_b200 as ecb
    from eventcalib_b200 import spline
    kn = spline.knot_vector(np.linspace(0, 1, 20), 6)
    ctx.cost_setup([6], [kn])
    with pytest.raises(ecb.EcbError):
        ctx.cost_set_residuals(np.zeros((2, 2)), np.zeros((2, 3)), np.array([0.5, 0.2]), np.zeros(2))
    with pytest.raises(ecb.EcbError):
        ctx.cost_set_residuals(np.zeros((1, 2)), np.zeros((1, 3)), np.array([1.5]), np.zeros(1))


def test_so3_variant_normal_equations(ctx, oracle_mod, problem):
    """a11: CalibReprojectionError_SO3 + LocalParameterizationSO3 (EventCalibSpline.hpp:65-156) — cost, J^T J, J^T r of the
    same residual set against the oracle's Jet<37> evaluation, 1e-9 relative"""
    ev, pb, P, n, oe, oc = problem
    Ps = oracle_mod.CostProblem([pb["n_cp"]], [pb["knots"]], pb["radius"], pb["huber"], so3=True)
    Ps.associate(ev["t"], ev["x"], ev["y"], pb["kf_t"], pb["circles"], pb["landmarks"], pb["step"])
    rot = pb["rot_cp"].reshape(-1, 4)
    rot = rot / np.linalg.norm(rot, axis=1, keepdims=True)   # Sophus::SO3d control points are unit quaternions
    x = (pb["intrinsics"], rot, pb["trans_cp"])
    from eventcalib_b200 import synth
    ctx.set_sensor(346, 260)                 # earlier tests re-used the context for other residual sets
    ctx.load_events(synth.to_records(ev))
    ctx.cost_setup([pb["n_cp"]], [pb["knots"]], pb["radius"], pb["huber"])
    assert ctx.cost_associate(pb["kf_t"], pb["circles"], pb["landmarks"], pb["step"]) == Ps.n_residuals
    ctx.cost_set_rotation_model(1)
    try:
        c_ref, H_ref, g_ref = Ps.normal_eq(*x)
        c, H, g = ctx.cost_normal_eq(*x)
        assert abs(ctx.cost_eval(*x) - c_ref) <= RTOL * c_ref
        assert abs(c - c_ref) <= RTOL * c_ref
        for s in range(H.shape[0]):
            np.testing.assert_allclose(H[s], H_ref[s], rtol=0, atol=RTOL * np.abs(H_ref[s]).max())
            np.testing.assert_allclose(g[s], g_ref[s], rtol=0, atol=RTOL * np.abs(g_ref[s]).max())
        # a different function from the quaternion-spline variant
        c_q, _, _ = P.normal_eq(*x)
        assert abs(c_q - c_ref) > 1e-6 * c_ref
    finally:
        ctx.cost_set_rotation_model(0)


def test_fused_exchange_two_ranks_on_one_gpu(oracle_mod):
    """The fused reduce + exchange (ecb_cost_normal_eq_exchange) with two contexts acting as two ranks on one GPU:
    each holds one half of the events, both write their slots into both receive buffers; every rank's result equals
    the sum of the two single-rank normal equations bit for bit, over several epochs (both parities)."""
    import torch
    import eventcalib_b200 as ecb
    from eventcalib_b200 import synth, calib_problem
    ev = synth.make_stream(200000, 346, 260, t0=5.0, duration=0.3, seed=1004, return_truth=True)
    pb = calib_problem.build(ev, seed=3)
    n = len(ev["t"])
    cut = n // 2 + 777          # the halves overlap in one knot span
    rec = synth.to_records(ev)
    ranks = []
    for r, (lo, hi) in enumerate(((0, cut), (cut, n))):
        c = ecb.Context(0)
        c.set_sensor(346, 260)
        c.load_events(rec[lo:hi])
        c.cost_setup([pb["n_cp"]], [pb["knots"]], pb["radius"], pb["huber"])
        assert c.cost_associate(pb["kf_t"], pb["circles"], pb["landmarks"], pb["step"]) > 1000
        ranks.append(c)
    lay = ranks[0].cost_layout()
    nbytes = ranks[0].exchange_buffer_bytes(2)
    bufs = [c.device_alloc(nbytes) for c in ranks]
    outs = [torch.zeros(lay["out_doubles"], dtype=torch.float64, device="cuda") for _ in ranks]
    try:
        rng = np.random.default_rng(0)
        for epoch in (1, 2, 3):
            x = (pb["intrinsics"] * (1 + 1e-3 * rng.normal(size=9)), pb["rot_cp"], pb["trans_cp"])
            single = []
            for c in ranks:
                cst, H, g = c.cost_normal_eq(*x)
                single.append((cst, H, g))
            # two ranks on ONE device: all send sides first, then the (spinning) receive sides — see the header
            for r in (0, 1):
                ranks[r].cost_normal_eq_exchange(*x, r, bufs, epoch, outs[r].data_ptr(), want_cost=False, phases=1)
            c0 = ranks[0].cost_normal_eq_exchange(*x, 0, bufs, epoch, outs[0].data_ptr(), want_cost=True, phases=2)
            c1 = ranks[1].cost_normal_eq_exchange(*x, 1, bufs, epoch, outs[1].data_ptr(), want_cost=True, phases=2)
            assert c0 == c1
            a, b = outs[0].cpu().numpy(), outs[1].cpu().numpy()
            assert np.array_equal(a, b)
            ns = lay["total_spans"]
            blk = a[:ns * 1122].reshape(ns, 1122)
            H = blk[:, :1089].reshape(ns, 33, 33)
            g = blk[:, 1089:]
            assert np.array_equal(H, single[0][1] + single[1][1]) and np.array_equal(g, single[0][2] + single[1][2])
            assert a[ns * 1122] == single[0][0] + single[1][0] == c1
    finally:
        for c, p in zip(ranks, bufs):
            c.synchronize()
            c.device_free(p)
            c.close()
