"""End-to-end parity of the GPU-driven LM (ecb_calibrate: CUDA normal equations + host banded solve) with the
independent dense restatement driven by the dual-number oracle (tests/lm_oracle.py): same accept/reject sequence,
cost trajectory and final intrinsics / control points within 1e-9 relative."""
import numpy as np
import pytest

import lm_oracle

pytestmark = pytest.mark.gpu


def test_calibrate_matches_oracle_lm(ctx, oracle_mod):
    import eventcalib_b200 as ecb
    from eventcalib_b200 import synth, calib_problem
    ev = synth.make_stream(150000, 346, 260, t0=5.0, duration=0.6, seed=1004, return_truth=True, rot_amp=(0.35, 0.35, 0.25),
                           dist=92.0)
    pb = calib_problem.build(ev, seed=2, intr_noise=0.02)
    ctx.set_sensor(346, 260)
    ctx.load_events(synth.to_records(ev))
    ctx.cost_setup([pb["n_cp"]], [pb["knots"]], pb["radius"], pb["huber"])
    n = ctx.cost_associate(pb["kf_t"], pb["circles"], pb["landmarks"], pb["step"])
    P = oracle_mod.CostProblem([pb["n_cp"]], [pb["knots"]], pb["radius"], pb["huber"])
    P.associate(ev["t"], ev["x"], ev["y"], pb["kf_t"], pb["circles"], pb["landmarks"], pb["step"])
    assert n == P.n_residuals
    K = 20
    i1, r1, t1, summ, tr1 = ctx.calibrate([pb["n_cp"]], pb["intrinsics"], pb["rot_cp"], pb["trans_cp"],
                                          ecb.lm_options(max_iterations=K))
    i2, r2, t2, tr2, term = lm_oracle.solve(P, [pb["n_cp"]], pb["intrinsics"], pb["rot_cp"], pb["trans_cp"], max_iterations=K)
    tr2 = np.array(tr2)
    assert len(tr1) == len(tr2)
    np.testing.assert_array_equal(tr1[:, 3], tr2[:, 3])
    np.testing.assert_allclose(tr1[:, 0], tr2[:, 0], rtol=1e-9)
    np.testing.assert_allclose(i1, i2, rtol=1e-9)
    np.testing.assert_allclose(r1.reshape(-1, 4), r2, rtol=0, atol=1e-9)
    np.testing.assert_allclose(t1.reshape(-1, 3), t2, rtol=1e-9, atol=1e-9)
    assert summ["final_cost"] < summ["initial_cost"]
    # fx, fy move towards the ground truth
    assert abs(i1[0] / pb["truth_intrinsics"][0] - 1) < abs(pb["intrinsics"][0] / pb["truth_intrinsics"][0] - 1)


def test_calibrate_so3_variant_matches_oracle_lm(ctx, oracle_mod):
    """useSO3: 1 — SO(3) spline residuals, x+ = x * exp(delta) in the LM step"""
    import eventcalib_b200 as ecb
    from eventcalib_b200 import synth, calib_problem
    ev = synth.make_stream(100000, 346, 260, t0=5.0, duration=0.4, seed=1004, return_truth=True, rot_amp=(0.35, 0.35, 0.25),
                           dist=92.0)
    pb = calib_problem.build(ev, seed=2, intr_noise=0.02)
    rot0 = pb["rot_cp"].reshape(-1, 4)
    rot0 = rot0 / np.linalg.norm(rot0, axis=1, keepdims=True)
    ctx.set_sensor(346, 260)
    ctx.load_events(synth.to_records(ev))
    ctx.cost_setup([pb["n_cp"]], [pb["knots"]], pb["radius"], pb["huber"])
    ctx.cost_set_rotation_model(1)
    try:
        n = ctx.cost_associate(pb["kf_t"], pb["circles"], pb["landmarks"], pb["step"])
        P = oracle_mod.CostProblem([pb["n_cp"]], [pb["knots"]], pb["radius"], pb["huber"], so3=True)
        P.associate(ev["t"], ev["x"], ev["y"], pb["kf_t"], pb["circles"], pb["landmarks"], pb["step"])
        assert n == P.n_residuals
        K = 12
        i1, r1, t1, summ, tr1 = ctx.calibrate([pb["n_cp"]], pb["intrinsics"], rot0, pb["trans_cp"],
                                              ecb.lm_options(max_iterations=K, rotation_model=1))
        i2, r2, t2, tr2, term = lm_oracle.solve(P, [pb["n_cp"]], pb["intrinsics"], rot0, pb["trans_cp"], max_iterations=K, so3=True)
        tr2 = np.array(tr2)
        assert len(tr1) == len(tr2)
        np.testing.assert_array_equal(tr1[:, 3], tr2[:, 3])
        np.testing.assert_allclose(tr1[:, 0], tr2[:, 0], rtol=1e-9)
        np.testing.assert_allclose(i1, i2, rtol=1e-9)
        np.testing.assert_allclose(r1.reshape(-1, 4), r2, rtol=0, atol=1e-9)
        np.testing.assert_allclose(t1.reshape(-1, 3), t2, rtol=1e-9, atol=1e-9)
        assert summ["final_cost"] < summ["initial_cost"]
    finally:
        ctx.cost_set_rotation_model(0)
