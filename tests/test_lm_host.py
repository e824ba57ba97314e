"""CPU tests of the host-side LM state machine (ecb_lm_*, eventcalib_b200/csrc/ecb_lm.cu): driven by the ORACLE's
normal equations it must follow the same iterate sequence as the independent dense-numpy restatement
(tests/lm_oracle.py), i.e. the banded-arrowhead Cholesky + trust-region logic are checked without a GPU."""
import numpy as np
import pytest

import lm_oracle


def _packed(c, H, g):
    ns = H.shape[0]
    out = np.zeros(ns * 1122 + 2)
    blk = out[:ns * 1122].reshape(ns, 1122)
    blk[:, :1089] = H.reshape(ns, 1089)
    blk[:, 1089:] = g
    out[ns * 1122] = c
    return out


@pytest.fixture(scope="module")
def small_problem(oracle_mod):
    from eventcalib_b200 import synth, calib_problem
    ev = synth.make_stream(60000, 346, 260, t0=5.0, duration=0.4, seed=77, return_truth=True, rot_amp=(0.35, 0.35, 0.25),
                           dist=92.0)
    pb = calib_problem.build(ev, seed=1, intr_noise=0.01)
    P = oracle_mod.CostProblem([pb["n_cp"]], [pb["knots"]], pb["radius"], pb["huber"])
    P.associate(ev["t"], ev["x"], ev["y"], pb["kf_t"], pb["circles"], pb["landmarks"], pb["step"])
    return pb, P


def _run_host_lm(pb, P, **opt):
    import eventcalib_b200 as ecb
    lm = ecb.LmState([pb["n_cp"]], ecb.lm_options(**opt))
    x = (pb["intrinsics"], pb["rot_cp"], pb["trans_cp"])
    st = lm.begin(*x, _packed(*P.normal_eq(*x)))
    while st == 0:
        st, ci, cr, ct = lm.propose()
        if st != 0:
            break
        fb = lm.feedback(P.cost(ci, cr, ct))
        if fb == 1:
            st = lm.update(_packed(*P.normal_eq(ci, cr, ct)))
        elif fb == 0:
            st = 0
        else:
            st = fb
    return lm


def test_lm_matches_dense_restatement(small_problem):
    pb, P = small_problem
    assert pb["n_cp"] >= 6
    lm = _run_host_lm(pb, P, max_iterations=12)
    i1, r1, t1, summ = lm.state()
    tr1 = lm.trace()
    i2, r2, t2, tr2, term = lm_oracle.solve(P, [pb["n_cp"]], pb["intrinsics"], pb["rot_cp"], pb["trans_cp"], max_iterations=12)
    tr2 = np.array(tr2)
    assert len(tr1) == len(tr2)
    np.testing.assert_array_equal(tr1[:, 3], tr2[:, 3])           # same accept / reject sequence
    np.testing.assert_allclose(tr1[:, 0], tr2[:, 0], rtol=1e-9)    # same cost trajectory
    np.testing.assert_allclose(tr1[:, 2], tr2[:, 2], rtol=1e-6)    # same trust-region radii
    np.testing.assert_allclose(i1, i2, rtol=1e-9)
    np.testing.assert_allclose(r1, r2, rtol=0, atol=1e-9)
    np.testing.assert_allclose(t1, t2, rtol=1e-9, atol=1e-9)
    assert summ["final_cost"] < 0.7 * summ["initial_cost"]


def test_lm_recovers_intrinsics(small_problem):
    pb, P = small_problem
    lm = _run_host_lm(pb, P, max_iterations=50)
    i1, _, _, summ = lm.state()
    err0 = np.abs(pb["intrinsics"][:2] / pb["truth_intrinsics"][:2] - 1).max()
    err1 = np.abs(i1[:2] / pb["truth_intrinsics"][:2] - 1).max()
    assert err1 < 0.6 * err0          # fx, fy move towards the ground truth
    assert summ["termination"] in (2, 3, 4, 5)


def test_lm_fixed_iterations_runs_exactly(small_problem):
    pb, P = small_problem
    lm = _run_host_lm(pb, P, max_iterations=7, fixed_iterations=1)
    assert lm.state()[3]["iterations"] == 7
