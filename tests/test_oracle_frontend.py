"""CPU checks of the front-end oracle: hashing / unordered_set order known answers (SURVEY.md Appendix E and
tests/golden/uset_order.npz), EventFrame window semantics, fitCircle against an independent least-squares solve,
and the synthetic stream / frontend sanity."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(__file__), "golden", "uset_order.npz")


def test_hash_known_answers(oracle_mod):
    lib = oracle_mod.port()
    assert lib.orc_hash_double(0.0) == 0
    assert lib.orc_hash_double(1.0) == 8386164645967068059
    assert lib.orc_hash_double(2.0) == 6369015886390043782
    assert lib.orc_hash_double(4.0) == 397442141738980197
    assert lib.orc_hash_p2(0.0, 0.0) == 175247769566
    assert lib.orc_hash_p2(1.0, 0.0) == 4747032605431675226
    assert lib.orc_hash_p2(0.0, 1.0) == 8386164817454775227
    assert lib.orc_hash_p2(3.0, 4.0) == 2475419552798222426


def test_unordered_set_order(oracle_mod):
    pix = np.array([[x, y] for y in range(3) for x in range(5)], float)
    want = [(4, 2), (3, 2), (2, 2), (1, 0), (4, 1), (2, 0), (3, 1), (3, 0), (4, 0), (0, 1), (0, 0), (1, 1), (2, 1), (0, 2), (1, 2)]
    assert [tuple(map(int, p)) for p in oracle_mod.uset_order(pix)] == want
    g = np.load(GOLD)
    for j in range(int(g["n_cases"])):
        assert np.array_equal(oracle_mod.uset_order(g["in_%d" % j].astype(float)).astype(np.int16), g["out_%d" % j])


def test_event_frame_semantics(oracle_mod):
    t = np.array([1.0, 1.0, 2.0, 2.5, 3.0, 3.0, 4.0])
    x = np.array([5, 6, 5, 7, 8, 5, 9], float)
    y = np.zeros(7)
    p = np.array([1, 1, 0, 1, 0, 1, 1], np.uint8)
    P, N, lo, hi = oracle_mod.event_frame(t, x, y, p, 1.0, 3.0)     # CLOSED on both ends
    assert (lo, hi) == (0, 6)
    # pixel 5 has both polarities -> cancelled from both sets (EventFrame.cpp:23-32); duplicates collapse
    assert sorted(P[:, 0].tolist()) == [6.0, 7.0] and sorted(N[:, 0].tolist()) == [8.0]
    P, N, lo, hi = oracle_mod.event_frame(t, x, y, p, 1.5, 2.9)
    assert (lo, hi) == (2, 4)
    P, N, lo, hi = oracle_mod.event_frame(t, x, y, p, 10.0, 11.0)
    assert lo == hi and len(P) == 0 and len(N) == 0


def test_fit_circle_against_independent_lstsq(oracle_mod):
    rng = np.random.default_rng(3)
    for _ in range(50):
        c, r = rng.uniform(30, 200, 2), rng.uniform(5, 16)
        th = rng.uniform(0, 2 * np.pi, int(rng.integers(10, 150)))
        pts = np.rint(np.stack([c[0] + r * np.cos(th), c[1] + r * np.sin(th)], 1) + rng.normal(0, 0.5, (len(th), 2)))
        # Kasa: minimise sum (x^2 + y^2 - 2 a x - 2 b y - c)^2
        A = np.stack([2 * pts[:, 0], 2 * pts[:, 1], np.ones(len(pts))], 1)
        sol, *_ = np.linalg.lstsq(A, (pts ** 2).sum(1), rcond=None)
        want = np.array([sol[0], sol[1], np.sqrt(sol[0] ** 2 + sol[1] ** 2 + sol[2])])
        got = oracle_mod.fit_circle(pts[: len(pts) // 3], pts[len(pts) // 3:])
        np.testing.assert_allclose(got, want, rtol=1e-8)


def test_radius_threshold(oracle_mod):
    # 346x260, 9x4 asymmetric, 5.5 / 1.75 cm  
    assert abs(oracle_mod.radius_threshold(346, 260, 9, 4, 1, 5.5, 1.75) - 15.511363636363637) < 1e-12
    import eventcalib_b200 as ecb
    for args in [(346, 260, 9, 4, True, 5.5, 1.75), (640, 480, 7, 5, False, 3.0, 1.0), (1280, 720, 9, 4, True, 5.5, 1.75)]:
        assert ecb.radius_threshold(*args) == oracle_mod.radius_threshold(*[float(a) if i < 2 else a for i, a in enumerate(args)])


def test_frontend_finds_the_board(oracle_mod):
    from eventcalib_b200 import synth
    ev = synth.make_stream(30000, 346, 260, t0=5.0, duration=0.015, seed=11)
    assert np.all(np.diff(ev["t"]) > 0) and ev["x"].max() <= 345 and ev["y"].max() <= 259
    win = synth.tiling_windows(5.0, 5.015, 1.5e-3)
    rthr = oracle_mod.radius_threshold(346, 260, 9, 4, 1, 5.5, 1.75)
    for ref in (True, False):
        tot, nev, per = oracle_mod.frontend_windows(ev["t"], ev["x"], ev["y"], ev["p"], win, Rthr=rthr, threads=2, ref=ref)
        assert nev == len(ev["t"]) and per.min() >= 30 and per.max() <= 36
    # both fitCircle modes give the same pairing on clean data
    P, N, lo, hi = oracle_mod.event_frame(ev["t"], ev["x"], ev["y"], ev["p"], win[0, 0], win[0, 1])
    a = oracle_mod.extract(P, N, Rthr=rthr, fitCircle=0)
    b = oracle_mod.extract(P, N, Rthr=rthr, fitCircle=1)
    assert a["enough"] and b["enough"] and len(a["cand"]) >= 30
    # records round-trip through the reference's binary format
    import tempfile
    with tempfile.NamedTemporaryFile(suffix=".bin") as f:
        synth.write_bin(f.name, ev)
        assert os.path.getsize(f.name) == 25 * len(ev["t"])
        back = synth.read_bin(f.name)
        assert all(np.array_equal(back[k], ev[k]) for k in "txyp")


def test_rectify_features_oracle_sanity(oracle_mod):
    """rectifyFeatures restatement (CirclesEventFrame.cpp:417-609): with projections through the ground-truth camera the
    rectified circles land on the projected centres and the frame is kept; shifted projections delete features."""
    from eventcalib_b200 import synth
    ev = synth.make_stream(6000, 346, 260, t0=5.0, duration=0.003, seed=5, return_truth=True)
    P, N, _, _ = oracle_mod.event_frame(ev["t"], ev["x"], ev["y"], ev["p"], 5.0, 5.0015)
    board, cam, traj = ev["board"], ev["camera"], ev["trajectory"]
    c = board.centres()
    k = board.radius / np.sqrt(2)
    offs = np.array([[0, 0, 0], [k, k, 0], [k, -k, 0], [-k, -k, 0], [-k, k, 0]])
    X = (c[:, None, :] + offs[None, :, :]).reshape(-1, 3)
    R, tw = traj.pose(np.full(len(X), 5.00075))
    u, v = synth.project(cam, R, tw, X)
    img = np.stack([u, v], 1).reshape(36, 5, 2)
    out, ok = oracle_mod.rectify(P, N, img, 346, 260)
    alive = out[:, 2] >= 0
    assert ok == 1 and alive.sum() >= 32
    assert np.abs(out[alive, :2] - img[alive, 0]).max() < 1.5
    out2, ok2 = oracle_mod.rectify(P, N, img + 9.0, 346, 260)
    assert ok2 == 0 and (out2[:, 2] < 0).sum() > 20
