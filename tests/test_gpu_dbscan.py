"""Parity of the CUDA DBSCAN (ecb_dbscan_run, through the C ABI) with the oracle on seeded inputs.

Bar: labels bit-exact, including the discovery-order cluster ids and the reference's tie / border rules
(dbscan.h:115-265, kdtree.cpp:148-179).  The oracle itself is pinned against the unmodified reference in
tests/test_oracle_dbscan.py.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rand_points(rng, n, W, H):
    pts = np.unique(np.stack([rng.integers(0, W, n), rng.integers(0, H, n)], axis=1), axis=0)
    rng.shuffle(pts)
    return pts.astype(np.float64)


def test_toy_known_answer(ctx):
    # end points have a single neighbour -> noise, not border members
    xy = np.array([[4 * i, 0] for i in range(10)] + [[100, 100]], float)
    rc, lab, nc = ctx.dbscan(xy, 4, 2)
    assert rc == 0 and nc == 1
    assert lab.tolist() == [-1, 0, 0, 0, 0, 0, 0, 0, 0, -1, -1]


def test_failed_status(ctx):
    import eventcalib_b200 as ecb
    assert ctx.dbscan(np.zeros((0, 2)), 4, 2)[0] == ecb.FAILED
    assert ctx.dbscan(np.zeros((3, 2)), 4, 0)[0] == ecb.FAILED


@pytest.mark.parametrize("eps,minpts", [(4, 2), (2, 2), (3, 3), (6, 5), (8, 8), (4, 1), (2.5, 2), (4.5, 3)])
def test_random_dense(ctx, oracle_mod, eps, minpts):
    rng = np.random.default_rng(int(eps * 100) + minpts)
    for it in range(12):
        n = int(rng.integers(1, 3000))
        W = int(rng.integers(12, 140))
        pts = _rand_points(rng, n, W, W)
        ref = oracle_mod.dbscan(pts, eps, minpts)
        rc, lab, nc = ctx.dbscan(pts, eps, minpts)
        assert rc == 0
        assert nc == len(ref["clusters"])
        assert np.array_equal(lab, ref["labels"]), "labels differ (n=%d W=%d)" % (len(pts), W)


def test_single_point_and_offsets(ctx, oracle_mod):
    rc, lab, nc = ctx.dbscan(np.array([[5.0, 7.0]]), 4, 2)
    assert rc == 0 and nc == 0 and lab.tolist() == [-1]
    rng = np.random.default_rng(5)
    pts = _rand_points(rng, 800, 60, 40) + np.array([1000.0, 2000.0])  # bounding-box origin handling
    ref = oracle_mod.dbscan(pts, 4, 2)
    rc, lab, nc = ctx.dbscan(pts, 4, 2)
    assert np.array_equal(lab, ref["labels"])


def test_batch(ctx, oracle_mod):
    rng = np.random.default_rng(11)
    sets = [_rand_points(rng, int(rng.integers(1, 1500)), 90, 70) for _ in range(40)]
    off = np.concatenate([[0], np.cumsum([len(s) for s in sets])])
    lab, nc, st = ctx.dbscan_batch(np.concatenate(sets), off, 4, 2)
    assert not st.any()
    for k, s in enumerate(sets):
        ref = oracle_mod.dbscan(s, 4, 2)
        assert nc[k] == len(ref["clusters"])
        assert np.array_equal(lab[off[k]:off[k + 1]], ref["labels"])


def test_large_problem_global_scratch(ctx, oracle_mod):
    # > shared-memory array budget: per-point arrays go to L2 scratch, 640x480 bitmap
    rng = np.random.default_rng(3)
    pts = _rand_points(rng, 60000, 640, 480)
    ref = oracle_mod.dbscan(pts, 4, 2)
    rc, lab, nc = ctx.dbscan(pts, 4, 2)
    assert nc == len(ref["clusters"])
    assert np.array_equal(lab, ref["labels"])


def test_unsupported_inputs_fail_loudly(ctx):
    import eventcalib_b200 as ecb
    with pytest.raises(ecb.EcbError):
        ctx.dbscan(np.array([[np.nan, 1.0], [2.0, 3.0]]), 4, 2)
    with pytest.raises(ecb.EcbError):
        ctx.dbscan(np.array([[1.0, 1.0], [np.inf, 1.0]]), 4, 2)


# ---- the general (grid-hash) path: everything DBSCAN<T,Float>::Run accepts (dbscan.h:40,70,115-177) -------------------------
def _check_general(ctx, oracle_mod, pts, eps, minpts, what):
    """ORDERED `Clusters`, `Noise` and labels against the unmodified reference (oracle/_ref; the restated oracle otherwise)."""
    pts = np.ascontiguousarray(pts, np.float64)
    if pts.shape[1] == 2:
        ref = (oracle_mod.ref_dbscan if oracle_mod.have_ref() else oracle_mod.dbscan)(pts, eps, minpts)
        rc, lab, clusters, noise = ctx.dbscan_ordered(pts, eps, minpts)
        rc2, lab2, nc2 = ctx.dbscan(pts, eps, minpts)  # unordered entry point: tree only if a pair needs it
        assert rc2 == 0 and nc2 == len(ref["clusters"]) and np.array_equal(lab2, ref["labels"]), what + ": unordered labels"
    else:
        if not oracle_mod.have_ref():
            pytest.skip("n-D reference wrapper needs oracle/_ref")
        ref = oracle_mod.ref_dbscan_nd(pts, eps, minpts)
        rc, lab, clusters, noise = ctx.dbscan_nd(pts, eps, minpts)
    assert rc == 0 and ref["rc"] == 0
    assert np.array_equal(lab, ref["labels"]), what + ": labels differ"
    assert len(clusters) == len(ref["clusters"]), what
    for c, (g, r) in enumerate(zip(clusters, ref["clusters"])):
        assert np.array_equal(g, r), "%s: cluster %d member order differs" % (what, c)
    assert np.array_equal(noise, ref["noise"]), what
    return ref


@pytest.mark.parametrize("eps,minpts", [(1.7, 3), (0.9, 2), (3.3, 5), (2.0, 1)])
def test_general_float_coordinates(ctx, oracle_mod, eps, minpts):
    rng = np.random.default_rng(int(eps * 10) + minpts)
    for it in range(8):
        n = int(rng.integers(1, 2500))
        L = float(rng.uniform(5, 60))
        pts = rng.uniform(-L, L, (n, 2))
        ref = _check_general(ctx, oracle_mod, pts, eps, minpts, "uniform floats n=%d L=%.1f" % (n, L))
    assert len(ref["labels"])


def test_general_duplicates_and_integer_ties(ctx, oracle_mod):
    """Duplicate points are distinct pids and each other's neighbours (dbscan.h:218); on integer grids with integer eps the
    kd query's strict pruning (kdtree.cpp:166-171) drops exact-eps axis neighbours — both at once here."""
    rng = np.random.default_rng(77)
    for it in range(8):
        n = int(rng.integers(50, 2000))
        W = int(rng.integers(10, 60))
        pts = rng.integers(0, W, (n, 2)).astype(np.float64)  # many duplicates
        for eps, minpts in ((4, 2), (3, 4), (1, 1)):
            _check_general(ctx, oracle_mod, pts, eps, minpts, "integer duplicates n=%d W=%d eps=%g" % (n, W, eps))


def test_general_large_and_small_eps_on_pixels(ctx, oracle_mod):
    rng = np.random.default_rng(78)
    pts = _rand_points(rng, 1500, 120, 90)
    for eps, minpts in ((20, 4), (16, 30), (0.5, 1), (0.0, 1), (17.5, 10)):   # outside the bitmap kernel's [1, 15]
        _check_general(ctx, oracle_mod, pts, eps, minpts, "pixels eps=%g" % eps)
    _check_general(ctx, oracle_mod, np.concatenate([pts, pts[:300]]), 0.0, 1, "eps 0 with duplicates")
    _check_general(ctx, oracle_mod, pts - 60.0, 4, 2, "negative pixel coordinates")
    _check_general(ctx, oracle_mod, pts * 300.0, 1200, 2, "extent beyond the bitmap")


def test_general_half_integer_grid_and_sorted_input(ctx, oracle_mod):
    rng = np.random.default_rng(79)
    g = np.stack(np.meshgrid(np.arange(40), np.arange(30)), -1).reshape(-1, 2) * 0.5 + 0.25
    keep = rng.random(len(g)) < 0.6
    pts = g[keep]
    rng.shuffle(pts)
    for eps in (1.0, 1.5, 2.5):   # exact-eps axis neighbours in binary fractions: the pruning rule fires
        _check_general(ctx, oracle_mod, pts, eps, 3, "half-integer grid eps=%g" % eps)
    srt = pts[np.lexsort((pts[:, 1], pts[:, 0]))]  # sorted input: the insertion tree degenerates into long chains
    _check_general(ctx, oracle_mod, srt, 1.5, 3, "sorted input")
    _check_general(ctx, oracle_mod, np.repeat(np.array([[3.0, 4.0]]), 40, 0), 1.0, 5, "one point 40 times")


def test_general_large_float_problem(ctx, oracle_mod):
    rng = np.random.default_rng(80)
    c = rng.uniform(0, 400, (60, 2))
    pts = (c[rng.integers(0, 60, 40000)] + rng.normal(0, 3.0, (40000, 2)))
    _check_general(ctx, oracle_mod, pts, 1.2, 4, "40k clustered floats")


@pytest.mark.parametrize("dim", [1, 3, 4])
def test_general_other_dimensions(ctx, oracle_mod, dim):
    rng = np.random.default_rng(90 + dim)
    for it in range(4):
        n = int(rng.integers(1, 1500))
        pts = rng.uniform(0, 12, (n, dim)) if it % 2 == 0 else rng.integers(0, 8, (n, dim)).astype(np.float64)
        eps = 0.05 * 12 if dim == 1 else (2.0 if it % 2 else 1.3)
        _check_general(ctx, oracle_mod, pts, eps, 3, "dim %d case %d" % (dim, it))


def test_general_batch(ctx, oracle_mod):
    rng = np.random.default_rng(81)
    sets = [rng.uniform(0, 30, (int(rng.integers(1, 900)), 2)) for _ in range(25)]
    sets[3] = np.round(sets[3])          # duplicates + integer ties inside a float batch
    off = np.concatenate([[0], np.cumsum([len(s) for s in sets])])
    ref_fn = oracle_mod.ref_dbscan if oracle_mod.have_ref() else oracle_mod.dbscan
    lab, nc, st, clusters = ctx.dbscan_batch_ordered(np.concatenate(sets), off, 2.0, 3)
    lab2, nc2, st2 = ctx.dbscan_batch(np.concatenate(sets), off, 2.0, 3)
    assert np.array_equal(lab, lab2) and np.array_equal(nc, nc2)
    for k, pts in enumerate(sets):
        ref = ref_fn(pts, 2.0, 3)
        assert np.array_equal(lab[off[k]:off[k + 1]], ref["labels"]), "problem %d" % k
        assert len(clusters[k]) == len(ref["clusters"])
        for g, r in zip(clusters[k], ref["clusters"]):
            assert np.array_equal(g, r)


@pytest.mark.parametrize("eps,minpts", [(4, 2), (2, 2), (3, 3), (6, 5), (4, 1), (2.5, 2)])
def test_ordered_clusters_equal_the_reference_lists(ctx, oracle_mod, eps, minpts):
    """`Clusters` as ORDERED lists (BFS pop order, kd result-list neighbour order; dbscan.h:229-259, kdtree.cpp:148-179,
    469-486) and `Noise`, against the unmodified reference DBSCAN compiled in oracle/_ref (the restated oracle when the
    prebuilt reference library is absent)."""
    ref_fn = oracle_mod.ref_dbscan if oracle_mod.have_ref() else oracle_mod.dbscan
    rng = np.random.default_rng(1000 + int(eps * 10) + minpts)
    for it in range(10):
        n = int(rng.integers(1, 2500))
        W = int(rng.integers(12, 120))
        pts = _rand_points(rng, n, W, W)
        ref = ref_fn(pts, eps, minpts)
        rc, lab, clusters, noise = ctx.dbscan_ordered(pts, eps, minpts)
        assert rc == 0 and np.array_equal(lab, ref["labels"])
        assert len(clusters) == len(ref["clusters"])
        for c, (g, r) in enumerate(zip(clusters, ref["clusters"])):
            assert np.array_equal(g, r), "cluster %d member order differs (n=%d W=%d)" % (c, len(pts), W)
        assert np.array_equal(noise, ref["noise"])


def test_ordered_circle_frames_and_batch(ctx, oracle_mod):
    from eventcalib_b200 import synth
    ref_fn = oracle_mod.ref_dbscan if oracle_mod.have_ref() else oracle_mod.dbscan
    ev = synth.make_stream(30000, 346, 260, t0=5.0, duration=0.015, seed=21)
    sets = []
    for w in synth.tiling_windows(5.0, 5.015, 1.5e-3):
        P, N, _, _ = oracle_mod.event_frame(ev["t"], ev["x"], ev["y"], ev["p"], w[0], w[1])
        sets += [P, N]
    off = np.concatenate([[0], np.cumsum([len(s) for s in sets])])
    lab, nc, st, clusters = ctx.dbscan_batch_ordered(np.concatenate(sets), off, 4, 2)
    assert (st == 0).all()
    for k, pts in enumerate(sets):
        ref = ref_fn(pts, 4, 2)
        assert np.array_equal(lab[off[k]:off[k + 1]], ref["labels"])
        assert len(clusters[k]) == len(ref["clusters"])
        for g, r in zip(clusters[k], ref["clusters"]):
            assert np.array_equal(g, r)
