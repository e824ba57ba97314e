"""Parity of the batched CUDA front end (ecb_load_events + ecb_frontend_run) with the oracle.

Per window: event range, per-polarity pixel sets after dedupe and +/- cancellation (EventFrame.cpp:10-36),
pid order (order_mode 0 = first arrival, order_mode 1 = the reference's libstdc++ unordered_set iteration order), DBSCAN labels bit-exact, kept clusters, medians, candidate pairs
exact and circle centres / radii within 1e-9 relative (CirclesEventFrame.cpp:61-312,361-415).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-9


def _first_arrival(ev, lo, hi, pol):
    x = ev["x"][lo:hi].astype(np.int64)
    y = ev["y"][lo:hi].astype(np.int64)
    p = ev["p"][lo:hi]
    key = y * 100000 + x
    own = key[p == pol]
    other = set(key[p != pol].tolist())
    seen = set()
    out = []
    for k in own.tolist():
        if k not in seen:
            seen.add(k)
            if k not in other:
                out.append((k % 100000, k // 100000))
    return np.array(out, np.float64).reshape(-1, 2)


def _check_stream(ctx, oracle_mod, ev, windows, width, height, fit_circle, eps=4.0, min_pts=2, order_mode=0,
                  median_mode=0, cluster_min=5, stats=None, rows_cols=36):
    import eventcalib_b200 as ecb
    from eventcalib_b200 import synth
    ctx.set_sensor(width, height)
    n = ctx.load_events(synth.to_records(ev))
    assert n == len(ev["t"])
    rthr = ecb.radius_threshold(width, height, 9, 4, True, 5.5, 1.75)
    prm = ecb.default_params(eps=eps, min_pts=min_pts, fit_circle=fit_circle, radius_threshold=rthr,
                             order_mode=order_mode, median_mode=median_mode, cluster_min=cluster_min, rows_cols=rows_cols)
    ctx.frontend_run(windows, prm)
    summ = ctx.summary()
    pts = [ctx.points(0), ctx.points(1)]
    cand = ctx.candidates(max(64, int(summ["n_candidates"].max())))
    if stats is not None:
        stats["max_kept"] = max(stats.get("max_kept", 0), int(summ["n_kept"].max()))
    n_cand_total = 0
    for w, (t0, t1) in enumerate(windows):
        P, N, lo, hi = oracle_mod.event_frame(ev["t"], ev["x"], ev["y"], ev["p"], t0, t1)
        s = summ[w]
        assert (s["ev_lo"], s["ev_hi"]) == (lo, hi)
        assert s["status"] == 0
        V = []
        for pol, ref_set in ((0, N), (1, P)):
            o, k = int(s["point_offset"][pol]), int(s["n_points"][pol])
            xy = pts[pol][0][o:o + k]
            assert k == len(ref_set)
            assert set(map(tuple, xy.tolist())) == set(map(tuple, ref_set.tolist()))
            if order_mode == 1:   # the reference's own pid order: iteration order of the real std::unordered_set
                assert np.array_equal(xy, ref_set), "pid order differs from the libstdc++ set order, window %d" % w
            else:
                assert np.array_equal(xy, _first_arrival(ev, lo, hi, pol))
            V.append(xy)
        # median_mode 1: the oracle runs the real std::nth_element over the BFS-ordered member lists, like the reference
        r = oracle_mod.extract(V[1], V[0], eps=eps, minS=min_pts, fitCircle=fit_circle, Rthr=rthr,
                               canonical_median=(median_mode == 0), clusterMin=cluster_min, rows_cols=rows_cols)
        for pol, key in ((0, "n"), (1, "p")):
            o, k = int(s["point_offset"][pol]), int(s["n_points"][pol])
            assert np.array_equal(pts[pol][1][o:o + k], r[key + "_labels"]), "labels differ window %d pol %d" % (w, pol)
            assert s["n_clusters"][pol] == len(r[key + "_clusters"])
            raw, size, med = ctx.clusters(w, pol, cap=max(512, int(s["n_kept"][pol])))
            assert np.array_equal(raw, r["kept_" + key])
            assert np.array_equal(size, [len(r[key + "_clusters"][c]) for c in raw])
            if r["enough"]:
                assert np.array_equal(med, r["med_" + key])
        assert s["n_candidates"] == len(r["cand"])
        n_cand_total += len(r["cand"])
        if len(r["cand"]):
            g = cand[w, :len(r["cand"])]
            assert np.array_equal(g[:, :2], r["cand"][:, :2])
            np.testing.assert_allclose(g[:, 2:], r["cand"][:, 2:], rtol=RTOL, atol=0)
    return n_cand_total


@pytest.mark.parametrize("order_mode", [0, 1])
@pytest.mark.parametrize("fit_circle", [0, 1])
def test_davis346_windows(ctx, oracle_mod, fit_circle, order_mode):
    from eventcalib_b200 import synth
    ev = synth.make_stream(120000, 346, 260, t0=5.0, duration=0.06, seed=1001)
    win = synth.tiling_windows(5.0, 5.06, 1.5e-3)
    total = _check_stream(ctx, oracle_mod, ev, win, 346, 260, fit_circle, order_mode=order_mode)
    assert total > 30 * len(win)  # the synthetic board is found in (almost) every window


@pytest.mark.parametrize("fit_circle", [0, 1])
def test_reference_exact_mode(ctx, oracle_mod, fit_circle):
    """order_mode 1 + median_mode 1: the reference's own pid order (libstdc++ unordered_set) and its own cluster centres
    (std::nth_element over DBSCAN's BFS-ordered member lists) — no documented deviation left on this path."""
    from eventcalib_b200 import synth
    ev = synth.make_stream(240000, 346, 260, t0=5.0, duration=0.12, seed=1001)
    win = synth.tiling_windows(5.0, 5.12, 1.5e-3)
    total = _check_stream(ctx, oracle_mod, ev, win, 346, 260, fit_circle, order_mode=1, median_mode=1)
    assert total > 30 * len(win)
    # the tie clusters are really there: the canonical medians differ somewhere on this stream
    import eventcalib_b200 as ecb
    rthr = ecb.radius_threshold(346, 260, 9, 4, True, 5.5, 1.75)
    meds = []
    for mm in (0, 1):
        ctx.frontend_run(win, ecb.default_params(fit_circle=fit_circle, radius_threshold=rthr, order_mode=1, median_mode=mm))
        meds.append(np.concatenate([ctx.clusters(w, pol)[2] for w in range(len(win)) for pol in (0, 1)]))
    assert len(meds[0]) == len(meds[1]) and (meds[0] != meds[1]).any()


def _shape_stream(seed, sorted_noise):
    """Hand-built window: clusters that are symmetric about the image diagonal (x <-> y), so most squared norms occur twice and
    the median norm of nearly every cluster is tied — each of them goes through the member-order pass.  Sizes / densities
    straddle the limits of its small-cluster path (48 members, 32 visible neighbours, tree depth 48): 5x5 .. 8x8 blocks,
    anti-diagonal lines of 47 / 48 / 49 / 60 pixels, rings, a filled disc; `sorted_noise` adds isolated pixels that arrive
    first in ascending x — with order_mode 0 they make the emulated kd-tree a chain more than 48 levels deep."""
    rng = np.random.default_rng(seed)
    px = []
    c = 12
    for k in (5, 6, 7, 8):                      # k x k blocks centred on the diagonal
        px += [(c + i, c + j) for i in range(k) for j in range(k)]
        c += k + 14
    for L in (47, 48, 49, 60):                  # anti-diagonal lines through a diagonal pixel
        px += [(c + 30 + i - L // 2, c + 30 - (i - L // 2)) for i in range(L)]
        c += 24
    for r, cy in ((5.0, 120.0), (7.5, 170.0)):  # rings (left of the diagonal shapes)
        for a in np.linspace(0, 2 * np.pi, int(8 * r), endpoint=False):
            px.append((int(round(30.0 + r * np.cos(a))), int(round(cy + r * np.sin(a)))))
    px += [(238 + i, 238 + j) for i in range(-6, 7) for j in range(-6, 7) if i * i + j * j <= 36]   # filled disc on the diagonal
    for a in np.linspace(0, 2 * np.pi, 40, endpoint=False):                                      # a small ring far to the right
        px.append((int(round(322.0 + 5.0 * np.cos(a))), int(round(120.0 + 5.0 * np.sin(a)))))
    px = list(dict.fromkeys(px))
    assert all(0 <= x < 346 and 0 <= y < 260 for x, y in px)
    px = [px[i] for i in rng.permutation(len(px))]
    # isolated pixels along y = 0 with ascending x: inserted first and in this order (order_mode 0) they form ONE chain, so a
    # cluster at large x hangs more than 48 levels deep in the emulated tree
    noise = [(6 * i, 0) for i in range(57)] if sorted_noise else []
    noise = [(x, y) for x, y in noise if 0 <= x < 346 and 0 <= y < 260]
    pos = noise + px
    # negative polarity: plain 3x3 blocks in the strip x >= 270 that the shapes never reach (nothing cancels)
    neg = [(272 + 12 * (b % 6) + i, 10 + 40 * (b // 6) + j) for b in range(12) for i in range(3) for j in range(3)]
    assert not set(pos) & set(neg)
    ev = {"t": [], "x": [], "y": [], "p": []}
    t = 5.0
    for pol, pts in ((1, pos), (0, neg)):
        for x, y in pts:
            t += 1e-7
            ev["t"].append(t)
            ev["x"].append(x)
            ev["y"].append(y)
            ev["p"].append(pol)
    return {"t": np.array(ev["t"]), "x": np.array(ev["x"], np.uint16), "y": np.array(ev["y"], np.uint16),
            "p": np.array(ev["p"], np.uint8)}


@pytest.mark.parametrize("order_mode", [0, 1])
def test_member_order_small_cluster_limits(ctx, oracle_mod, order_mode):
    """median_mode 1 on clusters around every limit of k_bfs_order's small-cluster path (ecb_bfs.cu: bfs_small) and on the
    walk it falls back to: labels, kept clusters and std::nth_element medians equal the reference restatement."""
    ev = _shape_stream(5, sorted_noise=True)
    win = np.array([[4.0, 6.0]])
    n_cand = _check_stream(ctx, oracle_mod, ev, win, 346, 260, 0, order_mode=order_mode, median_mode=1, rows_cols=4)
    assert n_cand >= 0
    # the medians are really order-dependent here: the canonical rule picks a different member somewhere
    import eventcalib_b200 as ecb
    rthr = ecb.radius_threshold(346, 260, 9, 4, True, 5.5, 1.75)
    meds = []
    for mm in (0, 1):
        ctx.frontend_run(win, ecb.default_params(radius_threshold=rthr, order_mode=order_mode, median_mode=mm, rows_cols=4))
        meds.append(np.concatenate([ctx.clusters(0, pol)[2] for pol in (0, 1)]))
    assert len(meds[0]) == len(meds[1]) >= 16 and (meds[0] != meds[1]).any()


def test_overlapping_and_empty_windows(ctx, oracle_mod):
    from eventcalib_b200 import synth
    ev = synth.make_stream(40000, 346, 260, t0=5.0, duration=0.02, seed=7)
    win = np.array([[5.0, 5.0015], [5.0005, 5.003], [4.0, 4.5], [5.019, 9.0], [5.002, 5.002], [5.0, 5.02]])
    _check_stream(ctx, oracle_mod, ev, win, 346, 260, 0)
    _check_stream(ctx, oracle_mod, ev, win, 346, 260, 0, order_mode=1)


def test_set_order_window_sizes(ctx, oracle_mod):
    """libstdc++ order emulation across every rehash boundary (13, 29, 59, ... buckets): windows of growing length."""
    from eventcalib_b200 import synth
    ev = synth.make_stream(60000, 346, 260, t0=5.0, duration=0.03, seed=11, noise_frac=0.3)
    lens = [2e-6, 5e-6, 8e-6, 1.5e-5, 3e-5, 6e-5, 1.3e-4, 2.7e-4, 5.5e-4, 1.1e-3, 2.3e-3, 5e-3, 1e-2, 3e-2]
    win = np.array([[5.0 + 1e-4 * k, 5.0 + 1e-4 * k + L] for k, L in enumerate(lens)])
    _check_stream(ctx, oracle_mod, ev, win, 346, 260, 1, order_mode=1)


def test_vga_large_windows(ctx, oracle_mod):
    from eventcalib_b200 import synth
    ev = synth.make_stream(300000, 640, 480, t0=0.0, duration=0.03, seed=1003)
    win = synth.tiling_windows(0.0, 0.03, 10e-3)
    _check_stream(ctx, oracle_mod, ev, win, 640, 480, 1)
    _check_stream(ctx, oracle_mod, ev, win, 640, 480, 1, order_mode=1, median_mode=1)


def test_hd_sensor_noise_sweep(ctx, oracle_mod):
    """BASELINE config C5 in miniature: 1280x720, polarity noise, 1 ms windows, eps / minPts sweep.  The bit planes of
    this sensor exceed one CTA's shared memory, so the kernels run them from per-CTA L2 scratch."""
    from eventcalib_b200 import synth
    ev = synth.make_stream(200000, 1280, 720, t0=0.0, duration=0.002, seed=1005, noise_frac=0.2, flip_frac=0.05)
    win = synth.tiling_windows(0.0, 0.002, 1e-3)
    _check_stream(ctx, oracle_mod, ev, win, 1280, 720, 1, order_mode=1, median_mode=1)
    for eps, mp in ((2, 2), (3, 5), (6, 3), (8, 8)):
        _check_stream(ctx, oracle_mod, ev, win, 1280, 720, 0, eps=float(eps), min_pts=mp)


def test_many_kept_clusters_grow_the_tables(ctx, oracle_mod):
    """The reference has no cap on the clusters a window keeps (CirclesEventFrame.cpp:89-117).  max_clusters = 0 sizes the
    kept-cluster tables from the data: noisy windows with hundreds of kept clusters come back complete (status 0, every
    kept cluster / median / candidate equal to the oracle's); a fixed capacity truncates and says so."""
    import eventcalib_b200 as ecb
    from eventcalib_b200 import synth
    ev = synth.make_stream(120000, 1280, 720, t0=0.0, duration=0.002, seed=77, noise_frac=0.6, flip_frac=0.05)
    win = synth.tiling_windows(0.0, 0.002, 1e-3)
    st = {}
    _check_stream(ctx, oracle_mod, ev, win, 1280, 720, 1, eps=8.0, min_pts=2, cluster_min=2, order_mode=1, median_mode=1, stats=st)
    _check_stream(ctx, oracle_mod, ev, win, 1280, 720, 0, eps=3.0, min_pts=2, cluster_min=2, stats=st)
    assert st["max_kept"] > 512, st   # beyond the old static limits (128 default, 512 maximum)
    ev = synth.make_stream(60000, 346, 260, t0=0.0, duration=0.003, seed=78, noise_frac=0.7, flip_frac=0.1)
    win = synth.tiling_windows(0.0, 0.003, 1.5e-3)
    st = {}
    _check_stream(ctx, oracle_mod, ev, win, 346, 260, 1, eps=2.0, min_pts=2, cluster_min=2, order_mode=1, median_mode=1, stats=st)
    assert st["max_kept"] > 128, st
    # fixed capacity: truncated tables are flagged
    prm = ecb.default_params(eps=2.0, min_pts=2, cluster_min=2, max_clusters=16)
    ctx.frontend_run(win, prm)
    assert all(int(v) & ecb.PB_CLUSTER_CAP for v in ctx.summary()["status"])


def test_eps_minpts_sweep(ctx, oracle_mod):
    from eventcalib_b200 import synth
    ev = synth.make_stream(30000, 346, 260, t0=5.0, duration=0.015, seed=1005, noise_frac=0.2, flip_frac=0.05)
    win = synth.tiling_windows(5.0, 5.015, 1.5e-3)
    for eps in (2, 3, 4, 6, 8):
        for mp in (2, 3, 5, 8):
            _check_stream(ctx, oracle_mod, ev, win, 346, 260, 0, eps=float(eps), min_pts=mp)


def ecb_params(ctx):
    import eventcalib_b200 as ecb
    return ecb.default_params(fit_circle=1, radius_threshold=ecb.radius_threshold(346, 260, 9, 4, True, 5.5, 1.75), order_mode=1,
                              median_mode=1)


def test_rejects_bad_streams(ctx):
    import eventcalib_b200 as ecb
    from eventcalib_b200 import synth
    ev = synth.make_stream(1000, 346, 260, duration=0.001, seed=3)
    ctx.set_sensor(346, 260)
    bad = dict(ev)
    bad["x"] = ev["x"].copy()
    bad["x"][10] = 12.5
    with pytest.raises(ecb.EcbError):
        ctx.load_events(synth.to_records(bad))


def test_unsorted_stream_is_ordered_like_the_multimap_load(ctx, oracle_mod):
    """The reference loads the records into a std::multimap keyed by the stamp (eventCameraCalib.cpp:154-163): any file
    order is accepted and equal stamps keep their file order.  Device: a stable radix sort by stamp in ecb_load_events_*."""
    from eventcalib_b200 import synth
    ev = synth.make_stream(40000, 346, 260, t0=5.0, duration=0.015, seed=31)
    ev["t"] = np.round(ev["t"], 5)  # many equal stamps: stability matters (first arrival decides the pid order)
    rng = np.random.default_rng(5)
    perm = rng.permutation(len(ev["t"]))
    shuf = {k: (v[perm] if isinstance(v, np.ndarray) and len(v) == len(perm) else v) for k, v in ev.items()}
    order = np.argsort(shuf["t"], kind="stable")  # what the multimap iteration yields
    srt = {k: (v[order] if isinstance(v, np.ndarray) and len(v) == len(perm) else v) for k, v in shuf.items()}
    win = synth.tiling_windows(5.0, 5.015, 1.5e-3)
    ctx.set_sensor(346, 260)
    prm = ecb_params(ctx)
    ctx.load_events(synth.to_records(shuf))
    ctx.frontend_run(win, prm)
    got = (ctx.summary().copy(), ctx.points(0), ctx.points(1), ctx.candidates(64).copy())
    ctx.load_events(synth.to_records(srt))
    ctx.frontend_run(win, prm)
    ref = (ctx.summary().copy(), ctx.points(0), ctx.points(1), ctx.candidates(64).copy())
    assert np.array_equal(got[0], ref[0])
    for a, b in zip(got[1] + got[2], ref[1] + ref[2]):
        assert np.array_equal(a, b)
    assert np.array_equal(got[3], ref[3])
    # and against the oracle's window on the stably sorted stream
    s0 = got[0][0]
    P, N, lo, hi = oracle_mod.event_frame(srt["t"], srt["x"], srt["y"], srt["p"], win[0, 0], win[0, 1])
    assert (int(s0["ev_lo"]), int(s0["ev_hi"])) == (lo, hi)
    assert np.array_equal(got[2][0][:len(P)], P) and np.array_equal(got[1][0][:len(N)], N)


def test_fit_circles_api(ctx, oracle_mod):
    rng = np.random.default_rng(0)
    sets, ref = [], []
    for k in range(50):
        c = rng.uniform(50, 200, 2)
        r = rng.uniform(5, 15)
        th = rng.uniform(0, 2 * np.pi, int(rng.integers(8, 200)))
        p = np.rint(np.stack([c[0] + r * np.cos(th), c[1] + r * np.sin(th)], 1))
        sets.append(p)
        ref.append(oracle_mod.fit_circle(p[: len(p) // 2], p[len(p) // 2:]))
    off = np.concatenate([[0], np.cumsum([len(s) for s in sets])])
    out = ctx.fit_circles(np.concatenate(sets), off)
    np.testing.assert_allclose(out, np.array(ref), rtol=RTOL)


def _image_points(truth, t_mid, shift=(0.0, 0.0), scale=1.0):
    """projected circle centre + four quadrant points (CirclesEventFrame.cpp:431-456) through the ground-truth camera,
    rounded to float like cv::projectPoints' Point2f output"""
    from eventcalib_b200 import synth
    board, cam, traj = truth["board"], truth["camera"], truth["trajectory"]
    c = board.centres()
    k = board.radius / np.sqrt(2) * scale
    offs = np.array([[0, 0, 0], [k, k, 0], [k, -k, 0], [-k, -k, 0], [-k, k, 0]])
    X = (c[:, None, :] + offs[None, :, :]).reshape(-1, 3)
    R, tw = traj.pose(np.full(len(X), t_mid))
    u, v = synth.project(cam, R, tw, X)
    img = np.stack([u + shift[0], v + shift[1]], axis=1).reshape(len(c), 5, 2)
    return img.astype(np.float32).astype(np.float64)


@pytest.mark.parametrize("fit_circle", [0, 1])
def test_rectify_features(ctx, oracle_mod, fit_circle):
    """a6 rectifyFeatures (CirclesEventFrame.cpp:417-609) batched over frames x circles: rectified centres / radii within
    1e-9 of the oracle, deleted features and frame verdicts identical; good, shifted and mis-scaled projections."""
    import eventcalib_b200 as ecb
    from eventcalib_b200 import synth
    ev = synth.make_stream(90000, 346, 260, t0=5.0, duration=0.045, seed=31, return_truth=True)
    win = synth.tiling_windows(5.0, 5.045, 1.5e-3)
    ctx.set_sensor(346, 260)
    ctx.load_events(synth.to_records(ev))
    rthr = ecb.radius_threshold(346, 260, 9, 4, True, 5.5, 1.75)
    ctx.frontend_run(win, ecb.default_params(fit_circle=fit_circle, radius_threshold=rthr, order_mode=1, median_mode=1))
    summ = ctx.summary()
    pts = [ctx.points(0), ctx.points(1)]
    frames = np.arange(len(win), dtype=np.int32)
    imgs = []
    for w in frames:
        mid = 0.5 * (win[w, 0] + win[w, 1])
        variant = w % 5
        shift = {0: (0, 0), 1: (0.4, -0.3), 2: (7.0, 2.0), 3: (0, 0), 4: (-2.5, 1.5)}[variant]
        imgs.append(_image_points(ev, mid, shift=shift, scale=1.6 if variant == 3 else 1.0))
    imgs = np.array(imgs)
    imgs[7, 3, 0] = [-4.0, 10.0]      # a centre outside the image
    out, ok = ctx.rectify(frames, imgs)
    n_alive = 0
    for w in frames:
        s = summ[w]
        V = [pts[pol][0][int(s["point_offset"][pol]):int(s["point_offset"][pol]) + int(s["n_points"][pol])] for pol in (0, 1)]
        ref, ref_ok = oracle_mod.rectify(V[1], V[0], imgs[w], 346, 260, fitCircle=fit_circle)
        assert np.array_equal(out[w, :, 2] < 0, ref[:, 2] < 0), "deleted features differ, frame %d" % w
        alive = ref[:, 2] >= 0
        n_alive += int(alive.sum())
        np.testing.assert_allclose(out[w][alive], ref[alive], rtol=RTOL, atol=0)
        assert ok[w] == ref_ok
    assert n_alive > 10 * len(win) and (ok == 1).any() and (ok == 0).any()


def test_closed_window_bounds_and_equal_stamps(ctx, oracle_mod):
    """EventFrame.cpp:14-15: lower_bound(first) .. upper_bound(second) — both ends closed; events with EQUAL time stamps
    (multimap keeps file order) straddling a window bound are all inside"""
    from eventcalib_b200 import synth
    ev = synth.make_stream(20000, 346, 260, t0=5.0, duration=0.01, seed=77)
    t = ev["t"].copy()
    for k in (3000, 3001, 3002, 9000, 9001, 15000):      # runs of identical stamps
        t[k] = t[k - 1]
    ev = dict(ev, t=t)
    win = np.array([[t[2999], t[9001]],        # starts and ends exactly on repeated stamps
                    [t[9001], t[15000]],       # shares its first stamp with the previous window's last
                    [t[100], t[100]],          # a single instant
                    [np.nextafter(t[3002], 10), np.nextafter(t[9000], 0)],   # just inside the repeated stamps
                    [t[0], t[-1]]])
    _check_stream(ctx, oracle_mod, ev, win, 346, 260, 1, order_mode=1, median_mode=1)
    s = ctx.summary()
    assert s["ev_lo"][0] == 2999 and s["ev_hi"][0] == 9002 and s["ev_lo"][1] == 8999 and s["ev_hi"][2] - s["ev_lo"][2] == 1
