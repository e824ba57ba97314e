#!/usr/bin/env python
"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): front end in the reference-exact modes, ordered
DBSCAN, rectify, cost evaluation (both rotation models) on a few windows.

  compute-sanitizer --tool memcheck  python profiles/tools/sanitize_small.py
  compute-sanitizer --tool racecheck python profiles/tools/sanitize_small.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import eventcalib_b200 as ecb  # noqa: E402
from eventcalib_b200 import synth, calib_problem  # noqa: E402

ev = synth.make_stream(12000, 346, 260, t0=5.0, duration=0.006, seed=42, return_truth=True)
win = synth.tiling_windows(5.0, 5.006, 1.5e-3)
ctx = ecb.Context(0)
ctx.set_sensor(346, 260)
ctx.load_events(synth.to_records(ev))
rthr = ecb.radius_threshold(346, 260, 9, 4, True, 5.5, 1.75)
ctx.frontend_run(win, ecb.default_params(fit_circle=1, radius_threshold=rthr, order_mode=1, median_mode=1))
s = ctx.summary()
xy, lab = ctx.points(1)
img = np.zeros((len(win), 36, 5, 2))
img[..., 0] = 100.0
img[..., 1] = 100.0
ctx.rectify(np.arange(len(win), dtype=np.int32), img)
rc, lab, clusters, noise = ctx.dbscan_ordered(xy[: int(s["n_points"][0][1])], 4.0, 2)
ev2 = synth.make_stream(8000, 346, 260, t0=5.0, duration=0.05, seed=4, return_truth=True)
pb = calib_problem.build(ev2, seed=0)
ctx.load_events(synth.to_records(ev2))
ctx.cost_setup([pb["n_cp"]], [pb["knots"]], pb["radius"], pb["huber"])
ctx.cost_associate(pb["kf_t"], pb["circles"], pb["landmarks"], pb["step"])
x = (pb["intrinsics"], pb["rot_cp"], pb["trans_cp"])
ctx.cost_normal_eq(*x)
ctx.cost_eval(*x)
ctx.cost_set_rotation_model(1)
ctx.cost_normal_eq(*x)
ctx.cost_set_rotation_model(0)
# round 2: the device LM loop (blocked band Cholesky, staged back substitution) ...
lm = ecb.DeviceLm(ctx, [pb["n_cp"]], ecb.lm_options(max_iterations=3))
out = lm.run(pb["intrinsics"], pb["rot_cp"], pb["trans_cp"])
lm.close()
# ... the general (grid-hash) DBSCAN path: float coordinates, duplicates on an integer grid (tie rule + tree), 3-D, a batch ...
rng = np.random.default_rng(0)
ctx.dbscan_ordered(rng.uniform(0, 20, (400, 2)), 1.3, 3)
ctx.dbscan(rng.uniform(0, 20, (400, 2)), 1.3, 3)
ctx.dbscan_ordered(rng.integers(0, 14, (300, 2)).astype(np.float64), 3.0, 2)
ctx.dbscan_nd(rng.uniform(0, 8, (300, 3)), 1.2, 3)
ctx.dbscan_nd(rng.uniform(0, 8, (200, 4)), 1.5, 3)
sets = [rng.uniform(0, 15, (int(rng.integers(1, 120)), 2)) for _ in range(6)]
ctx.dbscan_batch_ordered(np.concatenate(sets), np.concatenate([[0], np.cumsum([len(q) for q in sets])]), 2.0, 3)
# ... and an unsorted event file (stable device sort by stamp)
perm = rng.permutation(len(ev["t"]))
ctx.load_events(synth.to_records({k: (v[perm] if isinstance(v, np.ndarray) and len(v) == len(perm) else v) for k, v in ev.items()}))
ctx.frontend_run(win, ecb.default_params(fit_circle=1, radius_threshold=rthr, order_mode=1, median_mode=1))
ctx.close()
print("sanitize run ok: %d windows, %d clusters, LM iterations %d" % (len(win), len(clusters), out["iterations"]))
