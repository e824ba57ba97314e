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
ctx.close()
print("sanitize run ok: %d windows, %d clusters" % (len(win), len(clusters)))
