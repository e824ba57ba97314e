#!/usr/bin/env python
"""Side measurements of the BASELINE configs that are not the bench line (one GPU, front end only, device-resident):

  C3 per GPU : 640x480, 25 M events (1/8 of 200 M), 10 ms tiling windows (~1e5 events each -> per-point arrays in L2 scratch)
  C5 slice   : 1280x720, 20 M events at 100 Mev/s, 1 ms windows, 20 % noise + 5 % polarity flips, eps x minPts sweep
               (bit planes in L2 scratch)

  python profiles/tools/side_configs.py > profiles/rX_side_configs.jsonl
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import eventcalib_b200 as ecb  # noqa: E402
from eventcalib_b200 import synth  # noqa: E402


def run(name, width, height, n_events, rate, window, seed, sweeps, noise=0.05, flip=0.0, steps=3):
    dur = n_events / rate
    ev = synth.make_stream(n_events, width, height, t0=0.0, duration=dur, seed=seed, noise_frac=noise, flip_frac=flip,
                           workers=min(16, os.cpu_count() or 1))
    win = synth.tiling_windows(0.0, dur, window)
    rec = synth.to_records(ev)
    d_raw = torch.empty(n_events * 25 + 16, dtype=torch.uint8, device="cuda")
    d_raw[: n_events * 25].copy_(torch.from_numpy(rec.view(np.uint8).reshape(-1)))
    ctx = ecb.Context(0)
    ctx.set_sensor(width, height)
    rthr = ecb.radius_threshold(width, height, 9, 4, True, 5.5, 1.75)
    for eps, mp in sweeps:
        prm = ecb.default_params(eps=float(eps), min_pts=mp, fit_circle=1, radius_threshold=rthr, order_mode=1, median_mode=1)
        ctx.set_profiling(True)
        ms = []
        for it in range(steps + 1):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ctx.load_events_device(d_raw.data_ptr(), n_events)
            ctx.frontend_run(win, prm)
            ctx.synchronize()
            if it:
                ms.append((time.perf_counter() - t0) * 1e3)
        st = {k: round(v, 3) for k, v in ctx.stage_ms().items() if v > 0}
        s = ctx.summary()
        print(json.dumps({"config": name, "sensor": [width, height], "events": n_events, "windows": int(len(win)),
                          "window_ms": window * 1e3, "eps": eps, "min_pts": mp, "ms_per_pass": float(np.median(ms)),
                          "events_per_s": n_events / (np.median(ms) * 1e-3), "stage_ms": st,
                          "points": int(s["n_points"].sum()), "clusters": int(s["n_clusters"].sum()),
                          "candidates_per_window": float(s["n_candidates"].mean()), "status_or": int(np.bitwise_or.reduce(s["status"]))}),
              flush=True)
    ctx.close()


if __name__ == "__main__":
    run("C3 per GPU", 640, 480, 25_000_000, 10e6, 10e-3, 1003, [(4, 2)])
    run("C5 slice", 1280, 720, 20_000_000, 100e6, 1e-3, 1005, [(e, m) for e in (2, 4, 8) for m in (2, 5)], noise=0.2, flip=0.05,
        steps=2)
