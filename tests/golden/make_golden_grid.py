"""Generates tests/golden/circles_grid.npz: candidate circle centres of synthetic windows and the grid order OpenCV's
findCirclesGrid (CALIB_CB_ASYMMETRIC_GRID, same CirclesGridFinder as the reference's points-in overload,
cv_calib/src/cv_calib.cpp:8-88) returns for them.  OpenCV has no points-in Python entry point, so every candidate is rendered
as a filled disc (4x supersampled) and the stock blob detector re-finds the centres; the returned centres are mapped back
to candidate indices.  Needs cv2 (present in the build container); run from the repo root:

    python tests/golden/make_golden_grid.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from eventcalib_b200 import synth  # noqa: E402

S = 4
out = {}
n_case = 0
rng = np.random.default_rng(7)
for seed, (W, H) in ((1001, (346, 260)), (7, (346, 260)), (31, (346, 260)), (1003, (640, 480))):
    rate = 2e6 * (W * H) / (346 * 260)
    n_ev = int(rate * 0.03)
    ev = synth.make_stream(n_ev, W, H, t0=5.0, duration=0.03, seed=seed)
    rthr = oracle.radius_threshold(W, H, 9, 4, 1, 5.5, 1.75)
    for k, w in enumerate(synth.tiling_windows(5.0, 5.03, 1.5e-3)):
        if k % 2:
            continue
        P, N, _, _ = oracle.event_frame(ev["t"], ev["x"], ev["y"], ev["p"], w[0], w[1])
        r = oracle.extract(P, N, fitCircle=1, Rthr=rthr)
        pts = r["cand"][:, 2:4].copy()
        if len(pts) < 36:
            continue
        rad = float(np.median(r["cand"][:, 4]))
        extra = int(rng.integers(0, 3))            # false candidates away from the grid
        for _ in range(extra):
            for _try in range(50):
                q = np.array([rng.uniform(10, W - 10), rng.uniform(10, H - 10)])
                if np.min(np.linalg.norm(pts - q, axis=1)) > 4.5 * rad:
                    pts = np.vstack([pts, q])
                    break
        pts = pts[rng.permutation(len(pts))]
        img = np.full((H * S, W * S), 255, np.uint8)
        for x, y in pts:
            cv2.circle(img, (int(round(x * S)), int(round(y * S))), int(round(0.8 * rad * S)), 0, -1, cv2.LINE_AA)
        ok, centers = cv2.findCirclesGrid(img, (4, 9), flags=cv2.CALIB_CB_ASYMMETRIC_GRID)
        order = np.full(36, -1, np.int32)
        if ok:
            ce = centers.reshape(-1, 2) / S
            order = np.array([int(np.argmin(((pts - p) ** 2).sum(1))) for p in ce], np.int32)
            assert len(set(order.tolist())) == 36
        out["pts_%d" % n_case] = pts
        out["order_%d" % n_case] = order
        n_case += 1
out["n"] = np.array(n_case)
found = sum(int(out["order_%d" % i][0] >= 0) for i in range(n_case))
print("cases", n_case, "found by OpenCV", found)
np.savez_compressed(os.path.join(os.path.dirname(__file__), "circles_grid.npz"), **out)
