"""Generates tests/golden/circles_grid_clustering.npz: candidate centres (ideal projections of the 9 x 4 asymmetric board under
in-plane rotations and tilts, jittered, plus 1 - 3 false candidates well away from the grid) and what OpenCV returns for them
with CALIB_CB_ASYMMETRIC_GRID and with CALIB_CB_ASYMMETRIC_GRID | CALIB_CB_CLUSTERING — the two calls of the reference
(CirclesEventFrame.cpp:332-336).  Rendering / blob detection as in make_golden_grid.py.  Needs cv2; run from the repo root:

    python tests/golden/make_golden_grid_clustering.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from eventcalib_b200 import synth  # noqa: E402

S, W, H = 4, 346, 260
board, cam = synth.Board(), synth.Camera()
c = board.centres()
ctr = c.mean(0)
rng = np.random.default_rng(11)
out = {}
n_case = 0


def ask(img, flags):
    ok, centers = cv2.findCirclesGrid(img, (4, 9), flags=flags)
    order = np.full(36, -1, np.int32)
    if ok:
        ce = centers.reshape(-1, 2) / S
        order = np.array([int(np.argmin(((pts - p) ** 2).sum(1))) for p in ce], np.int32)
        if len(set(order.tolist())) != 36:
            order[:] = -1
    return order


for k in range(60):
    th, tilt, tilt2 = rng.uniform(0, 2 * np.pi), rng.uniform(-0.6, 0.6), rng.uniform(-0.6, 0.6)
    cz, sz = np.cos(th), np.sin(th)
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    ct, st = np.cos(tilt), np.sin(tilt)
    c2, s2 = np.cos(tilt2), np.sin(tilt2)
    R = Rz @ np.array([[1, 0, 0], [0, ct, -st], [0, st, ct]]) @ np.array([[c2, 0, s2], [0, 1, 0], [-s2, 0, c2]])
    tw = ctr + R @ np.array([0, 0, -rng.uniform(95, 130)])
    u, v = synth.project(cam, np.repeat(R[None], 36, 0), np.repeat(tw[None], 36, 0), c)
    pts = np.stack([u, v], 1) + rng.normal(0, 0.15, (36, 2))
    if pts.min() < 12 or pts[:, 0].max() > W - 12 or pts[:, 1].max() > H - 12:
        continue
    d = np.sort(np.linalg.norm(pts[:, None] - pts[None], axis=2), axis=1)[:, 1]
    rad = 0.3 * float(np.median(d))
    for _ in range(int(rng.integers(1, 4))):
        for _try in range(200):
            q = np.array([rng.uniform(8, W - 8), rng.uniform(8, H - 8)])
            if np.min(np.linalg.norm(pts - q, axis=1)) > 2.2 * float(d.max()):
                pts = np.vstack([pts, q])
                break
    pts = pts[rng.permutation(len(pts))]
    img = np.full((H * S, W * S), 255, np.uint8)
    for x, y in pts:
        cv2.circle(img, (int(round(x * S)), int(round(y * S))), max(2, int(round(rad * S))), 0, -1, cv2.LINE_AA)
    out["pts_%d" % n_case] = pts
    out["std_%d" % n_case] = ask(img, cv2.CALIB_CB_ASYMMETRIC_GRID)
    out["clu_%d" % n_case] = ask(img, cv2.CALIB_CB_ASYMMETRIC_GRID | cv2.CALIB_CB_CLUSTERING)
    n_case += 1
out["n"] = np.array(n_case)
print("cases", n_case, "standard finds", sum(int(out["std_%d" % i][0] >= 0) for i in range(n_case)), "clustering finds",
      sum(int(out["clu_%d" % i][0] >= 0) for i in range(n_case)))
np.savez_compressed(os.path.join(os.path.dirname(__file__), "circles_grid_clustering.npz"), **out)
