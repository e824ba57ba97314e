"""Generates tests/golden/circles_grid_hard.npz: HARD views for the grid finder (include/ecb/circles_grid.hpp) with OpenCV's
answers — the reference's call sequence, findCirclesGrid(CALIB_CB_ASYMMETRIC_GRID) and on failure the same with
CALIB_CB_CLUSTERING (CirclesEventFrame.cpp:332-336).  Views: ideal board projections under random in-plane rotation, tilts up
to 0.5 rad about both axes, distance 85 - 120, centre noise 0.1 / 0.3 / 0.6 px, and one of: nothing else, 1 - 3 outliers,
one circle missing, one missing + 1 - 2 outliers, 4 - 8 outliers.  OpenCV has no points-in Python entry point: the candidates
are rendered as filled discs (4x supersampled), the stock blob detector re-finds them and the returned centres are mapped back
to candidate indices.  Needs cv2 (present in the build container); run from the repo root:

    python tests/golden/make_golden_grid_hard.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from eventcalib_b200 import synth  # noqa: E402

S = 4
W, H = 346, 260
board, cam = synth.Board(), synth.Camera()
c = board.centres()
ctr = c.mean(0)
rng = np.random.default_rng(1)
KINDS = ["clean", "outliers", "missing", "missing+outlier", "many_outliers"]
out = {}
n_case = 0
for it in range(400):
    th, tilt, tilt2, dist = rng.uniform(0, 2 * np.pi), rng.uniform(-0.5, 0.5), rng.uniform(-0.5, 0.5), rng.uniform(85, 120)
    cz, sz = np.cos(th), np.sin(th)
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    ct, st = np.cos(tilt), np.sin(tilt)
    Rx = np.array([[1, 0, 0], [0, ct, -st], [0, st, ct]])
    c2, s2 = np.cos(tilt2), np.sin(tilt2)
    Ry = np.array([[c2, 0, s2], [0, 1, 0], [-s2, 0, c2]])
    R = Rz @ Rx @ Ry
    tw = ctr + R @ np.array([rng.uniform(-8, 8), rng.uniform(-8, 8), -dist])
    u, v = synth.project(cam, np.repeat(R[None], 36, 0), np.repeat(tw[None], 36, 0), c)
    pts = np.stack([u, v], 1) + rng.normal(0, rng.choice([0.1, 0.3, 0.6]), (36, 2))
    if pts[:, 0].min() < 8 or pts[:, 0].max() > W - 8 or pts[:, 1].min() < 8 or pts[:, 1].max() > H - 8:
        continue
    d = np.linalg.norm(pts[:, None] - pts[None], axis=2) + np.eye(36) * 1e9
    rad = 0.3 * d.min()
    kind = int(rng.integers(0, len(KINDS)))
    if "missing" in KINDS[kind]:
        pts = np.delete(pts, rng.integers(0, 36), axis=0)
    n_out = [0, int(rng.integers(1, 4)), 0, int(rng.integers(1, 3)), int(rng.integers(4, 9))][kind]
    for _ in range(n_out):
        for _try in range(50):
            q = np.array([rng.uniform(10, W - 10), rng.uniform(10, H - 10)])
            if np.min(np.linalg.norm(pts - q, axis=1)) > 5.1 * rad:
                pts = np.vstack([pts, q])
                break
    pts = pts[rng.permutation(len(pts))]
    img = np.full((H * S, W * S), 255, np.uint8)
    for x, y in pts:
        cv2.circle(img, (int(round(x * S)), int(round(y * S))), max(2, int(round(rad * S))), 0, -1, cv2.LINE_AA)
    which = 1
    ok, centers = cv2.findCirclesGrid(img, (4, 9), flags=cv2.CALIB_CB_ASYMMETRIC_GRID)
    if not ok:
        which = 2
        ok, centers = cv2.findCirclesGrid(img, (4, 9), flags=cv2.CALIB_CB_ASYMMETRIC_GRID | cv2.CALIB_CB_CLUSTERING)
    order = np.full(36, -1, np.int32)
    if ok:
        ce = centers.reshape(-1, 2) / S
        order = np.array([int(np.argmin(((pts - p) ** 2).sum(1))) for p in ce], np.int32)
    else:
        which = 0
    out["pts_%d" % n_case] = pts
    out["order_%d" % n_case] = order
    out["meta_%d" % n_case] = np.array([kind, which], np.int32)
    n_case += 1
out["n"] = np.array(n_case)
found = sum(int(out["order_%d" % i][0] >= 0) for i in range(n_case))
print("cases", n_case, "found by OpenCV", found)
np.savez_compressed(os.path.join(os.path.dirname(__file__), "circles_grid_hard.npz"), **out)
