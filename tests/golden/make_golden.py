"""Generates tests/golden/*.npz from the UNMODIFIED reference DBSCAN compiled in place (oracle/_ref) and from the
oracle's libstdc++ unordered_set order.  Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

The fixtures pin the CPU restatement (and through it the CUDA path) on machines where /root/reference is absent."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

assert oracle.have_ref(), "oracle/_ref missing: run `make -C oracle` with /root/reference present"
rng = np.random.default_rng(20261017)
cases = {}
k = 0
for eps, mp in [(4, 2), (4, 2), (4, 2), (2, 2), (3, 3), (6, 5), (8, 8), (4, 1), (2.5, 2), (4.5, 3)]:
    n = int(rng.integers(50, 900))
    W = int(rng.integers(15, 70))
    pts = np.unique(np.stack([rng.integers(0, W, n), rng.integers(0, W, n)], 1), axis=0)
    rng.shuffle(pts)
    r = oracle.ref_dbscan(pts.astype(float), eps, mp)
    cases["pts_%d" % k] = pts.astype(np.int16)
    cases["par_%d" % k] = np.array([eps, mp], float)
    cases["labels_%d" % k] = r["labels"]
    cases["members_%d" % k] = np.concatenate(r["clusters"]) if r["clusters"] else np.zeros(0, np.uint32)
    cases["sizes_%d" % k] = np.array([len(c) for c in r["clusters"]], np.int32)
    cases["noise_%d" % k] = r["noise"]
    cases["kdq_%d" % k] = oracle.kd_range(pts.astype(float), 0, eps, ref=True)
    k += 1
cases["n_cases"] = np.array(k)
np.savez_compressed(os.path.join(os.path.dirname(__file__), "dbscan_reference.npz"), **cases)

# libstdc++ unordered_set<Vector2d, EigenMatrixHash> iteration order (EventFrame.cpp:12-35): insertion sequences -> order
us = {}
for j, n in enumerate([5, 13, 14, 30, 200, 1200]):
    pts = np.unique(np.stack([rng.integers(0, 346, 2 * n), rng.integers(0, 260, 2 * n)], 1), axis=0)
    rng.shuffle(pts)
    pts = pts[:n].astype(float)
    us["in_%d" % j] = pts.astype(np.int16)
    us["out_%d" % j] = oracle.uset_order(pts).astype(np.int16)
us["n_cases"] = np.array(6)
us["hash_1_0"] = np.array([oracle.port().orc_hash_double(1.0)], np.uint64)
us["hash_p_3_4"] = np.array([oracle.port().orc_hash_p2(3.0, 4.0)], np.uint64)
np.savez_compressed(os.path.join(os.path.dirname(__file__), "uset_order.npz"), **us)
print("wrote", k, "dbscan cases and 6 unordered_set cases")
