"""The drop-in claim, executed: the reference's OWN CirclesEventFrame.cpp / EventFrame.cpp (compiled where they lie by
oracle/Makefile into oracle/_ref/libref_dropin.so, in the build container) with `#include <dbscan.h>` resolved to the PRODUCT's
include/ecb/dbscan.h — so `DBSCAN<Eigen::Vector2d, double> dbscan; dbscan.Run(&positiveEvents_, 2, eps, minS)` of
CirclesEventFrame.cpp:66-72 compiles unchanged and runs on the GPU through the C ABI (ecb_dbscan_run_ordered).  Everything
downstream of `Clusters` (size filter, std::nth_element medians, pairing, fitCircle, grid order, rectifyFeatures) is the
reference's own code, so its outputs must equal the golden outputs of the all-reference build (tests/golden/reference_source.npz)
— exactly, not within a tolerance: the ordered `Clusters` lists are the only thing that crossed the boundary."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "reference_source.npz"))


@pytest.fixture(scope="module")
def dropin(oracle_mod):
    if not oracle_mod.have_ref_dropin():
        pytest.skip("oracle/_ref/libref_dropin.so not built (it is compiled from /root/reference in the build container)")
    return oracle_mod.ref_dropin_lib()


def _events():
    return G["ev_t"], G["ev_x"].astype(np.float64), G["ev_y"].astype(np.float64), G["ev_p"]


@pytest.mark.parametrize("fit_circle", [0, 1])
def test_reference_extract_features_through_the_product_dbscan(oracle_mod, dropin, fit_circle):
    t, x, y, p = _events()
    found = 0
    for i, w in enumerate(G["windows"]):
        r = oracle_mod.ref_extract(t, x, y, p, float(w[0]), float(w[1]), 346, 260, fit_circle, lib=dropin)
        assert r["found"] == bool(G["extract_found_%d_%d" % (i, fit_circle)])
        reached = bool(G["extract_reached_%d_%d" % (i, fit_circle)])
        assert (r["cand_f32"] is not None) == reached
        if reached:
            np.testing.assert_array_equal(r["cand_f32"], G["extract_cand_%d_%d" % (i, fit_circle)])
        if r["found"]:
            found += 1
            np.testing.assert_array_equal(r["features"], G["extract_features_%d_%d" % (i, fit_circle)])
    assert found >= 1


@pytest.mark.parametrize("fit_circle", [0, 1])
def test_reference_rectify_features_through_the_product_dbscan(oracle_mod, dropin, fit_circle):
    """rectifyFeatures expands inliers to whole DBSCAN clusters (pClusters_ / nClusters_, CirclesEventFrame.cpp:522-553): the
    cluster lists the product's DBSCAN::Run handed back must carry the reference's refit to the same circles."""
    t, x, y, p = _events()
    for i, w in enumerate(G["windows"]):
        rc, out, fid = oracle_mod.ref_rectify(t, x, y, p, float(w[0]), float(w[1]), 346, 260, fit_circle,
                                              G["rectify_img_%d_%d" % (i, fit_circle)], G["rectify_fxy_%d_%d" % (i, fit_circle)],
                                              lib=dropin)
        assert rc == int(G["rectify_rc_%d_%d" % (i, fit_circle)])
        np.testing.assert_array_equal(out, G["rectify_out_%d_%d" % (i, fit_circle)])
        np.testing.assert_array_equal(fid, G["rectify_fid_%d_%d" % (i, fit_circle)])
