"""The C++ host side over the C ABI: CLI argument / error behaviour (CPU) and, on the GPU, the façade self-test and a
full CLI run on a synthetic .bin with the reference's config keys (parameter/event_calibration/example.yaml)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "eventcalib_b200", "host")
CLI = os.path.join(HOST, "unit_test_eventCameraCalib")

YAML = """%YAML:1.0
# same keys as the reference's parameter/event_calibration/example.yaml
StartTime: 5
EndTime: 10
MotionTimeStep: 5e-4
FrameEventNumThreshold: 4000
Camera.width: 346
Camera.height: 260
Is_Pattern_Asymmetric: 1
BoardSize_Rows: 9
BoardSize_Cols: 4
Square_Size: 5.5
Circles_Radius: 1.75
Calibrate_NrOfFrameToUse: 200
Calibrate_UseFisheyeModel: 0
Calibrate_FixAspectRatio: 1
Calibrate_AssumeZeroTangentialDistortion: 1
Calibrate_FixPrincipalPointAtTheCenter: 1
Fix_K1: 0
Fix_K2: 0
Fix_K3: 0
Fix_K4: 1
Fix_K5: 1
dbscan_eps: 4
dbscan_startMinSample: 2
clusterMinSample: 5
knn_num: 3
fitCircle: 0
useSO3: 0
reduceMap: 0
Viewer.Facing: [ 1,0,0,0,1,0,0,0,1 ]
"""


@pytest.fixture(scope="module")
def built():
    import eventcalib_b200.build as b
    b.build()
    subprocess.check_call(["make", "-s", "-C", HOST])
    return CLI


def test_usage_and_missing_settings(built, tmp_path):
    r = subprocess.run([built], capture_output=True, text=True)
    assert r.returncode == 1 and "Usage: ./unit_test_eventCameraCalib settingFilePath binFilePath SavePath" in r.stderr
    r = subprocess.run([built, "a", "b"], capture_output=True, text=True)
    assert r.returncode == 1
    r = subprocess.run([built, str(tmp_path / "nope.yaml"), "x.bin", str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 255 and "Failed to open settings file at:" in r.stderr   # exit(-1), eventCameraCalib.cpp:115-118


@pytest.mark.gpu
def test_facade_selftest(built):
    r = subprocess.run([os.path.join(HOST, "test_facade")], capture_output=True, text=True)
    assert r.returncode == 0 and "facade ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_cli_on_synthetic_stream(built, tmp_path):
    from eventcalib_b200 import synth
    ev = synth.make_stream(500000, 346, 260, t0=5.0, duration=0.25, seed=1001)
    # a few events before StartTime must be skipped by the loader
    pre = synth.make_stream(1000, 346, 260, t0=4.0, duration=0.01, seed=1)
    full = {k: np.concatenate([pre[k], ev[k]]) for k in "txyp"}
    synth.write_bin(str(tmp_path / "ev.bin"), full)
    (tmp_path / "cfg.yaml").write_text(YAML)
    save = tmp_path / "out"
    env = dict(os.environ, ECB_PIECES="3")
    r = subprocess.run([built, str(tmp_path / "cfg.yaml"), str(tmp_path / "ev.bin"), str(save)], capture_output=True, text=True,
                       stdin=subprocess.DEVNULL, timeout=600, env=env)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr
    assert "Events from 5 second to" in r.stdout and "frames in Map." in r.stdout and "press Enter to exit..." in r.stdout
    frames = int([l for l in r.stdout.splitlines() if l.endswith("frames in Map.")][0].split()[0])
    assert frames >= 15
    cand = np.loadtxt(str(save / "candidates.txt"))
    assert cand.shape[1] == 5 and len(np.unique(cand[:, 0])) == frames
    assert np.all((cand[:, 4] > 0) & (cand[:, 4] < 16))      # circle radii in pixels (< circleRadiusThreshold)
    # every frame holds the 36 circles in board order
    assert np.all(np.bincount(np.unique(cand[:, 0], return_inverse=True)[1]) == 36)
    # the adaptive window loop (eventCameraCalib.cpp:49-81) replayed window by window through the Python binding
    import eventcalib_b200 as ecb
    from test_circles_grid import _lib as grid_lib, _order as grid_order
    glib = grid_lib()
    ctx = ecb.Context(0)
    ctx.set_sensor(346, 260)
    ctx.load_events(synth.to_records(ev))
    rthr = ecb.radius_threshold(346, 260, 9, 4, True, 5.5, 1.75)
    prm = ecb.default_params(fit_circle=0, radius_threshold=rthr, order_mode=1, median_mode=1)
    step, t_end = 5e-4, float(ev["t"][-1])
    ln, gap, pstep = 3 * step, 5 * step, (t_end - 5.0) / 3
    import gate_py
    gate = gate_py.Gate(9, 4, step)
    stamps = []
    pieces = []
    for k in range(3):
        lo, hi = t_end - pstep * (k + 1), t_end - pstep * k
        pieces.append([lo, lo + ln, hi])
    while any(b < hi for a, b, hi in pieces):          # the CLI's wavefront: one window per unfinished piece per round
        for pc in pieces:
            a, b, hi = pc
            if not b < hi:
                continue
            ctx.frontend_run(np.array([[a, b]]), prm)
            s = ctx.summary()[0]
            n_ev = int(s["n_points"].sum())
            ok = False
            if s["n_candidates"] >= 36:   # extractFeatures(): candidates found and ordered as the 9 x 4 grid
                cpts = ctx.candidates(128)[0, :int(s["n_candidates"]), 2:4]
                ok, order36 = grid_order(glib, cpts.astype(np.float32).astype(np.float64))
            if ok:   # tracking->process (eventCameraCalib.cpp:60): EventCalibIni::track on the ordered centres
                ok = gate.process((a + b) / 2, cpts[order36])
            if ok:
                stamps.append((a + b) / 2)
                a = b + gap
                b = a + ln
            elif n_ev > 4000 or (b - a) > 3 * ln:
                a += step
                b = a + ln
            else:
                b += step
            pc[0], pc[1] = a, b
    ctx.close()
    assert len(stamps) == frames
    np.testing.assert_allclose(np.sort(stamps), np.unique(cand[:, 0]), rtol=0, atol=1e-12)

    # the back half ran through (its results are checked on a well-posed stream in the next test)
    assert "frames in Map after Initialization." in r.stdout and "Intrinsics after optimization:" in r.stdout
    assert (save / "TrajectoryByEvent.txt").exists()


@pytest.mark.gpu
def test_cli_end_to_end_calibration(built, tmp_path):
    """The whole CLI (eventCameraCalib.cpp:104-233) on a synthetic stream whose camera orbits the board with +-0.35 rad tilts
    (the calibration is well posed), fitCircle 1: initialisation, spline optimisation on the GPU, trajectory file — checked
    against the generator's ground-truth camera and trajectory."""
    from eventcalib_b200 import synth
    ev = synth.make_stream(2000000, 346, 260, t0=5.0, duration=1.0, seed=1001, return_truth=True, workers=4,
                           rot_amp=(0.35, 0.35, 0.3), orbit=True)
    truth_cam, truth_traj = ev["camera"], ev["trajectory"]
    synth.write_bin(str(tmp_path / "ev.bin"), ev)
    (tmp_path / "cfg.yaml").write_text(YAML.replace("fitCircle: 0", "fitCircle: 1"))
    save = tmp_path / "out"
    env = dict(os.environ, ECB_PIECES="4")
    r = subprocess.run([built, str(tmp_path / "cfg.yaml"), str(tmp_path / "ev.bin"), str(save)], capture_output=True, text=True,
                       stdin=subprocess.DEVNULL, timeout=600, env=env)
    print("\n".join(l for l in r.stdout.splitlines() if not l.startswith("Frame ")))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr
    frames = int([l for l in r.stdout.splitlines() if l.endswith("frames in Map.")][0].split()[0])

    # ---- the back half: initialisation (EventCalibIni::cvCalibration), spline optimisation, trajectory file ----
    out = r.stdout
    line = lambda key: [l for l in out.splitlines() if key in l][0]
    assert "Calibration succeeded" in out
    rms = float(line("Re-projection error reported by calibrateCamera:").split(":")[1])
    assert 0 < rms < 1.0                                      # centres of fitted circles: sub-pixel
    kept = int(line("frames in Map after Initialization.").split()[0])
    assert 10 < kept <= frames
    before = np.array(line("Intrinsics before optimization:").split(":")[1].split(), float)
    after = np.array(line("Intrinsics after optimization:").split(":")[1].split(), float)
    truth = truth_cam.intrinsics()
    assert before.shape == (9,) and after.shape == (9,)
    assert before[2] == 172.5 and before[3] == 129.5 and before[0] == before[1]   # CALIB_FIX_PRINCIPAL_POINT / FIX_ASPECT_RATIO
    assert abs(before[0] / truth[0] - 1) < 2e-2                # the frame-based initialisation
    # the event-based optimisation lands on the ground-truth camera of the generator: focal lengths within 1 %,
    # principal point within 2 px, and it is closer than the frame-based initialisation was
    assert abs(after[0] / truth[0] - 1) < 1e-2 and abs(after[1] / truth[1] - 1) < 1e-2, (before, after, truth)
    assert abs(after[2] - truth[2]) < 2 and abs(after[3] - truth[3]) < 2
    # The distortion terms after[4:9] (inverse radial polynomial k1 .. k5) are NOT asserted: on a 1 s stream with the board near
    # the image centre only their combined effect on the rim points is observable, the individual coefficients of r^4 .. r^10
    # are determined to no better than their own magnitude (the LM stops with k2 .. k5 far from the generator's values while
    # the reprojection cost and the poses below agree with the ground truth; profiles/r1e_cli_c1_fitcircle0.log).  Whether
    # Ceres would stop elsewhere on the same data cannot be known here (a12's solve is unpinned).  What is
    # checked instead is the thing they model: the undistortion they encode, at the pixels the rims occupy.
    x = np.linspace(-0.35, 0.35, 15)               # normalised radius range covered by the board in this stream
    r2 = x * x
    s_after = 1 + sum(after[4 + k] * r2 ** (k + 1) for k in range(5))
    s_truth = 1 + sum(truth[4 + k] * r2 ** (k + 1) for k in range(5))
    assert np.max(np.abs(s_after / s_truth - 1)) < 5e-3
    summ = line("Solver Summary:")
    c0, c1 = (float(x) for x in summ.split("cost")[1].split(",")[0].split("->"))
    assert c1 < c0
    # TUM trajectory (SystemBase.cpp:122-150): one line per key frame, pose = ground truth within 1 cm / 0.5 degree
    tum = np.loadtxt(str(save / "TrajectoryByEvent.txt"))
    assert tum.shape == (kept, 8)
    Rw, tw = truth_traj.pose(tum[:, 0])
    assert np.abs(tum[:, 1:4] - tw).max() < 1.0, np.abs(tum[:, 1:4] - tw).max()
    from scipy.spatial.transform import Rotation as Rot
    ang = (Rot.from_quat(tum[:, 4:8]) * Rot.from_matrix(Rw).inv()).magnitude()
    assert np.degrees(ang).max() < 0.5, np.degrees(ang).max()


def _run_cli(built, cfg, binf, save, devices, pieces="6"):
    env = dict(os.environ, ECB_PIECES=pieces, ECB_DEVICES=devices)
    env.pop("ECB_DEVICE", None)
    r = subprocess.run([built, str(cfg), str(binf), str(save)], capture_output=True, text=True, stdin=subprocess.DEVNULL,
                       timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr
    return r


@pytest.mark.gpu
def test_cli_multi_gpu_equals_single_gpu(built, tmp_path):
    """The C++ multi-GPU host (include/ecb/multi_gpu.hpp): the reference's time pieces (eventCameraCalib.cpp:172-180) dealt out
    to GPU shards, one host thread + context per shard.  Listing device 0 three times exercises the whole sharding path —
    record partition at piece boundaries, window routing, per-shard threads, merge, sharded rectifyFeatures — on a one-GPU
    box: frames, candidate circles, the initialisation report and the final calibration equal the single-context run BIT FOR
    BIT (same text, same files)."""
    from eventcalib_b200 import synth
    ev = synth.make_stream(800000, 346, 260, t0=5.0, duration=0.4, seed=1001, return_truth=True, workers=4,
                           rot_amp=(0.35, 0.35, 0.3), orbit=True)
    synth.write_bin(str(tmp_path / "ev.bin"), ev)
    (tmp_path / "cfg.yaml").write_text(YAML.replace("fitCircle: 0", "fitCircle: 1"))
    one = _run_cli(built, tmp_path / "cfg.yaml", tmp_path / "ev.bin", tmp_path / "out1", "0")
    three = _run_cli(built, tmp_path / "cfg.yaml", tmp_path / "ev.bin", tmp_path / "out3", "0,0,0")
    assert "on 1 GPU(s)" in one.stderr and "on 3 GPU(s)" in three.stderr
    assert int([l for l in one.stdout.splitlines() if l.endswith("frames in Map.")][0].split()[0]) >= 10
    assert one.stdout == three.stdout
    assert (tmp_path / "out1" / "candidates.txt").read_bytes() == (tmp_path / "out3" / "candidates.txt").read_bytes()
    assert (tmp_path / "out1" / "TrajectoryByEvent.txt").read_bytes() == (tmp_path / "out3" / "TrajectoryByEvent.txt").read_bytes()


@pytest.mark.gpu
def test_cli_two_real_gpus(built, tmp_path):
    """Two distinct devices: detection and initialisation bit for bit as on one GPU; the spline optimisation runs the replicated
    device LM with the normal equations summed over NVLink — final intrinsics and trajectory within 1e-9 / 1e-8."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    from eventcalib_b200 import synth
    ev = synth.make_stream(2000000, 346, 260, t0=5.0, duration=1.0, seed=1001, return_truth=True, workers=4,
                           rot_amp=(0.35, 0.35, 0.3), orbit=True)
    synth.write_bin(str(tmp_path / "ev.bin"), ev)
    (tmp_path / "cfg.yaml").write_text(YAML.replace("fitCircle: 0", "fitCircle: 1"))
    one = _run_cli(built, tmp_path / "cfg.yaml", tmp_path / "ev.bin", tmp_path / "out1", "0", pieces="8")
    two = _run_cli(built, tmp_path / "cfg.yaml", tmp_path / "ev.bin", tmp_path / "out2", "0,1", pieces="8")
    assert "on 2 GPU(s)" in two.stderr
    head = lambda s: s.split("Solver Summary:")[0]
    assert head(one.stdout) == head(two.stdout)
    assert (tmp_path / "out1" / "candidates.txt").read_bytes() == (tmp_path / "out2" / "candidates.txt").read_bytes()
    line = lambda out, key: [l for l in out.splitlines() if key in l][0]
    a1 = np.array(line(one.stdout, "Intrinsics after optimization:").split(":")[1].split(), float)
    a2 = np.array(line(two.stdout, "Intrinsics after optimization:").split(":")[1].split(), float)
    np.testing.assert_allclose(a2[:4], a1[:4], rtol=1e-7)   # host LM on one GPU vs replicated device LM on two
    t1, t2 = np.loadtxt(str(tmp_path / "out1" / "TrajectoryByEvent.txt")), np.loadtxt(str(tmp_path / "out2" / "TrajectoryByEvent.txt"))
    np.testing.assert_allclose(t2, t1, rtol=0, atol=1e-5)


@pytest.mark.gpu
def test_cli_reference_signatures_and_images(built, tmp_path):
    """The reference's own argument lists on the façade (include/ecb/compat/): `EventCalibSpline(MapBase::Ptr, EventContainer::Ptr,
    bool useSO3, bool reduceMap, double, double)` (EventCalibSpline.hpp:19) does what the CLI's step-by-step path does — same
    report lines, same intrinsics, same trajectory; `CirclesEventFrame::rectifyFeatures(const std::unordered_set<int>&, Rcw, tcw)`
    and `findCenter(const Eigen::Vector2d&) -> LandmarkBase::Ptr` (CirclesEventFrame.hpp:43-65) reproduce the batched result;
    and SavePath/image/<timestamp>.png exists for every key frame (eventCameraCalib.cpp:214-227)."""
    from eventcalib_b200 import synth
    ev = synth.make_stream(1000000, 346, 260, t0=5.0, duration=0.5, seed=1001, return_truth=True, workers=4,
                           rot_amp=(0.35, 0.35, 0.3), orbit=True)
    synth.write_bin(str(tmp_path / "ev.bin"), ev)
    (tmp_path / "cfg.yaml").write_text(YAML.replace("fitCircle: 0", "fitCircle: 1"))
    a = _run_cli(built, tmp_path / "cfg.yaml", tmp_path / "ev.bin", tmp_path / "outA", "0")
    env = dict(os.environ, ECB_PIECES="6", ECB_DEVICES="0", ECB_REFERENCE_SIGNATURES="1")
    b = subprocess.run([built, str(tmp_path / "cfg.yaml"), str(tmp_path / "ev.bin"), str(tmp_path / "outB")], capture_output=True,
                       text=True, stdin=subprocess.DEVNULL, timeout=900, env=env)
    assert b.returncode == 0, b.stdout[-2000:] + b.stderr
    keep = lambda out: [l for l in out.splitlines() if not l.startswith("Solver Summary:")]
    assert keep(a.stdout) == keep(b.stdout)                      # report lines incl. the intrinsics before / after, 12 digits
    ta, tb = np.loadtxt(str(tmp_path / "outA" / "TrajectoryByEvent.txt")), np.loadtxt(str(tmp_path / "outB" / "TrajectoryByEvent.txt"))
    np.testing.assert_allclose(tb, ta, rtol=0, atol=2e-10)       # written with 10 decimals
    msg = [l for l in b.stderr.splitlines() if l.startswith("reference signatures:")][0]
    assert "rectifyFeatures(outlierIdxs, Rcw, tcw) true" in msg
    n_feat, n_alive = int(msg.split("true, ")[1].split()[0]), int(msg.split(" of ")[1].split()[0])
    assert n_feat == n_alive >= 29
    assert float(msg.split("batched result ")[1].split(",")[0]) < 1e-5    # same circles (image points pass through float)
    assert int(msg.split("landmark for ")[1].split()[0]) == n_feat
    # per-key-frame debug images
    kept = int([l for l in a.stdout.splitlines() if l.endswith("frames in Map after Initialization.")][0].split()[0])
    pngs = sorted((tmp_path / "outA" / "image").glob("*.png"))
    assert len(pngs) == kept
    import cv2
    im = cv2.imread(str(pngs[0]))
    assert im.shape == (260, 346, 3)
    assert (im.reshape(-1, 3) == (255, 255, 255)).all(1).sum() > 200    # white rectified circles / median marks
    assert ((im[:, :, 2] == 200) | (im[:, :, 2] == 100)).sum() > 300    # cluster pixels (R = 200 positive, 100 negative)
