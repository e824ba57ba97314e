#!/usr/bin/env python
"""Wall-clock of the drop-in CLI (eventcalib_b200/host/unit_test_eventCameraCalib) on BASELINE config C1: a synthetic
1 M-event DAVIS346 circle-grid .bin (0.5 s, camera orbiting the board so that the calibration is well posed) with the keys of
parameter/event_calibration/example.yaml.  Prints the CLI's report lines and the wall time of three runs (process start, CUDA
context creation, file read, window loop, initialisation, spline optimisation, trajectory file)."""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from eventcalib_b200 import synth  # noqa: E402
import eventcalib_b200.build as b  # noqa: E402
from test_host_cli import YAML  # noqa: E402

b.build()
host = os.path.join(ROOT, "eventcalib_b200", "host")
subprocess.check_call(["make", "-s", "-C", host])
fit = sys.argv[1] if len(sys.argv) > 1 else "0"
with tempfile.TemporaryDirectory() as d:
    ev = synth.make_stream(1000000, 346, 260, t0=5.0, duration=0.5, seed=1001, return_truth=True, workers=4,
                           rot_amp=(0.35, 0.35, 0.3), orbit=True)
    synth.write_bin(os.path.join(d, "ev.bin"), ev)
    open(os.path.join(d, "cfg.yaml"), "w").write(YAML.replace("fitCircle: 0", "fitCircle: " + fit))
    for run in range(3):
        t0 = time.perf_counter()
        r = subprocess.run([os.path.join(host, "unit_test_eventCameraCalib"), os.path.join(d, "cfg.yaml"), os.path.join(d, "ev.bin"),
                            os.path.join(d, "out")], capture_output=True, text=True, stdin=subprocess.DEVNULL)
        dt = time.perf_counter() - t0
        if run == 0:
            print("\n".join(l for l in r.stdout.splitlines() if not l.startswith("Frame ")))
            print(r.stderr.strip())
            print("ground truth intrinsics:", " ".join("%.9g" % v for v in ev["camera"].intrinsics()))
        print("run %d: rc %d, wall %.3f s (fitCircle %s, 1 M events)" % (run, r.returncode, dt, fit))
