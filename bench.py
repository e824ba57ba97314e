#!/usr/bin/env python
"""bench.py — events/s of the EventCalib hot path (detection: ingest -> windows -> DBSCAN -> circle fit).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--events E]

Workload at every N: BASELINE.json configs[1] ("C2") per GPU — a synthetic 20 M-event DAVIS346 (346x260)
circle-grid stream, 2 Mev/s, fixed tiling windows of 1.5 ms (= 3 x MotionTimeStep), both polarities, DBSCAN
eps 4 / minPts 2, cluster filter and circle fit (fitCircle: 1).  One "step" = one pass of the whole front end
over that stream.  N > 1: windows are independent, so rank r holds the r-th 10 s slice of a longer stream
(weak scaling, no data-path collective); value = all events of all ranks / max-over-ranks device time.

  value   device-resident: the packed 25-byte records are already in HBM when the timed region starts
  e2e     through the C ABI with HOST buffers: pinned-host records -> H2D -> front end -> D2H of the per-window
          summaries and candidate circles, every step
  roofline / cpu_baseline / clocks: see DESIGN.md §Measurement.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "events/sec (detection+residual eval)"
UNIT = "events/s"
WIDTH, HEIGHT = 346, 260
WINDOW = 1.5e-3
RATE = 2.0e6  # events / s of stream time
ALGO_BYTES_PER_EVENT = 29.0  # 25 B record read + 4 B label write


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_workload(n_events, rank):
    from eventcalib_b200 import synth
    dur = n_events / RATE
    t0 = 5.0 + rank * dur
    workers = max(1, min(16, (os.cpu_count() or 2) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))
    ev = synth.make_stream(n_events, WIDTH, HEIGHT, t0=t0, duration=dur, seed=1002 + rank, workers=workers)
    win = synth.tiling_windows(t0, t0 + dur, WINDOW)
    return ev, win


def frontend_params(order_mode=1, median_mode=1):
    import eventcalib_b200 as ecb
    rthr = ecb.radius_threshold(WIDTH, HEIGHT, 9, 4, True, 5.5, 1.75)
    return ecb.default_params(eps=4.0, min_pts=2, cluster_min=5, knn_num=3, fit_circle=1, radius_threshold=rthr,
                              rows_cols=36, order_mode=order_mode, median_mode=median_mode), rthr


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.p:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def cpu_reference_rate(ev, win, rthr, budget_s=12.0, threads=None):
    """Reference CPU path on a bounded sample of the workload's windows, threaded like the reference
    (hardware_concurrency()-2 std::threads over windows, eventCameraCalib.cpp:172-190; Ceres num_threads = hw-2,
    EventCalibSpline.cpp:242):  detection = verbatim reference DBSCAN from oracle/_ref + restated glue (oracle port if
    _ref is absent);  residual evaluation = association + one dual-number (Jet<37>) Jacobian evaluation with dense
    per-span J^T J + one cost-only evaluation (oracle/ecb_oracle_cost.cpp)."""
    import oracle
    from eventcalib_b200 import calib_problem, synth
    oracle.build()
    kind = "reference" if oracle.have_ref() else "port"
    hw = os.cpu_count() or 4
    threads = threads or max(1, hw - 2)
    kw = dict(eps=4.0, minS=2, clusterMin=5, knn_num=3, fitCircle=1, Rthr=rthr, rows_cols=36, ref=True)
    # calibrate on a few windows, then size the sample for ~budget_s of wall time (detection is ~half of it)
    probe = win[: min(len(win), 4 * threads)]
    t0 = time.perf_counter()
    _, nev, _ = oracle.frontend_windows(ev["t"], ev["x"], ev["y"], ev["p"], probe, threads=threads, **kw)
    dt = max(time.perf_counter() - t0, 1e-6)
    rate = nev / dt
    n_s = int(min(len(win), max(len(probe), 0.4 * rate * budget_s / max(nev / len(probe), 1))))
    sample = win[:n_s]
    t0 = time.perf_counter()
    cand, nev, _ = oracle.frontend_windows(ev["t"], ev["x"], ev["y"], ev["p"], sample, threads=threads, **kw)
    dt_det = time.perf_counter() - t0
    # residual evaluation of the same events
    hi = int(np.searchsorted(ev["t"], sample[-1, 1], side="right"))
    cam, board = synth.Camera(WIDTH, HEIGHT), synth.Board()
    seed = int(round((ev["t"][0] - 5.0) / max(len(ev["t"]) / RATE, 1e-9)))
    pb = calib_problem.build_from_truth(cam, synth.Trajectory(1002 + seed, board, 78.0), board, float(ev["t"][0]),
                                        float(ev["t"][hi - 1]))
    P = oracle.CostProblem([pb["n_cp"]], [pb["knots"]], pb["radius"], pb["huber"])
    t0 = time.perf_counter()
    P.associate(ev["t"][:hi], ev["x"][:hi], ev["y"][:hi], pb["kf_t"], pb["circles"], pb["landmarks"], pb["step"])
    P.eval_mt(pb["intrinsics"], pb["rot_cp"], pb["trans_cp"], threads)
    dt_res = time.perf_counter() - t0
    dt = dt_det + dt_res
    return dict(value=nev / dt, unit=UNIT, cores=threads, kind=kind, seconds=dt,
                detection_events_per_s=nev / dt_det, residual_eval_events_per_s=nev / dt_res,
                sample="first %d of %d windows (%d events, %d residuals), %d std::threads of %d host cores; detection: %s; "
                       "residual eval: single-thread association + Jet<37> Jacobian evaluation + cost-only evaluation (port)" % (
                    n_s, len(win), nev, P.n_residuals, threads, hw,
                    "verbatim reference DBSCAN + restated glue (oracle/_ref)" if kind == "reference" else "oracle port")), cand


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path, same config / metric."""
    if rank != 0:
        return
    n_events = args.events
    ev, win = make_workload(min(n_events, 4_000_000), 0)  # a bounded slice is enough for the CPU arm
    _, rthr = frontend_params()
    vals = []
    cb = None
    for i in range(args.warmup + args.steps):
        cb, _ = cpu_reference_rate(ev, win, rthr, budget_s=max(2.0, 40.0 / max(1, args.warmup + args.steps)))
        if i >= args.warmup:
            vals.append(cb)
    v = float(np.mean([c["value"] for c in vals]))
    secs = float(np.mean([c["seconds"] for c in vals]))
    cb = dict(vals[-1])
    cb["value"] = v
    cb.pop("seconds", None)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": secs * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32 pixels / f64 fit", "data": "synthetic",
            "config": {"workload": "C2 per GPU: synthetic DAVIS346 circle-grid stream, 1.5 ms tiling windows, "
                                   "DBSCAN eps 4 minPts 2 + circle fit; CPU arm runs a bounded sample of it"},
            "cpu_baseline": cb, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print_json(line)


def build_cost_problem(ev_truth, world, n_events):
    """Global calibration problem: one spline segment per rank's time slice (EventCalibSpline splits the map into
    segments at gaps, EventCalibSpline.cpp:319-348); every rank knows all segments, holds only its own residuals."""
    from eventcalib_b200 import calib_problem, synth
    cam, board = ev_truth["camera"], ev_truth["board"]
    dur = n_events / RATE
    segs = []
    for r in range(world):
        traj = synth.Trajectory(1002 + r, board, 78.0)
        segs.append(calib_problem.build_from_truth(cam, traj, board, 5.0 + r * dur + 0.5 / RATE, 5.0 + (r + 1) * dur - 0.5 / RATE,
                                                   seed=r))
    return segs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--events", type=int, default=20_000_000)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--slices", type=int, default=0, help="time slices (contexts/streams/host threads) of the e2e pipeline")
    ap.add_argument("--order-mode", type=int, default=1, help="pid order: 0 first arrival, 1 libstdc++ unordered_set order (the reference's)")
    ap.add_argument("--median-mode", type=int, default=1, help="cluster centre: 0 canonical, 1 std::nth_element over BFS order (the reference's)")
    ap.add_argument("--slice-plan", default="", help="relative sizes of the e2e time slices, e.g. 1,2,3,3,2,1 (overrides --slices)")
    ap.add_argument("--lm-iters", type=int, default=50, help="LM iterations of the C4 side measurement (0 = skip)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE line, the result JSON: everything else that writes to file descriptor 1 (NCCL prints its
    # version banner there) is sent to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    json_out = os.fdopen(json_fd, "w")
    global print_json
    def print_json(line):
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import eventcalib_b200 as ecb
    from eventcalib_b200 import synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    ev, win = make_workload(args.events, rank)
    n = len(ev["t"])
    rec = synth.to_records(ev)
    pinned = torch.empty(n * 25, dtype=torch.uint8, pin_memory=True)
    pinned.numpy()[:] = rec.view(np.uint8).reshape(-1)
    d_raw = torch.empty(n * 25 + 16, dtype=torch.uint8, device="cuda")
    d_raw[: n * 25].copy_(pinned, non_blocking=False)
    # the PCIe ceiling of the end-to-end number: plain pinned -> device copy of the same records (outside every timed region)
    h2d_gbs = 0.0
    for _ in range(3):
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        c0.record()
        d_raw[: n * 25].copy_(pinned, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        h2d_gbs = max(h2d_gbs, n * 25 / (c0.elapsed_time(c1) * 1e-3) / 1e9)

    # one explicit (non-default) stream shared by torch (copies, NCCL, timing events) and the library's kernels
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = ecb.Context(local, stream.cuda_stream)
    ctx.set_sensor(WIDTH, HEIGHT)
    prm, rthr = frontend_params(args.order_mode, args.median_mode)

    # residual evaluation: key frames / circles / spline segments from the ground truth (host-side initialisation is
    # outside the hot path); rank r owns segment r
    segs = build_cost_problem(dict(camera=synth.Camera(WIDTH, HEIGHT), board=synth.Board()), world, n)
    mine = segs[rank]
    n_cp = [sg["n_cp"] for sg in segs]
    ctx.cost_setup(n_cp, [sg["knots"] for sg in segs], mine["radius"], mine["huber"])
    intr = mine["intrinsics"]
    rot = np.concatenate([sg["rot_cp"] for sg in segs])
    trans = np.concatenate([sg["trans_cp"] for sg in segs])
    ctx.load_events_device(d_raw.data_ptr(), n)
    n_res = ctx.cost_associate(mine["kf_t"], mine["circles"], mine["landmarks"], mine["step"])
    lay = ctx.cost_layout()
    d_ne = torch.zeros(lay["out_doubles"], dtype=torch.float64, device="cuda")
    h_ne = torch.empty(lay["out_doubles"], dtype=torch.float64, pin_memory=True)
    d_cost = torch.zeros(1, dtype=torch.float64, device="cuda")

    # N > 1: the inter-GPU sum of the normal equations is fused into the span reduction (peer-mapped receive buffers over
    # NVLink, exchanged once through CUDA IPC); NCCL all-reduce is the fallback when peer mapping is unavailable
    xch = None
    if world > 1 and not os.environ.get("ECB_NO_P2P"):
        ok = 1
        try:
            my_buf = ctx.device_alloc(ctx.exchange_buffer_bytes(world))
            h = torch.from_numpy(ctx.ipc_export(my_buf)).cuda()
            hs = [torch.empty_like(h) for _ in range(world)]
            dist.all_gather(hs, h)
            ptrs = [my_buf if r == rank else ctx.ipc_open(hs[r].cpu().numpy()) for r in range(world)]
            xch = {"ptrs": ptrs, "epoch": 0, "buf": my_buf}
        except Exception as e:  # noqa: BLE001
            sys.stderr.write("rank %d: peer mapping unavailable (%s), using NCCL\n" % (rank, e))
            ok = 0
        t_ok = torch.tensor([ok], device="cuda")
        dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
        if int(t_ok.item()) == 0:
            xch = None

    def normal_eq_all_ranks(i_, r_, t_):
        """packed J^T J / J^T r / cost of ALL ranks' residuals in d_ne (device)"""
        if xch:
            xch["epoch"] += 1
            ctx.cost_normal_eq_exchange(i_, r_, t_, rank, xch["ptrs"], xch["epoch"], d_ne.data_ptr(), want_cost=False)
        else:
            ctx.cost_normal_eq(i_, r_, t_, d_out=d_ne.data_ptr(), host=False)
            if world > 1:
                dist.all_reduce(d_ne)

    xch_check = None
    if xch:  # once: the fused exchange against NCCL's all-reduce of the same evaluation
        normal_eq_all_ranks(intr, rot, trans)
        a_x = d_ne.clone()
        ctx.cost_normal_eq(intr, rot, trans, d_out=d_ne.data_ptr(), host=False)
        dist.all_reduce(d_ne)
        xch_check = float((a_x - d_ne).abs().max().item() / max(float(d_ne.abs().max().item()), 1e-300))

    def residual_eval():
        """one LM iteration's worth of evaluation: association + J^T J / J^T r (+ inter-GPU sum) + cost-only"""
        ctx.cost_associate(mine["kf_t"], mine["circles"], mine["landmarks"], mine["step"])
        normal_eq_all_ranks(intr, rot, trans)
        c = ctx.cost_eval(intr, rot, trans)
        if world > 1:
            d_cost[0] = c
            dist.all_reduce(d_cost)

    def step_device():
        ctx.load_events_device(d_raw.data_ptr(), n)
        ctx.frontend_run(win, prm)
        residual_eval()

    # ---- end to end from host buffers: S time slices, each with its own context / stream / host thread, so the H2D copy
    # of slice k+1 overlaps the kernels of slice k (the C ABI is re-entrant per context; ctypes releases the GIL).
    from concurrent.futures import ThreadPoolExecutor
    from eventcalib_b200 import sharding
    # slice sizes: the pipeline is balanced (H2D time ~ kernel time), so its length is  first upload + all kernels  or
    # all uploads + last slice's kernels — short first and last slices, long ones in between (--slice-plan fractions)
    if args.slices <= 0:  # 8 slices on one GPU; with N ranks on one host keep about one host thread per core
        args.slices = 8 if world == 1 else max(2, min(8, (os.cpu_count() or 16) // world))
    plan = [float(v) for v in args.slice_plan.split(",")] if args.slice_plan else [1.0] * max(1, args.slices)
    S = len(plan)
    cuts = np.round(np.cumsum([0.0] + plan) / sum(plan) * len(win)).astype(int)
    slices = []
    for j in range(S):
        w0, w1 = int(cuts[j]), int(cuts[j + 1])
        lo = int(np.searchsorted(ev["t"], win[w0, 0], side="left")) if j > 0 else 0
        slices.append(dict(w0=w0, w1=w1, lo=lo))
    for j in range(S):
        sl = slices[j]
        sl["hi"] = slices[j + 1]["lo"] if j + 1 < S else n
        sl["ctx"] = ecb.Context(local)          # own non-blocking stream
        sl["ctx"].set_sensor(WIDTH, HEIGHT)
        sl["ctx"].cost_setup(n_cp, [sg["knots"] for sg in segs], mine["radius"], mine["huber"])
        ta, tb = ev["t"][sl["lo"]], ev["t"][sl["hi"] - 1]
        m = (mine["kf_t"] > ta - 6 * mine["step"]) & (mine["kf_t"] < tb + 6 * mine["step"])
        sl["kf_t"], sl["circles"] = mine["kf_t"][m].copy(), mine["circles"][m].copy()
    d_parts = torch.zeros(S, lay["out_doubles"], dtype=torch.float64, device="cuda")
    # caller-owned host result buffers: every slice writes its windows' summaries / candidate circles in place
    h_summ = np.zeros(len(win), ecb.SUMMARY_DTYPE)
    h_cand = np.zeros((len(win), 48, 5))
    pool = ThreadPoolExecutor(S)

    trace = os.environ.get("ECB_BENCH_TRACE")
    t_origin = [0.0]

    def slice_work(j):
        sl = slices[j]
        c = sl["ctx"]
        tm = [time.perf_counter()]
        c.load_events_ptr(pinned.data_ptr() + sl["lo"] * 25, sl["hi"] - sl["lo"])
        tm.append(time.perf_counter())
        c.frontend_run(win[sl["w0"]:sl["w1"]], prm)
        tm.append(time.perf_counter())
        c.cost_associate(sl["kf_t"], sl["circles"], mine["landmarks"], mine["step"])
        c.cost_normal_eq(intr, rot, trans, d_out=d_parts[j].data_ptr(), host=False)
        cost = c.cost_eval(intr, rot, trans)       # synchronises the slice's stream
        tm.append(time.perf_counter())
        out = c.summary(out=h_summ[sl["w0"]:sl["w1"]]), c.candidates(48, out=h_cand[sl["w0"]:sl["w1"]]), cost
        tm.append(time.perf_counter())
        if trace:
            sys.stderr.write("slice %d: start %.2f load_end %.2f frontend_end %.2f cost_end %.2f fetch_end %.2f ms\n" % (
                (j,) + tuple((x - t_origin[0]) * 1e3 for x in tm)))
        return out

    def step_e2e():
        t_origin[0] = time.perf_counter()
        res = list(pool.map(slice_work, range(S)))
        torch.sum(d_parts, dim=0, out=d_ne)
        if world > 1:
            dist.all_reduce(d_ne)
            d_cost[0] = sum(r[2] for r in res)
            dist.all_reduce(d_cost)
        h_ne.copy_(d_ne, non_blocking=False)
        return h_summ, h_cand

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(args.warmup):
        step_device()
    l0 = ctx.launches
    sampler = ClockSampler(local) if rank == 0 else None
    ms = timed(step_device, args.steps)
    launches = ctx.launches - l0
    value = world * n * args.steps / (ms * 1e-3)

    # per-kernel durations, measured live with CUDA events on the launching stream (separate pass)
    ctx.set_profiling(True)
    stage = {}
    for _ in range(args.steps):
        step_device()
        for k, v in ctx.stage_ms().items():
            stage.setdefault(k, []).append(v)
    ctx.set_profiling(False)
    stage = {k: float(np.mean(v)) for k, v in stage.items() if np.mean(v) > 0}

    # end to end through the C ABI with host buffers
    s, c = step_e2e()
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if sampler else None
    e2e_value = world * n * args.steps / (ms_e2e * 1e-3)
    h2d = n * 25 + win.nbytes + (9 + 7 * len(rot)) * 8 * 2
    d2h = s.nbytes + c.nbytes + h_ne.numel() * 8

    # C4 side measurement: fixed number of LM iterations (normal equations + all-reduce + replicated host solve)
    lm_info = None
    if args.lm_iters > 0:
        lm = ecb.LmState(n_cp, ecb.lm_options(max_iterations=args.lm_iters, fixed_iterations=1))
        barrier()
        t0 = time.perf_counter()

        def packed_at(i, r, t):
            normal_eq_all_ranks(i, r, t)
            h_ne.copy_(d_ne, non_blocking=False)
            return h_ne.numpy()

        def cost_at(i, r, t):
            cc = ctx.cost_eval(i, r, t)
            if world > 1:
                d_cost[0] = cc
                dist.all_reduce(d_cost)
                cc = float(d_cost.item())
            return cc

        st = lm.begin(intr, rot, trans, packed_at(intr, rot, trans))
        n_eval = 1
        while st == 0:
            st, ci, cr, ct = lm.propose()
            if st != 0:
                break
            fb = lm.feedback(cost_at(ci, cr, ct))
            if fb == 1:
                st = lm.update(packed_at(ci, cr, ct))
                n_eval += 1
            elif fb == 0:
                st = 0
            else:
                st = fb
        barrier()
        lm_s = time.perf_counter() - t0
        fi, _, _, summ = lm.state()
        lm_info = {"iterations": summ["iterations"], "successful_steps": summ["successful_steps"], "seconds": lm_s,
                   "s_per_iteration": lm_s / max(1, summ["iterations"]), "jacobian_evaluations": n_eval,
                   "initial_cost": summ["initial_cost"], "final_cost": summ["final_cost"], "termination": summ["termination"],
                   "residuals_all_gpus": None, "dimension": 9 + 6 * len(rot)}
        if world > 1:
            t = torch.tensor([float(n_res)], device="cuda", dtype=torch.float64)
            dist.all_reduce(t)
            lm_info["residuals_all_gpus"] = int(t.item())
        else:
            lm_info["residuals_all_gpus"] = int(n_res)

    if rank == 0:
        hbm, which = peaks()
        npts = int(s["n_points"].sum())
        # algorithmic bytes / flops per launch of each kernel (Kernels)
        algo = {"ingest": n * 37.0, "window": n * 4.0 + npts * 8.0, "cluster": npts * 12.0, "pair": npts * 8.0,
                "order": npts * 8.0,
                "assoc": n * 12.0 * 2 + n_res * 60.0, "normal_eq": n_res * 76.0, "cost": n_res * 76.0}
        flops = {"normal_eq": n_res * 2.0e3, "cost": n_res * 150.0}
        tot = sum(stage.values())
        kernels = {}
        for k in stage:
            kernels[k] = {"ms": stage[k], "share": stage[k] / tot, "algo_gbs": algo[k] / (stage[k] * 1e-3) / 1e9 if k in algo else None}
            if k in flops:
                kernels[k]["algo_tflops"] = flops[k] / (stage[k] * 1e-3) / 1e12
        dom = max((k for k in stage if k in algo), key=lambda k: stage[k])
        achieved = algo[dom] / (stage[dom] * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "k_" + dom, "achieved": achieved, "peak": hbm, "unit": "GB/s",
                    "frac": achieved / hbm, "traffic": None, "peak_source": which,
                    "algorithmic_bytes_per_launch": algo[dom],
                    "path_frac": (value / world) * ALGO_BYTES_PER_EVENT / 1e9 / hbm,
                    "note": "achieved = algorithmic bytes of the dominant kernel / its CUDA-event duration (DESIGN.md); "
                            "path_frac = whole-path 29 B/event x events/s / peak; the dominant kernels are shared-memory / "
                            "issue bound, not HBM bound (profiles/)", "kernels": kernels}
        # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/), scaled to this
        # run's event count when it differs
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            k = tr["kernels"].get("k_" + dom)
            if k:
                roofline["ncu"] = {m: k[m] for m in ("issue_active_pct", "warps_active_pct", "dram_throughput_pct", "fp64_pipe_active_pct") if m in k}
                roofline["traffic"] = (k["dram_bytes_read"] + k["dram_bytes_write"]) * (n / float(tr["events"]))
                roofline["traffic_source"] = tr["source"] + ("" if n == tr["events"] else " (scaled from %d events)" % tr["events"])
        except Exception:
            pass
        if "normal_eq" in stage:
            roofline["cost_kernel"] = {"bound": "tensor", "kernel": "k_normal_eq (FP64 DMMA)", "unit": "TFLOP/s",
                                       "achieved": flops["normal_eq"] / (stage["normal_eq"] * 1e-3) / 1e12, "peak": 37.0,
                                       "frac": flops["normal_eq"] / (stage["normal_eq"] * 1e-3) / 1e12 / 37.0,
                                       "peak_source": "measured FP64 DMMA m8n8k4 on this pool's B200 (profiles/r1_fp64_peak.txt)"}
        line = {"metric": "events/sec (detection+residual eval)", "value": value, "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "int32 pixels / f64 fit + f64 residuals", "data": "synthetic",
                "config": {"workload": "C2 per GPU: synthetic %d-event DAVIS346 (346x260) circle-grid stream, %d tiling "
                                       "windows of 1.5 ms, DBSCAN eps 4 minPts 2 + cluster filter + circle fit (fitCircle 1; pid order %s, "
                                       "cluster centres %s), then residual evaluation of the same events (association + "
                                       "J^T J/J^T r + cost, all-reduce when N>1)" % (
                                           n, len(win), "= libstdc++ unordered_set order like the reference" if args.order_mode == 1 else "= first arrival",
                                           "= std::nth_element over BFS-ordered members like the reference" if args.median_mode == 1 else "canonical"),
                           "events_per_gpu": n, "windows_per_gpu": int(len(win)), "residuals_per_gpu": int(n_res),
                           "control_points": int(len(rot)), "parallelism": "windows / spline segments sharded x%d" % world,
                           "normal_eq_exchange": (("fused into the span reduction: P2P stores into peer-mapped receive buffers "
                                                   "over NVLink, one-shot sum in rank order (max rel. diff vs NCCL all-reduce %.1e)" % xch_check)
                                                  if xch else ("NCCL all-reduce" if world > 1 else "single GPU")),
                           "l2": "inputs larger than L2 (%.0f MB of records per step)" % (n * 25 / 1e6),
                           "found_circles_per_window": float(s["n_candidates"].mean())},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
                "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": int(d2h), "pipeline": "%d time slices (relative sizes %s), one context/stream/host thread each "
                        "(H2D of slice k+1 overlaps the kernels of slice k)" % (S, ":".join("%g" % v for v in plan)),
                        "h2d_copy_gbs": h2d_gbs, "pcie_floor_ms": h2d / (h2d_gbs * 1e9) * 1e3 if h2d_gbs > 0 else None}}
        if lm_info:
            line["lm"] = lm_info
        if not args.no_cpu and world == 1:
            cb, _ = cpu_reference_rate(ev, win, rthr)
            cb.pop("seconds", None)
            line["cpu_baseline"] = cb
        print_json(line)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        if xch:
            for r_, p_ in enumerate(xch["ptrs"]):
                if r_ != rank:
                    ctx.ipc_close(p_)
            dist.barrier()
            ctx.device_free(xch["buf"])
        dist.destroy_process_group()
    pool.shutdown()
    for sl in slices:
        sl["ctx"].close()
    ctx.close()


if __name__ == "__main__":
    main()
