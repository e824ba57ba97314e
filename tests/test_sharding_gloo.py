"""N > 1 host logic on CPU (gloo, world_size 2): window sharding has no gaps/overlaps, and the sharded cost
evaluation + ONE sum all-reduce of the packed normal equations reproduces the single-process result, after which
the replicated host LM (ecb_lm_*) proposes bit-identical steps on every rank."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_window_shard_partitions():
    from eventcalib_b200.sharding import window_shard, event_range_for_windows
    for n in (0, 1, 7, 8, 6666):
        for world in (1, 2, 3, 8):
            blocks = [window_shard(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            assert max(b[1] - b[0] for b in blocks) - min(b[1] - b[0] for b in blocks) <= 1
    t = np.arange(100) * 1e-3
    w = np.array([[0.010, 0.0199], [0.020, 0.0299]])
    assert event_range_for_windows(t, w) == (10, 30)


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import oracle
    import eventcalib_b200 as ecb
    from eventcalib_b200 import synth, calib_problem, sharding
    from test_lm_host import _packed
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ev = synth.make_stream(40000, 346, 260, t0=5.0, duration=0.3, seed=21, return_truth=True, rot_amp=(0.3, 0.3, 0.2), dist=90.0)
    pb = calib_problem.build(ev, seed=1)
    x = (pb["intrinsics"], pb["rot_cp"], pb["trans_cp"])
    # shard the events by time: rank r evaluates only its slice's residuals
    lo, hi = sharding.window_shard(len(ev["t"]), rank, world)
    P = oracle.CostProblem([pb["n_cp"]], [pb["knots"]], pb["radius"], pb["huber"])
    P.associate(ev["t"][lo:hi], ev["x"][lo:hi], ev["y"][lo:hi], pb["kf_t"], pb["circles"], pb["landmarks"], pb["step"])
    packed = torch.from_numpy(_packed(*P.normal_eq(*x)))
    sharding.allreduce_normal_equations(packed)
    full = oracle.CostProblem([pb["n_cp"]], [pb["knots"]], pb["radius"], pb["huber"])
    full.associate(ev["t"], ev["x"], ev["y"], pb["kf_t"], pb["circles"], pb["landmarks"], pb["step"])
    ref = _packed(*full.normal_eq(*x))
    err = float(np.abs(packed.numpy() - ref).max() / np.abs(ref).max())
    lm = ecb.LmState([pb["n_cp"]], ecb.lm_options(max_iterations=3))
    lm.begin(*x, packed.numpy())
    st, ci, cr, ct = lm.propose()
    cand = torch.from_numpy(np.concatenate([ci, cr.ravel(), ct.ravel()]))
    gathered = [torch.zeros_like(cand) for _ in range(world)]
    dist.all_gather(gathered, cand)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    if rank == 0:
        np.save(out, np.array([err, float(same), float(st)]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_allreduce_matches_single_process(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "res.npy")
    port = 29500 + (os.getpid() % 400)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    err, same, st = np.load(out)
    assert err < 1e-12      # all-reduce of the shards == single-process normal equations (up to summation order)
    assert same == 1.0      # replicated LM: every rank proposes the identical candidate
    assert st == 0
