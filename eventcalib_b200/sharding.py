"""Host-side sharding of the path across GPUs (one process per GPU).

Detection: windows are independent (EventFrame copies its own events, EventFrame.cpp:10-36), so rank r takes a
contiguous block of windows = a contiguous byte range of the time-sorted .bin; no data-path collective.
Cost evaluation: residuals are independent given the parameters; each rank evaluates the residuals of its own
time range and ONE sum all-reduce of the packed normal equations (+ one scalar for the cost) feeds the
replicated host LM step.
"""
import numpy as np


def window_shard(n_windows, rank, world):
    """Contiguous block [lo, hi) of window indices owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_windows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def event_range_for_windows(t, windows):
    """Event index range [lo, hi) that covers every window of the block (windows: closed [t0,t1], time sorted)."""
    if len(windows) == 0:
        return 0, 0
    lo = int(np.searchsorted(t, windows[:, 0].min(), side="left"))
    hi = int(np.searchsorted(t, windows[:, 1].max(), side="right"))
    return lo, hi


def allreduce_normal_equations(packed, group=None):
    """Sum all-reduce of the packed [J^T J | J^T r | cost] buffer (torch tensor, CUDA with NCCL / CPU with gloo)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    return packed
