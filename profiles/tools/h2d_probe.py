"""PCIe probe for the end-to-end pipeline of bench.py: pinned -> device copy rate of the 500 MB record buffer as one copy,
as 8 concurrent per-stream copies, and as 8 per-stream copies while kernels keep the SMs / HBM busy."""
import time
import torch

n = 500_000_000
pinned = torch.empty(n, dtype=torch.uint8, pin_memory=True)
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
big = torch.empty(1 << 28, dtype=torch.float32, device="cuda")
streams = [torch.cuda.Stream() for _ in range(8)]
side = torch.cuda.Stream()


def run(split, busy):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if busy:
        with torch.cuda.stream(side):
            for _ in range(busy):
                big.mul_(1.0001)
    step = n // split
    ends = []
    for j in range(split):
        with torch.cuda.stream(streams[j % 8]):
            dev[j * step:(j + 1) * step].copy_(pinned[j * step:(j + 1) * step], non_blocking=True)
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            ends.append(e)
    for s in streams:
        s.synchronize()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return n / (t1 - t0) / 1e9


for split, busy in ((1, 0), (8, 0), (8, 40), (64, 0), (64, 40)):
    r = [run(split, busy) for _ in range(4)]
    print("copies %2d  concurrent kernels %2d : %.1f GB/s (best of 4: %s)" % (split, busy, max(r), " ".join("%.1f" % v for v in r)))
