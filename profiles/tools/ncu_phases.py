#!/usr/bin/env python
"""Per-phase breakdown of k_cluster from an `ncu --set full --import-source on` report: instructions, stall samples (= time
share: every resident warp is sampled whether it issues or waits), shared-memory wavefronts and the barrier / short-scoreboard
share of the samples, aggregated over the source lines between the `// ---- N.` phase markers of ecb_cluster.cu.

    cuobjdump -xelf all eventcalib_b200/libecb.so && nvdisasm -g -c ecb_cluster.sm_100a.cubin > cluster.sass
    ncu -i rep.ncu-rep --page source --csv --kernel-name regex:k_cluster > cl_src.csv
    python profiles/tools/ncu_phases.py cl_src.csv cluster.sass k_clusterItLb1 [--lines]
"""
import csv
import os
import re
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
src_csv, sass, func = sys.argv[1:4]
src = open(os.path.join(ROOT, "eventcalib_b200", "csrc", "ecb_cluster.cu")).read().splitlines()
marks = [(1, "helpers (union-find)")]
for i, l in enumerate(src, 1):
    m = re.match(r"\s*// ---- (\d+\w*\. .*?) -{3,}\s*$", l)
    if m:
        marks.append((i, m.group(1).strip()))
    elif "__global__ void" in l and "k_cluster" in l:
        marks.append((i, "prologue / work fetch"))
    elif re.match(r"\s*// ---- header", l):
        marks.append((i, "header"))
marks.sort()


def phase(line):
    name = marks[0][1]
    for ln, nm in marks:
        if ln <= line:
            name = nm
    return name


lines = open(sass).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and func in l and l.rstrip().endswith(":"))
cur, lastmain, ins = None, None, []
for l in lines[start + 1:]:
    if l.startswith("//-----") and ".text." in l:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        if cur[0] == "ecb_cluster.cu":
            lastmain = cur  # inlined helpers of other files are charged to the call site
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        ins.append(lastmain)
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
col = {k: hdr.index(k) for k in ("Instructions Executed", "# Samples", "L1 Wavefronts Shared Excessive", "L1 Wavefronts Shared",
                                 "stall_barrier", "stall_short_sb")}
body = [r for r in rows[2:] if len(r) > col["# Samples"]]
assert len(body) == len(ins), (len(body), len(ins))
agg = defaultdict(lambda: [0] * 6)
for r, lm in zip(body, ins):
    key = ("%s:%d" % lm if "--lines" in sys.argv else phase(lm[1])) if lm else "other"
    for j, k in enumerate(col):
        agg[key][j] += int(r[col[k]] or 0)
ti, ts, tw = (sum(v[j] for v in agg.values()) for j in (0, 1, 3))
print("k_cluster: %d warp instructions, %d stall samples, %d shared-memory wavefronts (%d excessive)\n" % (
    ti, ts, tw, sum(v[2] for v in agg.values())))
print("| phase | time (samples) | instructions | smem wavefronts | of which excessive | barrier share | short-scoreboard share |")
print("|---|---|---|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print("| %s | %.1f %% | %.1f %% | %.1f %% | %.1f %% | %.0f %% | %.0f %% |" % (
        k, 100 * v[1] / ts, 100 * v[0] / ti, 100 * v[3] / tw, 100 * v[2] / tw, 100 * v[4] / max(v[1], 1), 100 * v[5] / max(v[1], 1)))
