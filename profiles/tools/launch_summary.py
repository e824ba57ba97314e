#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.

  python profiles/tools/launch_summary.py gpurun_out/rX/launches.csv [skip_first_n_launches_per_kernel]
"""
import csv
import re
import sys
from collections import OrderedDict

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = OrderedDict()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ik]).replace("<unnamed>::", "").replace("void ", "")
    agg.setdefault(name, []).append(float(r[iv]) / 1e6)
tot = sum(sum(v[skip:]) / max(len(v[skip:]), 1) * 1 for v in agg.values())
print("| kernel | launches | avg ms | share of the per-launch sum |")
print("|---|---|---|---|")
for k, v in agg.items():
    w = v[skip:] or v
    avg = sum(w) / len(w)
    print("| %s | %d | %.4f | %.1f %% |" % (k, len(v), avg, 100 * avg / tot))
