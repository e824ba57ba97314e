#!/usr/bin/env python
"""Aggregate an ncu source-page (SASS) CSV per CUDA source line, using nvdisasm -g -c line markers.

  cuobjdump -xelf all eventcalib_b200/libecb.so ; nvdisasm -g -c ecb_cluster.sm_100a.cubin > cluster.sass
  ncu -i rep.ncu-rep --page source --csv --kernel-name regex:k_cluster > src.csv
  python profiles/tools/ncu_by_line.py src.csv cluster.sass k_clusterItE [top]
"""
import csv
import re
import sys
from collections import defaultdict

src_csv, sass, func = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
lines = open(sass).read().splitlines()
# locate function body
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and func in l and l.rstrip().endswith(":"))
cur = None
ins_line = []
for l in lines[start + 1:]:
    if l.startswith("//-----") and ".text." in l:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        ins_line.append(cur)
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
body = [r for r in rows[2:] if len(r) > isamp]
if len(body) != len(ins_line):
    print("warning: %d SASS rows in report vs %d in disassembly" % (len(body), len(ins_line)))
agg = defaultdict(lambda: [0, 0])
for r, ln in zip(body, ins_line):
    agg[ln][0] += int(r[ii])
    agg[ln][1] += int(r[isamp])
ti = sum(v[0] for v in agg.values())
ts = sum(v[1] for v in agg.values())
print("total warp instructions %d, samples %d" % (ti, ts))
print("%-28s %14s %7s %9s %7s" % ("line", "warp-inst", "%", "samples", "%"))
for ln, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-28s %14d %6.2f%% %9d %6.2f%%" % ("%s:%s" % ln if ln else "?", v[0], 100.0 * v[0] / ti, v[1], 100.0 * v[1] / max(ts, 1)))
