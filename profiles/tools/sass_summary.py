#!/usr/bin/env python
"""SASS opcode summary of libecb.so per kernel: which Blackwell-specific / pipe-defining instructions each kernel holds.

    python profiles/tools/sass_summary.py > profiles/r2s_sass_opcodes.md

UBLKCP = TMA bulk copy (cp.async.bulk), SYNCS = mbarrier, DMMA = FP64 tensor-pipe MMA, DFMA/DMUL/DADD = FP64 pipe,
REDUX = warp reduction unit, ATOMS / ATOMG / RED = shared / global atomics, MATCH / VOTE = warp match / ballot,
BAR = CTA barrier, LDS/STS = shared memory, LDG/STG = global memory.  UTC*MMA / LDTM (tcgen05 / TMEM) do not occur: the only
contraction of the path is FP64 and tcgen05 has no f64 kind.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "eventcalib_b200", "libecb.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
OPS = ["UBLKCP", "SYNCS", "DMMA", "DFMA", "DMUL", "DADD", "REDUX", "MATCH", "VOTE", "ATOMS", "ATOMG", "RED", "BAR", "LDS", "STS",
       "LDG", "STG", "UTCHMMA", "UTCQMMA", "LDTM"]
cur, arch = None, None
cnt = collections.OrderedDict()
tot = collections.Counter()
for line in txt.splitlines():
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch = m.group(1)
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(anonymous namespace\)::", "", cur)
        cur = re.sub(r"^void ", "", cur).split("(")[0]
        cnt[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1).split(".")[0]
        cnt[cur]["_n"] += 1
        tot["_n"] += 1
        if op in OPS:
            cnt[cur][op] += 1
            tot[op] += 1
used = [o for o in OPS if tot[o]]
print("# SASS opcode summary of eventcalib_b200/libecb.so (%s only, %d kernels, %d instructions)\n" % (arch, len(cnt), tot["_n"]))
print(__doc__.split("\n\n")[2].strip() + "\n")
print("| kernel | instr | " + " | ".join(used) + " |")
print("|---|---|" + "---|" * len(used))
for k, c in cnt.items():
    print("| `%s` | %d | " % (k, c["_n"]) + " | ".join(str(c[o]) if c[o] else "" for o in used) + " |")
print("| **total** | %d | " % tot["_n"] + " | ".join(str(tot[o]) for o in used) + " |")
