#!/usr/bin/env python
"""A/B builds of libecb.so: recompiles ONE source with extra -D flags and links it with the regular objects into
profiles/bin/libecb_<tag>.so.  Select a variant at run time with ECB_LIBRARY=<path> (eventcalib_b200.load_library).

    python profiles/tools/ab_build.py ecb_cluster.cu noagg -DECB_CL_AGG=0
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from eventcalib_b200 import build as b  # noqa: E402

src, tag, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
b.build()
out_dir = os.path.join(ROOT, "profiles", "bin")
os.makedirs(out_dir, exist_ok=True)
obj = os.path.join(out_dir, "%s_%s.o" % (src[:-3], tag))
subprocess.check_call([b.NVCC] + b.COMMON + b.PER_FILE.get(src, []) + flags + ["-c", os.path.join(b.CSRC, src), "-o", obj])
objs = [obj if s == src else os.path.join(b.OBJ, s[:-3] + ".o") for s in b.sources()]
out = os.path.join(out_dir, "libecb_%s.so" % tag)
subprocess.check_call([b.NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs)
print(out)
