#!/bin/bash
# A/B of library variants built by ab_build.py: runs bench.py (C2, device-resident stage times) once per variant.
#   profiles/tools/ab_run.sh <out.jsonl> <tag> [<tag> ...]     ("base" = the regular eventcalib_b200/libecb.so)
out=$1; shift
: > $out
for tag in "$@"; do
  lib=profiles/bin/libecb_$tag.so
  [ "$tag" = base ] && lib=eventcalib_b200/libecb.so
  ECB_LIBRARY=$PWD/$lib python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(json.dumps({'variant':'$tag','ms_per_step':d['ms_per_step'],'e2e_ms':d['e2e']['ms_per_step'],'stages':{k:round(v['ms'],4) for k,v in d['roofline']['kernels'].items()}}))" | tee -a $out
done
