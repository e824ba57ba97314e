#!/bin/bash
# A/B of library variants (profiles/tools/ab_build.py): bench.py C2 device-resident + e2e, one digest line per variant.
#   profiles/tools/ab_run2.sh <out.txt> [bench args --] <tag> [<tag> ...]     ("base" = eventcalib_b200/libecb.so)
out=$1; shift
args=""
if [[ " $* " == *" -- "* ]]; then while [ "$1" != "--" ]; do args="$args $1"; shift; done; shift; fi
mkdir -p $(dirname $out); : > $out
for tag in "$@"; do
  lib=profiles/bin/libecb_$tag.so
  [ "$tag" = base ] && lib=eventcalib_b200/libecb.so
  ECB_LIBRARY=$PWD/$lib python bench.py --steps 5 --warmup 3 --no-cpu $args 2>/dev/null | python profiles/tools/brief.py $tag | tee -a $out
done
