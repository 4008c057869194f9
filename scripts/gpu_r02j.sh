#!/bin/bash
set -u
OUT=gpurun_out
TAG=${1:-r02j}
mkdir -p $OUT
for v in pf0 pf5 pf2 team2k; do
  WFB_LIB=scripts/_build/libwfb_$v.so timeout 300 python scripts/c3_align_profile.py C3 32768:1 32768:1 > $OUT/${TAG}_$v.log 2> $OUT/${TAG}_$v.err; echo "$v rc=$? $(tail -1 $OUT/${TAG}_$v.log) $(grep 'main n=' $OUT/${TAG}_$v.err | tail -1 | cut -c1-70)"
done
