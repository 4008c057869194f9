#!/bin/bash
set -u
OUT=gpurun_out
TAG=${1:-r02r}
mkdir -p $OUT
for v in main team2k chunk512; do
  if [ $v = main ]; then unset WFB_LIB; else export WFB_LIB=scripts/_build/libwfb_$v.so; fi
  echo "== $v"
  timeout 300 python scripts/r01_workloads.py 2>&1 | tee $OUT/${TAG}_${v}_r01.log | cut -c1-200
  WFB_TRACE=1 timeout 300 python scripts/c3_sample_align.py C3 8 > $OUT/${TAG}_${v}_s8.log 2> $OUT/${TAG}_${v}_s8.err
  echo "stride 8: $(tail -1 $OUT/${TAG}_${v}_s8.log | cut -c1-90) $(grep 'persist ctas' $OUT/${TAG}_${v}_s8.err | tail -1 | cut -c1-60)"
done
