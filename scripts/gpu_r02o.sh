#!/bin/bash
set -u
OUT=gpurun_out
TAG=${1:-r02o}
mkdir -p $OUT
for e in 1 0; do
  for s in 8 2; do
    WFB_TEAM_EAGER=$e WFB_TRACE=1 timeout 300 python scripts/c3_sample_align.py C3 $s > $OUT/${TAG}_e${e}_s$s.log 2> $OUT/${TAG}_e${e}_s$s.err
    echo "eager=$e stride=$s: $(tail -1 $OUT/${TAG}_e${e}_s$s.log | cut -c1-80) $(grep 'persist ctas' $OUT/${TAG}_e${e}_s$s.err | tail -1 | cut -c1-60)"
  done
done
timeout 300 python scripts/c3_align_profile.py C3 32768:1 32768:1 > $OUT/${TAG}_c3.log 2> $OUT/${TAG}_c3.err; echo "c3 rc=$?"; cat $OUT/${TAG}_c3.log; grep "persist ctas\|helper" $OUT/${TAG}_c3.err | tail -2
