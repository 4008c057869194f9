#!/bin/bash
# round-2 measurement call: C3 alignment phase with the host-stage trace, phase timers, ncu source profile on a record sample
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 400 python scripts/c3_align_profile.py C3 32768:1 32768:1 > $OUT/r02b_c3_align.log 2> $OUT/r02b_c3_align.err; echo "align rc=$?"; cat $OUT/r02b_c3_align.log
WFB_LIB=scripts/_build/libwfb_timers.so timeout 400 python scripts/c3_phase_timers.py C3 1 > $OUT/r02b_phase_timers.log 2>&1; echo "timers rc=$?"; cat $OUT/r02b_phase_timers.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wfb_persist_kernel -s 1 -c 1 -f -o $OUT/r02b_persist_c3s8 python scripts/c3_sample_align.py C3 8 > $OUT/r02b_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 $OUT/r02b_ncu.log
