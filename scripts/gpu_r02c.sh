#!/bin/bash
# round-2 call c: GPU parity tests of the current build, then C3 alignment phase with and without row staging
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/r02c_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/r02c_gpu_tests.log
WFB_STAGE=0 timeout 300 python scripts/c3_align_profile.py C3 32768:1 32768:1 > $OUT/r02c_c3_nostage.log 2> $OUT/r02c_c3_nostage.err; echo "nostage rc=$?"; cat $OUT/r02c_c3_nostage.log
timeout 300 python scripts/c3_align_profile.py C3 32768:1 32768:1 > $OUT/r02c_c3_stage.log 2> $OUT/r02c_c3_stage.err; echo "stage rc=$?"; cat $OUT/r02c_c3_stage.log
grep -h "persist:\|main n=\|endsfree\|t+" $OUT/r02c_c3_stage.err | tail -32
