#!/bin/bash
set -u
OUT=gpurun_out
TAG=${1:-r02p}
mkdir -p $OUT
WFB_TRACE=stages timeout 300 python scripts/c3_align_profile.py C3 32768:1 32768:1 > $OUT/${TAG}_c3.log 2> $OUT/${TAG}_c3.err; echo "c3 rc=$?"; cat $OUT/${TAG}_c3.log
grep -h "t+" $OUT/${TAG}_c3.err | tail -18
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py -m gpu -x -q > $OUT/${TAG}_tests.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/${TAG}_tests.log
