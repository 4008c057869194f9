#!/bin/bash
set -u
OUT=gpurun_out
TAG=${1:-r02e}
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/${TAG}_gpu_tests.log
WFB_TRACE=1 timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-900 $OUT/${TAG}_bench.json
grep -h "endsfree\|head + tail" $OUT/${TAG}_bench.err | tail -8
