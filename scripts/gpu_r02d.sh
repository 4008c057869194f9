#!/bin/bash
set -u
OUT=gpurun_out
TAG=${1:-r02d}
mkdir -p $OUT
timeout 300 python scripts/c3_align_profile.py C3 32768:1 32768:1 > $OUT/${TAG}_c3.log 2> $OUT/${TAG}_c3.err; echo "c3 rc=$?"; cat $OUT/${TAG}_c3.log
grep -h "persist:\|main n=\|endsfree\|t+" $OUT/${TAG}_c3.err | tail -24
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "high_divergence or work_counts or golden_vectors or medium" > $OUT/${TAG}_tests.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/${TAG}_tests.log
