#!/bin/bash
# One short gpurun call for the round's last changes (candidate-filtered minmer build; scaled synthetic configs C4s / C5s):
#   /usr/local/graft/bin/gpurun --timeout 300 -- 'bash scripts/gpu_r02mm.sh r02mm [quick]'
set -u
TAG=${1:-r02mm}
MODE=${2:-all}
OUT=gpurun_out
mkdir -p $OUT
timeout 100 python scripts/gpu_mm_shot.py $TAG > $OUT/${TAG}_shot.log 2>&1; echo "mm shot rc=$?"; cut -c1-100,330-700 $OUT/${TAG}_shot.log | tail -14
if [ "$MODE" = quick ]; then
  timeout 100 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zsynthetic.py -q -x -m gpu --durations=8 \
    -k "minmers_match or index_build or l1_hit or l2_mappings or minmer_build_modes" > $OUT/${TAG}_parity.log 2>&1; echo "parity subset rc=$?"; tail -14 $OUT/${TAG}_parity.log | cut -c1-300
  exit 0
fi
timeout 170 python -m pytest tests/test_gpu_zsynthetic.py -q -x -s -m gpu --durations=0 > $OUT/${TAG}_zsynth.log 2>&1; echo "zsynthetic rc=$?"; tail -12 $OUT/${TAG}_zsynth.log | cut -c1-400
timeout 80 python bench.py --steps 1 --warmup 1 --no-cpu > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-300 $OUT/${TAG}_bench.json
