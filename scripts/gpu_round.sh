#!/bin/bash
# One gpurun call that re-establishes the round's GPU evidence for the current build (run from the repo root on the GPU box):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh r02'
# 1. pytest -m gpu (parity through the C ABI), 2. smoke(), 3. the bench line (never under a profiler), 4. the ncu launch list of the
# bench's headline section, 5. one `ncu --set full` capture of wfb_persist_kernel. Outputs land in gpurun_out/<tag>_*; copy the
# summaries to profiles/ afterwards (scripts/ncu_lines.py turns the .ncu-rep into the text summaries kept there).
set -u
TAG=${1:-rNN}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_gpu_tests.log 2>&1; echo "pytest rc=$?"
tail -3 $OUT/${TAG}_gpu_tests.log
timeout 300 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/${TAG}_smoke.log
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-600 $OUT/${TAG}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-map --no-record --no-pipeline > $OUT/${TAG}_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k wfb_persist_kernel -s 1 -c 1 -f -o $OUT/${TAG}_persist \
  python bench.py --steps 1 --warmup 1 --no-cpu --no-map --no-record --no-pipeline > $OUT/${TAG}_persist.log 2>&1; echo "ncu full rc=$?"
ls -la $OUT | tail -8
