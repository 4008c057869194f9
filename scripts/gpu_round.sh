#!/bin/bash
# One gpurun call that re-establishes the round's GPU evidence for the current build (run from the repo root on the GPU box):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash scripts/gpu_round.sh r03'
# 1. pytest -m gpu (parity through the C ABI), 2. smoke(), 3. the bench line (never under a profiler), 4. the ncu launch list of the bench
# command, 5. ncu metrics (duration, DRAM bytes, instruction counts, stall reasons) of wfb_persist_kernel on the full C3 launch — single-pass
# metrics only: `--set full` replays the kernel ~40 times and each replay restores the 30+ GB workspace (10 minutes for a 1 s kernel; use
# scripts/c3_sample_align.py with a stride for a source-level capture). Outputs land in gpurun_out/<tag>_*; copy summaries to profiles/.
set -u
TAG=${1:-rNN}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/${TAG}_gpu_tests.log
timeout 300 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/${TAG}_smoke.log
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-400 $OUT/${TAG}_bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu > $OUT/${TAG}_launches.log 2>&1; echo "ncu launches rc=$?"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,smsp__inst_executed_op_global_ld.sum,smsp__inst_executed_op_global_st.sum,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,smsp__inst_executed_op_shared_ld.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
timeout 900 ncu --metrics $M --clock-control none -k regex:wfb_persist_kernel -s 1 -c 1 --csv --log-file $OUT/${TAG}_persist_c3_metrics.csv \
  python scripts/c3_sample_align.py C3 1 > $OUT/${TAG}_persist_metrics.log 2>&1; echo "ncu metrics rc=$?"
ls -la $OUT | tail -8
