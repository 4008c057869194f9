#!/bin/bash
# Last short gpurun call of round 2: variants of the filtered minmer stream (departure cursor, shared-memory containers, chunk sizes) on all of
# scerevisiae8 — every variant must return the bytes of the unfiltered build —, the N-run / repeat parity cases under the two candidate
# defaults, and ncu metrics of the minmer kernels of those two.
set -u
TAG=${1:-r02mm3}
OUT=gpurun_out
mkdir -p $OUT
timeout 30 python scripts/gpu_mm_shot.py $TAG > $OUT/${TAG}_shot.log 2>&1; echo "mm shot rc=$?"
python - <<PY
import json
d = json.load(open("$OUT/${TAG}_mm.json"))
for r in d["runs"]:
    print(r["mode"], r.get("same_bytes_as_first"), "stream %.1f cand %.2f filt %.1f redo %.1f redo_chunks %d" % (r.get("stream_kernel_ms", -1), r.get("cand_kernel_ms", -1), r.get("filtered_stream_ms", -1), r.get("redo_ms", -1), r.get("redo_chunks", -1)) if "error" not in r else r["error"])
print(json.dumps(d.get("c3_map_phase"))[:300])
PY
for V in "WFB_MM_LCUR=1" "WFB_MM_LCUR=1 WFB_MM_FSMEM=1"; do
  env $V WFB_MM_FCHUNK=1024 timeout 25 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zsynthetic.py -q -x -m gpu -k "minmers_match or index_build or minmer_build_modes" > $OUT/${TAG}_parity_$(echo $V | tr ' =' '__').log 2>&1; echo "parity [$V] rc=$?"
done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
timeout 40 ncu --metrics $M --clock-control none -k regex:mm_ --csv --log-file $OUT/${TAG}_ncu_mm.csv python scripts/gpu_mm_shot.py ${TAG}ncu ncu > $OUT/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
