#!/bin/bash
set -u
OUT=gpurun_out
TAG=${1:-r02m}
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/${TAG}_gpu_tests.log
timeout 300 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/${TAG}_smoke.log
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("$OUT/${TAG}_bench.json"))
print("value %.1f e2e %.1f Mbp/s; phases"%(d["value"]/1e6,d["e2e"]["value"]/1e6), {k:round(v,1) for k,v in d["phases_ms"].items()}, "cpu %.2f Mbp/s"%(d["cpu_baseline"]["value"]/1e6), d["parity"]["paf_lines_identical"], d["roofline"]["frac"])
PY
