#!/bin/bash
set -u
OUT=gpurun_out
TAG=${1:-r02f}
mkdir -p $OUT
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("$OUT/${TAG}_bench.json"))
print("value %.1f e2e %.1f Mbp/s; phases"%(d["value"]/1e6,d["e2e"]["value"]/1e6), {k:round(v,1) for k,v in d["phases_ms"].items()}, "cpu %.2f Mbp/s"%(d["cpu_baseline"]["value"]/1e6), d["parity"]["paf_lines_identical"])
PY
for v in ns0 ns1 ns2 ns8; do
  WFB_LIB=scripts/_build/libwfb_$v.so timeout 300 python scripts/c3_align_profile.py C3 32768:1 32768:1 > $OUT/${TAG}_$v.log 2> $OUT/${TAG}_$v.err; echo "$v rc=$? $(tail -1 $OUT/${TAG}_$v.log) $(grep 'main n=' $OUT/${TAG}_$v.err | tail -1 | cut -c1-70)"
done
