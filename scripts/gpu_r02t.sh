#!/bin/bash
set -u
OUT=gpurun_out
TAG=${1:-r02t}
mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu > $OUT/${TAG}_launches.log 2>&1; echo "ncu launches rc=$?"
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("$OUT/${TAG}_launches.csv")) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if r[0]=="ID"]
h=rows[hdr[0]]; kn=h.index("Kernel Name"); mv=h.index("Metric Value")
t=collections.OrderedDict()
for r in rows[hdr[0]+1:]:
    try: t[r[kn].split("(")[0]] = t.get(r[kn].split("(")[0],0.0)+float(r[mv].replace(",",""))
    except: pass
tot=sum(t.values())
print("launches", len(rows)-hdr[0]-1, "total ms %.1f"%(tot/1e6))
for k,v in sorted(t.items(), key=lambda kv:-kv[1])[:12]: print("%6.2f%%  %10.2f ms  %s"%(100*v/tot, v/1e6, k[:90]))
PY
WFB_LIB=scripts/_build/libwfb_h96.so timeout 300 python scripts/r01_workloads.py 2>&1 | tee $OUT/${TAG}_h96_r01.log | cut -c1-200
