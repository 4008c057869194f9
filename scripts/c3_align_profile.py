#!/usr/bin/env python3
"""Tuning aid: the alignment phase of config C3 (scerevisiae8 all-vs-all) under different batch sizes / root orders, with the
library's per-launch trace (WFB_TRACE) on stderr. Usage: c3_align_profile.py [config] [batch:costsort ...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wfmash_b200 as wb  # noqa: E402
from tests import configrun, configs  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
variants = sys.argv[2:] or ["4096:0", "8192:1", "32768:1"]
cfg = configs.by_name(name)
t, q = configs.sequences(cfg)
MP, w = configrun.phase_params(wb, cfg)
mp, mst = wb.map_phase(t, q, MP)
g = configrun.golden()[name]
ref = sorted((d["head"], d["sha"]) for d in g["alignment"])
al = wb.Aligner(0)
os.environ["WFB_TRACE"] = "1"
for v in variants:
    b, cs = v.split(":")
    os.environ["WFB_COST_SORT"] = cs
    t0 = time.perf_counter()
    paf, ast = wb.align_phase(al, mp, t, q, window_length=w, batch_records=int(b))
    dt = time.perf_counter() - t0
    ok = configrun.digests(paf, 12) == ref
    print(f"RESULT {name} batch={b} cost_sort={cs} align_s={dt:.2f} kernel_ms={ast.kernel_ms:.0f} Mbp/s={ast.aligned_bp / dt / 1e6:.1f} identical={ok}", flush=True)
    sys.stderr.write(f"---- end of variant {v}\n")
al.close()
