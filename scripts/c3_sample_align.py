#!/usr/bin/env python3
"""ncu target: the main biWFA alignment (wfb_align_batch) of every `stride`-th mapping record of a BASELINE config, twice
(first call = warm-up). Usage: c3_sample_align.py [config] [stride]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wfmash_b200 as wb
from wfmash_b200 import pipeline
from tests import configrun, configs
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
stride = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cfg = configs.by_name(name)
t, q = configs.sequences(cfg)
MP, w = configrun.phase_params(wb, cfg)
mp, mst = wb.map_phase(t, q, MP)
rows = [ln for ln in mp.split(b"\n") if ln][::stride]
P = pipeline.Params(window_length=w, percentage_identity=float(mst.percentage_identity))
recs = pipeline.records_from_paf(b"".join(x + b"\n" for x in rows), t, q, P)
pairs = [(r["target"], r["query"]) for r in recs]
al = wb.Aligner(0)
for i in range(2):
    t0 = time.time()
    al.align_end2end_batch(pairs)
    st = al.last_stats
    print(f"records {len(pairs)} call {i}: {time.time() - t0:.2f} s kernel_ms {st.kernel_ms:.1f} persist_ms {st.break_kernel_ms:.1f} cells {st.cells} steps {st.score_steps} ovl {st.overlap_tests}", flush=True)
