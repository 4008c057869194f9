#!/usr/bin/env python3
"""Tuning aid: per-phase cycle counters of the score step on the records of a BASELINE config (default C3). Needs a library built
with -DWFB_PHASE_TIMERS, passed via WFB_LIB (this script builds it when WFB_LIB is unset). Usage: c3_phase_timers.py [config] [stride]"""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if "WFB_LIB" not in os.environ:
    from wfmash_b200 import build
    out = os.path.join(ROOT, "gpurun_out", "libwfb_timers.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    build.build_variant(out, build.DEFAULT_DEFS + ["-DWFB_PHASE_TIMERS"])
    os.environ["WFB_LIB"] = out
    os.execv(sys.executable, [sys.executable] + sys.argv)
import wfmash_b200 as wb
from wfmash_b200 import pipeline
from tests import configrun, configs

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
stride = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg = configs.by_name(name)
t, q = configs.sequences(cfg)
MP, w = configrun.phase_params(wb, cfg)
mp, mst = wb.map_phase(t, q, MP)
P = pipeline.Params(window_length=w, percentage_identity=float(mst.percentage_identity))
t0 = time.time()
recs = pipeline.records_from_paf(mp, t, q, P)[::stride]
pairs = [(r["target"], r["query"]) for r in recs]
print(f"{len(pairs)} records built in {time.time() - t0:.1f} s", flush=True)
al = wb.Aligner(0)
al.align_end2end_batch(pairs[:32])
buf = (ctypes.c_ulonglong * 32)()
L = wb.lib()
assert L.wfb_debug_phase_timers(buf) == 0
al.align_end2end_batch(pairs)
st = al.last_stats
assert L.wfb_debug_phase_timers(buf) == 0
v = list(buf)
print(f"records {len(pairs)} kernel_ms {st.kernel_ms:.1f} steps {st.score_steps} cells {st.cells}")
tot = sum(v[4 * b + 1] for b in range(4)) + v[16]
for b, nm in enumerate(["<=128", "<=1024", "<=4096", ">4096"]):
    s_, cyc, c0, wd = v[4 * b: 4 * b + 4]
    if s_:
        print(f"width {nm:7s} steps {s_:10d} cyc/step {cyc / s_:9.0f} thread0-own {c0 / s_:9.0f} mean width {wd / s_:8.0f} share of CTA cycles {cyc / tot:.3f}")
if v[17]:
    print(f"overlap calls {v[17]} cyc/call {v[16] / v[17]:.0f} share {v[16] / tot:.3f}")
GHZ = 1.965
if v[21]:
    print(f"base tasks {v[21]} mean {v[20] / v[21] / GHZ / 1e3:.1f} us (backtrace {v[22] / max(1, v[23]) / GHZ / 1e3:.1f} us) total {v[20] / GHZ / 1e9:.2f} CTA-s")
if v[25]:
    print(f"break tasks {v[25]} mean {v[24] / v[25] / GHZ / 1e3:.1f} us total {v[24] / GHZ / 1e9:.2f} CTA-s; waiting for tasks {v[26] / GHZ / 1e9:.2f} CTA-s; kernel x 296 CTAs = {st.kernel_ms * 0.296:.2f} CTA-s")
