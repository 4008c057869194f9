#!/usr/bin/env python3
"""Per-phase cycle counters of the score step (tuning only; needs a library built with -DWFB_PHASE_TIMERS, passed via
WFB_LIB). Aligns the bench workload once and prints, per wavefront-width bucket, steps, mean cycles per step, thread 0's
share before the barrier, and mean width."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import wfmash_b200 as wb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 861
recs = bench.make_records(0, n)
pairs = [(p, t) for p, t, _ in recs]
al = wb.Aligner(0)
al.align_end2end_batch(pairs[:32])
buf = (ctypes.c_ulonglong * 32)()
L = wb.lib()
assert L.wfb_debug_phase_timers(buf) == 0
al.align_end2end_batch(pairs)
st = al.last_stats
assert L.wfb_debug_phase_timers(buf) == 0
v = list(buf)
print(f"records {n} kernel_ms {st.kernel_ms:.1f} steps {st.score_steps} cells {st.cells}")
tot = sum(v[4 * b + 1] for b in range(4)) + v[16]
for b, name in enumerate(["<=128", "<=1024", "<=4096", ">4096"]):
    s_, cyc, c0, wd = v[4 * b: 4 * b + 4]
    if s_:
        print(f"width {name:7s} steps {s_:10d} cyc/step {cyc / s_:9.0f} thread0-own {c0 / s_:9.0f} mean width {wd / s_:8.0f} share of CTA cycles {cyc / tot:.3f}")
if v[17]:
    print(f"overlap calls {v[17]} cyc/call {v[16] / v[17]:.0f} share {v[16] / tot:.3f}")
GHZ = 1.965
if v[21]:
    print(f"base tasks {v[21]} mean {v[20] / v[21] / GHZ / 1e3:.1f} us (backtrace {v[22] / max(1, v[23]) / GHZ / 1e3:.1f} us) total {v[20] / GHZ / 1e9:.2f} CTA-s")
if v[25]:
    print(f"break tasks {v[25]} mean {v[24] / v[25] / GHZ / 1e3:.1f} us total {v[24] / GHZ / 1e9:.2f} CTA-s; waiting for tasks {v[26] / GHZ / 1e9:.2f} CTA-s; kernel x 296 CTAs = {st.kernel_ms * 0.296:.2f} CTA-s")
