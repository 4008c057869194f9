#!/usr/bin/env python3
"""Continuity with round 1: the same synthetic record sets through wfb_align_batch with the current library.
  * the round-1 bench workload (861 records, 5-25 kb, 1/2/5/10 %, 1 kb flanks): was 15.2 Mbp/s device-resident
  * a 200-record batch of the same shape: was 3.76 Mbp/s
  * one 50 kb record at 5 % alone: was 0.75 s
aligned bp = sum of the query lengths (the records' text side)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wfmash_b200 as wb  # noqa: E402
from wfmash_b200 import synth  # noqa: E402

al = wb.Aligner(0)
def run(name, recs, reps=3):
    pairs = [(p, t) for p, t, _ in recs]
    bp = sum(len(t) for _, t in pairs)
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        res = al.align_end2end_batch(pairs)
        dt = time.perf_counter() - t0
        st = al.last_stats
        assert all(r.status == 0 for r in res)
        if best is None or st.kernel_ms < best[1]:
            best = (dt, st.kernel_ms, st.break_kernel_ms)
    print(f"{name}: {len(pairs)} records, {bp / 1e6:.2f} Mbp: wall {best[0] * 1e3:.0f} ms, kernels {best[1]:.1f} ms (persist {best[2]:.1f}) -> "
          f"{bp / best[1] / 1e3:.2f} Mbp/s device, {bp / best[0] / 1e6:.2f} Mbp/s through the host-buffer call", flush=True)

run("r01 bench workload", synth.mapping_records(861, seed=1234, len_lo=5000, len_hi=25000, divergences=[0.01, 0.02, 0.05, 0.10], pad=1000))
run("200-record batch", synth.mapping_records(200, seed=1234, len_lo=5000, len_hi=25000, divergences=[0.01, 0.02, 0.05, 0.10], pad=1000))
run("one 50 kb record at 5 %", synth.mapping_records(1, seed=77, len_lo=50000, len_hi=50000, divergences=[0.05], pad=1000))
run("one 50 kb record at 20 %", synth.mapping_records(1, seed=78, len_lo=50000, len_hi=50000, divergences=[0.20], pad=0))
