#!/usr/bin/env python3
"""Tuning aid: the mapping phase of a BASELINE config (default C3) three times with the library's host-stage timeline (WFB_TRACE) on stderr."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["WFB_TRACE"] = "1"
import wfmash_b200 as wb  # noqa: E402
from tests import configrun, configs  # noqa: E402
name = sys.argv[1] if len(sys.argv) > 1 else "C3"
cfg = configs.by_name(name)
t, q = configs.sequences(cfg)
MP, w = configrun.phase_params(wb, cfg)
for i in range(3):
    t0 = time.perf_counter()
    mp, mst = wb.map_phase(t, q, MP)
    dt = time.perf_counter() - t0
    print(f"map_phase call {i}: {dt * 1e3:.0f} ms  index {mst.index_seconds * 1e3:.0f} ms (kernels {mst.index_kernel_ms:.0f}) ani {mst.ani_seconds * 1e3:.0f} ms (kernels {mst.ani_kernel_ms:.0f}) "
          f"map kernels {mst.map_kernel_ms:.0f} ms filter {mst.filter_seconds * 1e3:.0f} ms rows {mp.count(10)}", flush=True)
    sys.stderr.write(f"---- end of call {i}\n")
