#!/usr/bin/env python3
"""Tuning sweep of the breakpoint kernel's CTA shape / register budget (run on the GPU box).
Each config = (library variant, WFB_BREAK_THREADS); prints one line per config."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# usage: sweep_break.py [records] [variant:threads ...]   (variants = wfmash_b200/variants/lib_<variant>.so)
configs = [("q_256_2", 256), ("q_256_2", 128), ("q_256_3", 256), ("q_128_4", 128), ("q_128_6", 128), ("q_256_2", 192)]
recs = sys.argv[1] if len(sys.argv) > 1 else "861"
if len(sys.argv) > 2:
    configs = [tuple(a.split(":")) for a in sys.argv[2:]]  # variant:threads[:records]
for cfg in configs:
    var, thr = cfg[0], int(cfg[1])
    nrec = cfg[2] if len(cfg) > 2 else recs
    env = dict(os.environ, WFB_LIB=os.path.join(ROOT, "wfmash_b200", "variants", f"lib_{var}.so"), WFB_BREAK_THREADS=str(thr))
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "1", "--no-cpu", "--no-map", "--no-record", "--records", nrec],
                       env=env, capture_output=True, text=True, timeout=900)
    try:
        d = json.loads(p.stdout.strip().splitlines()[-1])
        st = d["config"]["stats"]
        print(f"{var:14s} recs={nrec:>5s} thr={thr:5d} value={d['value']/1e6:8.3f} Mbp/s e2e={d['e2e']['value']/1e6:8.3f} ms/step={d['ms_per_step']:9.1f} "
              f"break_ms={st['break_kernel_ms']:9.1f} cells={st['cells']/1e9:.2f}G ovl={st['overlap_tests']/1e9:.2f}G "
              f"frac={d['roofline']['frac']:.3f} clocks={d['clocks']['sm_mhz']}", flush=True)
    except Exception as e:
        print(var, thr, "FAILED", e, p.stderr[-500:], flush=True)
