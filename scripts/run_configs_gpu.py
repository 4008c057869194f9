#!/usr/bin/env python3
"""Runs BASELINE configs through wfb_map_phase + wfb_align_phase on the GPU, prints a summary per config and dumps the lines that
differ from the reference's (tests/golden/config_reference.json.gz) under gpurun_out/. Usage: run_configs_gpu.py [names...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import wfmash_b200 as wb  # noqa: E402
from tests import configrun  # noqa: E402

names = sys.argv[1:] or ["C2", "C2p80n5", "C3sub", "C3"]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
al = wb.Aligner(0)
for name in names:
    for rep in range(2 if name == "C3" else 1):
        r = configrun.run(wb, name, aligner=al)
        print(json.dumps(configrun.summary(r)), flush=True)
    bad = {k: r[k][:200] for k in ("mapping_only_ours", "mapping_only_ref", "alignment_only_ours", "alignment_only_ref") if r.get(k)}
    if bad:
        json.dump(bad, open(os.path.join(ROOT, "gpurun_out", f"config_diff_{name}.json"), "w"), indent=1)
        want = {h for h, _ in r.get("alignment_only_ref", [])}
        with open(os.path.join(ROOT, "gpurun_out", f"config_diff_{name}.paf"), "wb") as f:
            for ln in r.get("alignment_paf", b"").split(b"\n"):
                if ln and b"\t".join(ln.split(b"\t")[:12]).decode() in want:
                    f.write(ln[:20000] + b"\n")
al.close()
