#!/usr/bin/env python3
"""One short GPU run for the candidate-filtered minmer build: (1) wfb_minmers_build on all of scerevisiae8 (96 Mbp, s = 24) in every build
mode / chunk size — kernel times and that every mode returns the same bytes; (2) the C3 mapping phase (default build mode) against the
reference fixture. Results go to gpurun_out/<tag>_mm.json as they are produced."""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wfmash_b200 as wb  # noqa: E402
from tests import configrun, datasets  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
out_path = os.path.join(ROOT, "gpurun_out", f"{tag}_mm.json")
os.makedirs(os.path.dirname(out_path), exist_ok=True)
doc = {"runs": []}


def flush():
    with open(out_path, "w") as f:
        json.dump(doc, f, indent=1)


seqs = [x for _, x in datasets.load("yeast")]
ids = list(range(len(seqs)))
base = None
F = {"WFB_MM_FILTER": "1"}
modes = [("unfiltered", {"WFB_MM_FILTER": "0"}),
         ("filtered-1024", dict(F, WFB_MM_FCHUNK="1024")), ("filtered-1024", dict(F, WFB_MM_FCHUNK="1024")),
         ("filtered-1024-lcur", dict(F, WFB_MM_FCHUNK="1024", WFB_MM_LCUR="1")),
         ("filtered-1024-lcur-smem", dict(F, WFB_MM_FCHUNK="1024", WFB_MM_LCUR="1", WFB_MM_FSMEM="1")),
         ("filtered-512-lcur-smem", dict(F, WFB_MM_FCHUNK="512", WFB_MM_LCUR="1", WFB_MM_FSMEM="1")),
         ("filtered-2048-lcur-smem", dict(F, WFB_MM_FCHUNK="2048", WFB_MM_LCUR="1", WFB_MM_FSMEM="1")),
         ("filtered-4096-lcur-smem", dict(F, WFB_MM_FCHUNK="4096", WFB_MM_LCUR="1", WFB_MM_FSMEM="1")),
         ("filtered-1536-lcur", dict(F, WFB_MM_FCHUNK="1536", WFB_MM_LCUR="1")),
         ("filtered-2048-lcur", dict(F, WFB_MM_FCHUNK="2048", WFB_MM_LCUR="1"))]
if len(sys.argv) > 2 and sys.argv[2] == "ncu":  # the two candidate defaults only, once each, no mapping phase (run under ncu)
    modes = [modes[3], modes[4]]
for mode, env in modes:
    os.environ.update(env)
    try:
        t0 = time.perf_counter()
        mm, st = wb.minmers_build(seqs, ids, 15, 1000, 24)
        dt = time.perf_counter() - t0
        sha = hashlib.sha256(mm.tobytes()).hexdigest()
        if base is None:
            base = sha
        r = dict(mode=mode, minmers=int(len(mm)), same_bytes_as_first=sha == base, wall_s=round(dt, 3), **st.as_dict())
    except Exception as e:  # noqa: BLE001
        r = dict(mode=mode, error=repr(e))
    for k in env:
        os.environ.pop(k)
    doc["runs"].append(r)
    print(json.dumps(r), flush=True)
    flush()

for mode, env in ((("default", {}),) if not (len(sys.argv) > 2 and sys.argv[2] == "ncu") else ()):
    os.environ.update(env)
    try:
        for rep in range(2):
            r = configrun.run(wb, "C3", align=False)
        s = configrun.summary(r)
        s["mode"] = mode
        s["index_kernel_ms"] = r["map_stats"].index_kernel_ms
    except Exception as e:  # noqa: BLE001
        s = dict(mode=mode, error=repr(e))
    for k in env:
        os.environ.pop(k)
    doc.setdefault("c3_map_phase", []).append(s)
    print(json.dumps(s), flush=True)
    flush()
