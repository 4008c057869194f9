"""Runs one BASELINE config (tests/configs.py) through the library's two one-call phases — wfb_map_phase then wfb_align_phase, host
sequences in, PAF text out, exactly what `wfmash target.fa [query.fa]` does — and compares the text with what the reference's
UNMODIFIED skch::Map + align::Aligner wrote for the same files (tests/golden/config_reference.json.gz). Shared by the GPU tests,
bench.py and scripts/run_configs_gpu.py. TEST / BENCH INFRASTRUCTURE ONLY."""
import gzip
import hashlib
import json
import os
import time

import numpy as np

from tests import configs, util

_doc = None


def golden():
    global _doc
    if _doc is None:
        with gzip.open(os.path.join(util.GOLD, "config_reference.json.gz"), "rt") as f:
            _doc = json.load(f)
    return _doc


def digests(text: bytes, keep: int):
    return sorted((b"\t".join(ln.split(b"\t")[:keep]).decode(), hashlib.sha256(ln).hexdigest()) for ln in text.split(b"\n") if ln)


def phase_params(wb, cfg):
    prm = dict(cfg["params"])
    f = dict(prm.pop("filter", {}))
    ident = prm.pop("percentage_identity", None)
    w = prm.get("window_length", 1000)
    return wb.MapPhaseParams(filter=wb.FilterParams(window_length=w, **f), percentage_identity=0.0 if ident is None else ident, **prm), w


def run(wb, name, aligner=None, align=True, device=0):
    """-> dict: our texts, timings, and the comparison with the reference's lines."""
    cfg = configs.by_name(name)
    g = golden()[name]
    targets, queries = configs.sequences(cfg)
    MP, w = phase_params(wb, cfg)
    t0 = time.perf_counter()
    mp, mst = wb.map_phase(targets, queries, MP, device)
    t_map = time.perf_counter() - t0
    ours_m, ref_m = digests(mp, 14), sorted((d["head"], d["sha"]) for d in g["mapping"])
    out = dict(name=name, mapping_paf=mp, map_seconds=t_map, map_stats=mst, mapping_rows=len(ours_m), mapping_identical=bool(ours_m == ref_m),
               mapping_cols14_identical=[h for h, _ in ours_m] == [h for h, _ in ref_m],
               identity_identical=bool(np.float32(mst.percentage_identity) == np.float32(g["percentage_identity"])),
               mapping_only_ours=sorted(set(ours_m) - set(ref_m)), mapping_only_ref=sorted(set(ref_m) - set(ours_m)))
    if not align:
        return out
    own = aligner is None
    if own:
        aligner = wb.Aligner(device)
    t0 = time.perf_counter()
    paf, ast = wb.align_phase(aligner, mp, targets, queries if queries is not targets else targets, window_length=w)
    t_al = time.perf_counter() - t0
    if own:
        aligner.close()
    ours_a, ref_a = digests(paf, 12), sorted((d["head"], d["sha"]) for d in g["alignment"])
    out.update(alignment_paf=paf, align_seconds=t_al, align_stats=ast, alignment_lines=len(ours_a), alignment_identical=ours_a == ref_a,
               alignment_only_ours=sorted(set(ours_a) - set(ref_a)), alignment_only_ref=sorted(set(ref_a) - set(ours_a)),
               aligned_bp=int(ast.aligned_bp))
    return out


def summary(r):
    keep = ("name", "map_seconds", "mapping_rows", "mapping_identical", "mapping_cols14_identical", "identity_identical", "align_seconds", "alignment_lines",
            "alignment_identical", "aligned_bp")
    s = {k: r[k] for k in keep if k in r}
    s["mapping_mismatches"] = [len(r["mapping_only_ours"]), len(r["mapping_only_ref"])]
    if "alignment_only_ours" in r:
        s["alignment_mismatches"] = [len(r["alignment_only_ours"]), len(r["alignment_only_ref"])]
        s["align_kernel_ms"] = r["align_stats"].kernel_ms
        s["Mbp_per_s_end_to_end"] = r["aligned_bp"] / (r["map_seconds"] + r["align_seconds"]) / 1e6 if r["aligned_bp"] else 0.0
    st = r["map_stats"]
    s["map"] = dict(fragments=int(st.fragments), l2=int(st.l2_mappings), sketch=int(st.sketch_size), min_hits=int(st.minimum_hits), identity=float(st.percentage_identity),
                    index_s=st.index_seconds, kernels_ms=st.map_kernel_ms, filter_s=st.filter_seconds, stale_absorbed=int(st.stale_absorbed))
    return s
