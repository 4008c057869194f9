#!/usr/bin/env python3
"""Generates tests/golden/config_reference.json.gz: what the reference's UNMODIFIED two phases (skch::Map -> libmapperref.so, then
align::Aligner -> libalignref.so, compiled in place from /root/reference by oracle/Makefile) write for BASELINE.json's configs
on the real input files (tests/data/ = byte copies of /root/reference/data/):

  C1      data/reference.fa.gz vs data/reads.255bps.fa.gz, defaults  -> empty (reads shorter than -w 1k, computeMap.hpp:560-602)
  C1w250  the same with -w 250                                        -> non-empty
  C2      data/LPA.subset.fa.gz self, -k15 -w1k -P50k (ANI auto-identity, the CLI default)
  C3sub   data/scerevisiae8.fa.gz -Y '#': two genomes x three chromosomes (the subset the GPU parity test runs)
  C3      data/scerevisiae8.fa.gz -Y '#', all 8 genomes (only with --full: minutes of CPU)

For every config it stores the adopted identity, the mapping-PAF lines (columns 1-14, i.e. without the schedule-dependent ch:Z: tag,
plus a digest of the whole text of the -t 1 run) and, per alignment record, the first 12 PAF columns + a sha256 of the whole line
(cg:Z: included). Needs /root/reference (it cannot run on the GPU box); the GPU tests read the committed file.

    python tests/golden/make_config_golden.py [--full] [--only NAME]"""
import argparse
import contextlib
import gzip
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tests import aniutil, datasets, pipeutil, util  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "config_reference.json.gz")


@contextlib.contextmanager
def quiet():
    """The reference logs to stderr from C++; keep the terminal readable."""
    sys.stderr.flush()
    saved = os.dup(2)
    null = os.open(os.devnull, os.O_WRONLY)
    os.dup2(null, 2)
    try:
        yield
    finally:
        os.dup2(saved, 2)
        os.close(null)
        os.close(saved)


def configs(full):
    from tests import configs as C
    return [c for c in C.CONFIGS if full or not c["full_only"]]


def run_reference(cfg, threads):
    """-> dict(identity, mapping PAF text, alignment PAF text, seconds) from the unmodified reference."""
    from tests import configs as C
    M, A, S = util.load_ref("libmapperref.so"), util.load_ref("libalignref.so"), util.load_ref("libstatsref.so")
    assert M is not None and A is not None and S is not None, "oracle/_ref not built: make -C oracle ref"
    targets, queries = C.sequences(cfg)
    prm = dict(cfg["params"])
    t0 = time.time()
    ident = None
    if prm.get("percentage_identity") is None:
        with quiet():
            ident = aniutil.reference_identity(S, targets, queries, threads=threads)
        prm["percentage_identity"] = C.adopt_identity(ident)
    t_ani = time.time() - t0
    P = pipeutil.params(prm)
    same = queries is targets
    t0 = time.time()
    with quiet():
        mp = pipeutil.reference_map_phase(M, targets, P, threads=1, queries=None if same else queries)
    t_map = time.time() - t0
    A.ref_align_set_threads(threads)
    t0 = time.time()
    with quiet():
        al = pipeutil.reference_align_phase(A, mp, targets if same else targets + queries, P) if mp else b""
    t_align = time.time() - t0
    A.ref_align_set_threads(1)
    return dict(identity=ident, percentage_identity=prm["percentage_identity"], mapping_paf=mp, alignment_paf=al, seconds=dict(ani=t_ani, map=t_map, align=t_align))


def digest_lines(text, keep):
    lines = sorted(ln for ln in text.split(b"\n") if ln)
    return [{"head": b"\t".join(ln.split(b"\t")[:keep]).decode(), "sha": hashlib.sha256(ln).hexdigest()} for ln in lines]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--only", default=None)
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    ap.add_argument("--cache", default="/tmp/wfb_config_cache", help="scratch directory for the raw reference texts ('' = none)")
    a = ap.parse_args()
    doc = {}
    if os.path.exists(OUT):
        with gzip.open(OUT, "rt") as f:
            doc = json.load(f)
    for cfg in configs(a.full):
        if a.only and cfg["name"] != a.only:
            continue
        cache = os.path.join(a.cache, cfg["name"] + ".json")
        if a.cache and os.path.exists(cache + ".map.paf"):   # raw texts of an earlier run (scratch, never committed)
            r = json.load(open(cache))
            r["mapping_paf"], r["alignment_paf"] = open(cache + ".map.paf", "rb").read(), open(cache + ".aln.paf", "rb").read()
        else:
            r = run_reference(cfg, a.threads)
            if a.cache:
                os.makedirs(a.cache, exist_ok=True)
                open(cache + ".map.paf", "wb").write(r["mapping_paf"]); open(cache + ".aln.paf", "wb").write(r["alignment_paf"])
                json.dump({k: v for k, v in r.items() if not k.endswith("_paf")}, open(cache, "w"))
        mlines = [ln for ln in r["mapping_paf"].split(b"\n") if ln]
        alines = [ln for ln in r["alignment_paf"].split(b"\n") if ln]
        mapped_bp = sum(int(f[3]) - int(f[2]) for f in (ln.split(b"\t") for ln in mlines))
        # the reference's own "total aligned bp" counter (computeAlignments.hpp:481,528) adds qEndPos - qStartPos of the PARSED row, i.e.
        # after parseMashmapRow's query padding; the library's row parser equals the unmodified parseMashmapRow field by field
        # (tests/test_filter_cpu.py::test_mapping_paf_parse_matches_compiled_reference_live)
        import wfmash_b200 as wb
        w = cfg["params"].get("window_length", 1000)
        aligned_bp = 0
        for ln in mlines:
            row, _, _ = wb.mapping_paf_parse(ln, min(w, 5000), min(w, 5000), w * 128)
            aligned_bp += row.q_end - row.q_start
        entry = dict(identity=r["identity"], percentage_identity=r["percentage_identity"], mapping_rows=len(mlines), alignment_lines=len(alines),
                     mapped_query_bp=mapped_bp, aligned_bp=aligned_bp, mapping_sha_t1=hashlib.sha256(r["mapping_paf"]).hexdigest(),
                     mapping_sha_sorted=hashlib.sha256(b"\n".join(sorted(mlines))).hexdigest(),
                     mapping_sha_cols14=hashlib.sha256(b"\n".join(sorted(b"\t".join(ln.split(b"\t")[:14]) for ln in mlines))).hexdigest(),
                     alignment_sha_sorted=hashlib.sha256(b"\n".join(sorted(alines))).hexdigest(), reference_seconds=r["seconds"])
        entry["mapping"] = digest_lines(r["mapping_paf"], 14)
        entry["alignment"] = digest_lines(r["alignment_paf"], 12)
        doc[cfg["name"]] = entry
        print(cfg["name"], {k: v for k, v in entry.items() if k not in ("mapping", "alignment")}, flush=True)
    with gzip.open(OUT, "wt") as f:
        json.dump(doc, f)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
