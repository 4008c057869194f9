#!/usr/bin/env python3
"""tests/golden/tiny_target_reference.json.gz: what the reference's UNMODIFIED skch::Map (oracle/_ref/libmapperref.so, one thread) writes for a
sequence set with targets of exactly one window (1000 bases) and barely more. All minmers of such a target tie on (wpos, wpos_end), their order
is what the reference's unstable std::sort leaves, and computeL2MappedRegions (sketch evaluated after every insertion) reports other
conservedSketches / identities for another order (c#1#chr1 -> d#1#chrZ: 5 shared minmers, 4 with the ties in hash order). The set was found
by the round-2 map-phase fuzz; the file carries the sequences themselves. Re-creates the expected rows from the stored sequences:
    python tests/golden/make_tiny_target_golden.py          (needs /root/reference)"""
import gzip
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import pipeutil, util  # noqa: E402

PATH = os.path.join(ROOT, "tests", "golden", "tiny_target_reference.json.gz")

if __name__ == "__main__":
    M = util.load_ref("libmapperref.so")
    assert M is not None, "make -C oracle ref"
    with gzip.open(PATH, "rt") as f:
        doc = json.load(f)
    seqs = [(n, s.encode()) for n, s in doc["sequences"]]
    for name, c in doc["cases"].items():
        txt = pipeutil.reference_map_phase(M, seqs, pipeutil.params(c["params"]))
        c["rows"] = sorted(ln.decode() for ln in txt.split(b"\n") if ln)
        print(name, len(c["rows"]))
    with gzip.open(PATH, "wt") as f:
        json.dump(doc, f)
