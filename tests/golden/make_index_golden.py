#!/usr/bin/env python3
"""Regenerates tests/golden/index_reference.json.gz from the UNMODIFIED reference skch::Sketch compiled in place
(oracle/_ref/libsketchref.so = src/map/include/winSketch.hpp + sequenceIds.hpp behind oracle/ref_sketch_driver.cpp): per
case of tests/test_index_ref_cpu.py::CASES the sizes and one SHA-256 over minmerIndex and the hash-sorted
minmerPosLookupIndex. Run in the build container only (needs oracle/_ref, which needs /root/reference)."""
import gzip, json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from tests import util
from tests.test_index_ref_cpu import CASES, case_seqs, names_of, reference_index, digest

R = util.load_ref("libsketchref.so")
assert R is not None, "build oracle/_ref first (make -C oracle)"
cases = []
for seed, k, w, s, F, threads in CASES:
    seqs, ids, groups = case_seqs(seed)
    mi, hs, st, pts, _ = reference_index(R, seqs, names_of(groups), k, w, s, F, threads)
    cases.append({"seed": seed, "kws": [k, w, s], "n_minmers": len(mi), "n_hashes": len(hs), "n_points": len(pts), "sha": digest(mi, hs, st, pts)})
    print(cases[-1])
with gzip.GzipFile(os.path.join(HERE, "index_reference.json.gz"), "wb", mtime=0) as f:
    f.write(json.dumps({"cases": cases}).encode())
