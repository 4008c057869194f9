#!/usr/bin/env python3
"""Regenerates tests/golden/l2_reference.json.gz from the UNMODIFIED reference L2 code compiled in place
(oracle/_ref/libl2ref.so = src/map/include/slidingMap.hpp + mappingCore.hpp:306-442 behind oracle/ref_l2_driver.cpp):
every L2_mapLocus_t (fragment, locus, seqId, sharedSketchSize, meanOptimalPos, optimalStart, optimalEnd, strand) of
every L1 locus of tests.maputil.l2_case() for the parameter sets of tests/test_l2_emu_cpu.py::L2_CASES. The index and
the L1 loci the L2 stage reads are the oracle's (pinned separately by map_reference.json.gz); L2 itself is the reference's.
Run in the build container only (needs oracle/_ref, which needs /root/reference)."""
import gzip, json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from tests import util, maputil
from tests.test_l2_emu_cpu import L2_CASES

oracle = util.load_oracle()
ref = util.load_ref("libl2ref.so")
assert ref is not None, "build oracle/_ref first (make -C oracle)"
cases = []
for seed, mode, (k, w, s) in L2_CASES:
    seqs, ids, groups = maputil.l2_case(seed=seed)
    index = maputil.oracle_index(oracle, seqs, ids, k, w, s, 0.0002, 3)
    rows = maputil.l2_all_loci(ref, "ref", index, seqs, ids, groups, k, w, s, oracle, mode)
    cases.append({"seed": seed, "mode": list(mode), "kws": [k, w, s], "rows": rows.tolist()})
    print(seed, mode, (k, w, s), "L2 loci", len(rows))
with gzip.GzipFile(os.path.join(HERE, "l2_reference.json.gz"), "wb", mtime=0) as f:
    f.write(json.dumps({"cases": cases}).encode())
