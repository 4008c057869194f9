#!/usr/bin/env python3
"""Regenerates tests/golden/map_reference.json.gz from the UNMODIFIED reference mapping code compiled in place:
  * CommonFunc::addMinmers (src/map/include/commonFunc.hpp:439-708) via oracle/_ref/libmapref.so: per case of
    tests.maputil.minmer_cases() the number of MinmerInfo records, a SHA-256 over (hash, wpos, wpos_end, seqId, strand)
    in output order (ties of the
    reference's unstable final sort put in hash order, see tests.maputil.canonical) and the first 8 records in clear;
  * MappingCore::getSeedIntervalPoints + computeL1CandidateRegions (src/map/include/mappingCore.hpp:81-301, driven as
    Map::doL1Mapping does, computeMap.hpp:945-983) via oracle/_ref/libl1ref.so: every L1 locus
    (mode, fragment, seqId, rangeStartPos, rangeEndPos, intersectionSize) of tests.maputil.l1_case().
The reference has no byte-exact tests for src/map; these fixtures are outputs of the reference itself.
Run in the build container only (needs oracle/_ref, which needs /root/reference)."""
import gzip, json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from tests import util, maputil

oracle = util.load_oracle()
mref = util.load_ref("libmapref.so")
lref = util.load_ref("libl1ref.so")
assert mref is not None and lref is not None, "build oracle/_ref first (make -C oracle)"
F = ("hash", "wpos", "wpos_end", "seqId", "strand")
mm = []
for name, sq, k, w, s in maputil.minmer_cases():
    r = maputil.canonical(maputil.ref_add_minmers(mref, sq, k, w, s, 7))  # tie order of the unstable sort removed
    mm.append({"name": name, "k": k, "w": w, "s": s, "n": int(len(r)), "sha256": maputil.digest(r, F),
               "head": [[int(x[f]) for f in F] for x in r[:8]]})
seqs, ids, groups = maputil.l1_case()
k, w, s = 15, 1000, 29
# the index the L1 stage reads is the oracle's (Sketch::build needs htslib: unbuildable here); L1 itself is the reference's
index = maputil.oracle_index(oracle, seqs, ids, k, w, s, 0.0002, 3)
rows = maputil.l1_all_fragments(lref, "ref_l1_fragment", index, seqs, ids, groups, k, w, s, oracle)
with gzip.GzipFile(os.path.join(HERE, "map_reference.json.gz"), "wb", mtime=0) as f:
    f.write(json.dumps({"minmers": mm, "l1": {"k": k, "w": w, "s": s, "rows": rows.tolist()}}).encode())
print("minmer cases", len(mm), [m["n"] for m in mm], "L1 loci", len(rows))
