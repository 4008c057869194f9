#!/usr/bin/env python3
"""Regenerates tests/golden/ani_reference.json.gz: the identity the UNMODIFIED reference Stat::estimate_identity_for_groups
(oracle/_ref/libstatsref.so = src/map/include/map_stats.hpp compiled in place behind oracle/ref_stats_driver.cpp) returns for
every set of tests/test_ani_cpu.py::SETS over tests.aniutil.case() (as double bit patterns), and the SHA-256 of the group
sketches of the oracle's literal StreamingMinHash restatement. Run in the build container only (needs oracle/_ref)."""
import gzip, hashlib, json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from tests import util, aniutil
from tests.test_ani_cpu import SETS
from wfmash_b200 import pipeline

R = util.load_ref("libstatsref.so")
assert R is not None, "build oracle/_ref first (make -C oracle)"
seqs = aniutil.case()
ids = pipeline.SequenceIds(seqs, seqs, "#")
gids = sorted(set(ids.group)); dense = {g: i for i, g in enumerate(gids)}
grp = [dense[ids.group[ids.id_of[n]]] for n, _ in seqs]
sk, cnt = aniutil.oracle_group_sketches(util.load_oracle(), [x for _, x in seqs], grp, len(gids), 21, 4096)
ident = {name: float(aniutil.reference_identity(R, seqs[t0:t1], seqs[q0:q1], "#", pct, adj)).hex() for name, (q0, q1), (t0, t1), pct, adj in SETS}
print(ident)
with gzip.GzipFile(os.path.join(HERE, "ani_reference.json.gz"), "wb", mtime=0) as f:
    f.write(json.dumps({"identity": ident, "sketch_sha": hashlib.sha256(sk.tobytes() + cnt.tobytes()).hexdigest()}).encode())
