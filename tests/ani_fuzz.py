"""Differential fuzz of the ANI auto-identity path (wfb_ani_group_sketches under the host emulation build, WFB_LIB + wfb_ani_estimate_identity)
against the reference's UNMODIFIED Stat::estimate_identity_for_groups (oracle/_ref/libstatsref.so): random sequence sets, query / target slices,
percentiles and adjustments; the returned doubles must have the same bit pattern. TEST INFRASTRUCTURE. python tests/ani_fuzz.py SEED SECONDS [MAX]"""
import json
import os
import random
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wfmash_b200 as wb
from wfmash_b200 import pipeline
from tests import aniutil, util
R = util.load_ref("libstatsref.so")
if R is None:
    print(json.dumps({"cases": 0, "mismatches": 0, "skipped": "oracle/_ref/libstatsref.so not built"}))
    sys.exit(0)
MAX_N = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 30
oracle = util.load_oracle()
rnd = random.Random(int(sys.argv[1])); T_END = time.time() + float(sys.argv[2])
n = bad = 0
null = os.open(os.devnull, os.O_WRONLY); saved = os.dup(2)
while time.time() < T_END and bad < 3 and n < MAX_N:
    seqs = aniutil.case(seed=rnd.randrange(1, 10**6), length=rnd.choice([5_000, 12_000, 25_000]), small=rnd.random() < 0.5)
    m = len(seqs)
    q0 = rnd.randrange(0, m - 1); q1 = rnd.randrange(q0 + 1, m + 1); t0 = rnd.randrange(0, m - 1); t1 = rnd.randrange(t0 + 1, m + 1)
    pct = rnd.choice([25, 50, 50, 75, 90, 10]); adj = rnd.choice([-2.0, 0.0, 1.0, -5.0])
    P = pipeline.Params(percentage_identity=None, ani_percentile=pct, ani_adjustment=adj)
    ours = float(pipeline.estimate_identity(seqs[t0:t1], seqs[q0:q1], P)[0])
    os.dup2(null, 2)
    ref = float(aniutil.reference_identity(R, seqs[t0:t1], seqs[q0:q1], "#", pct, adj))
    os.dup2(saved, 2)
    n += 1
    if ours.hex() != ref.hex():
        bad += 1; print("MISMATCH", (q0, q1), (t0, t1), pct, adj, ours, ref, [(a, len(b)) for a, b in seqs], flush=True)
print(json.dumps({"cases": n, "mismatches": bad}))
