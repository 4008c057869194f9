"""Differential fuzz of the record path (wfb_biwfa_paf_batch: main biWFA, head / tail ends-free patches, erosion, swizzles, trimming, PAF
metrics; kernels under the host emulation build, WFB_LIB) against the reference's UNMODIFIED do_biwfa_alignment (oracle/_ref/libwflignref.so,
compiled in place: AVX2 build, hence term_group 8). Records whose ENDS are what patching and the swizzles act on: unrelated flanks, truncated
ends, tandem repeats of different copy number on the two sides, N runs, both strands, four filter sets. TEST INFRASTRUCTURE.
    python tests/paf_fuzz.py SEED SECONDS [MAX_RECORDS]"""
import json
import os
import random
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wfmash_b200 as wb
from tests import util
R = util.load_wflign_ref()
if R is None:
    print(json.dumps({"records": 0, "mismatches": 0, "skipped": "oracle/_ref/libwflignref.so not built"}))
    sys.exit(0)
rnd = random.Random(int(sys.argv[1]))
T_END = time.time() + float(sys.argv[2])
MAX_RECORDS = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 30
def rseq(n, alpha=b"ACGT"):
    return bytes(rnd.choice(alpha) for _ in range(n))
def mutate(s, d):
    out = bytearray(); i = 0
    while i < len(s):
        r = rnd.random()
        if r < d * 0.6: out.append(rnd.choice(b"ACGT")); i += 1
        elif r < d * 0.8: out.extend(rseq(rnd.choice([1, 1, 2, 3, 10, 40])))
        elif r < d: i += rnd.choice([1, 1, 2, 3, 10, 40])
        else: out.append(s[i]); i += 1
    return bytes(out)
al = wb.Aligner(0)
n = bad = 0
sets = [dict(), dict(min_alignment_length=32, min_block_identity=0.1), dict(disable_chain_patching=True), dict(min_identity=0.8)]
while time.time() < T_END and bad < 4 and n < MAX_RECORDS:
    recs = []
    for i in range(16):
        L = rnd.choice([40, 120, 300, 700, 1500, 3000])
        q = rseq(L, rnd.choice([b"ACGT", b"ACGT", b"ACGT", b"AC"]))
        t = mutate(q, rnd.choice([0.0, 0.01, 0.05, 0.1, 0.2, 0.3]))
        # ends: unrelated flanks / repeats / deletions of an end, which is what head / tail patching and the swizzles act on
        e = rnd.random()
        if e < 0.25: t = rseq(rnd.randint(1, 400)) + t + rseq(rnd.randint(1, 400))
        elif e < 0.4: t = t[rnd.randint(0, min(200, len(t))):]
        elif e < 0.55: q = q[: max(1, len(q) - rnd.randint(0, 200))]
        elif e < 0.7:
            u = rseq(rnd.randint(1, 6)); k = rnd.randint(3, 40)
            q = u * k + q; t = u * rnd.randint(1, 45) + t
        elif e < 0.8:
            u = rseq(rnd.randint(1, 6)); q = q + u * rnd.randint(3, 40); t = t + u * rnd.randint(1, 45)
        if rnd.random() < 0.1: q = q[:5] + b"N" * rnd.randint(1, 8) + q[5:]
        if not q or not t: continue
        recs.append(dict(query_name=f"q{i}", target_name=f"t{i}", query=q, target=t, query_total_length=len(q) + rnd.randint(0, 2000),
                         query_offset=rnd.randint(0, 500), target_total_length=len(t) + rnd.randint(0, 5000), target_offset=rnd.randint(0, 800),
                         query_is_rev=rnd.random() < 0.4, mashmap_estimated_identity=round(rnd.uniform(0.7, 1.0), 3), chain_id=i, chain_length=rnd.randint(1, 4), chain_pos=1))
    for r in recs:
        r["query_total_length"] = max(r["query_total_length"], r["query_offset"] + len(r["query"]))
        r["target_total_length"] = max(r["target_total_length"], r["target_offset"] + len(r["target"]))
    kw = rnd.choice(sets)
    lines, status = al.biwfa_paf_batch(recs, term_group=8, **kw)
    for i, (r, g) in enumerate(zip(recs, lines)):
        w = util.ref_paf(R, r, **kw)
        n += 1
        if g != w:
            bad += 1
            print("MISMATCH", kw, len(r["query"]), len(r["target"]), r["query_is_rev"], "\n ours:", g[:300], "\n ref :", w[:300], flush=True)
print(json.dumps({"records": n, "mismatches": bad}))
