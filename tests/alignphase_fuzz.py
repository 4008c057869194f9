"""Differential fuzz of the alignment phase's host side (wfb_align_phase: row parsing, query / target padding by chain position, clamped slices,
N-masking, strand correction, record batching, PAF / SAM re-emission; kernels under the host emulation build, WFB_LIB) against the reference's
UNMODIFIED align::Aligner::compute (oracle/_ref/libalignref.so): hand-made mapping rows at and around the sequence ends, both strands, shifted /
resized target intervals, random chain tags, PAF and SAM + MD. TEST INFRASTRUCTURE. python tests/alignphase_fuzz.py SEED SECONDS [MAX_CASES]"""
import json
import os
import random
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wfmash_b200 as wb
from wfmash_b200 import pipeline
from tests import pipeutil, util
A = util.load_ref("libalignref.so")
if A is None:
    print(json.dumps({"cases": 0, "rows": 0, "mismatches": 0, "skipped": "oracle/_ref/libalignref.so not built"}))
    sys.exit(0)
MAX_CASES = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 30
rnd = random.Random(int(sys.argv[1])); T_END = time.time() + float(sys.argv[2])
al = wb.Aligner(0)
n = bad = recs = 0
null = os.open(os.devnull, os.O_WRONLY); saved = os.dup(2)
while time.time() < T_END and bad < 3 and n < MAX_CASES:
    seqs = pipeutil.case(seed=rnd.randrange(1, 10**6), length=rnd.choice([6_000, 9_000, 14_000]))
    seqs = seqs[:3]
    P = pipeline.Params()
    rows = []
    for i in range(rnd.randint(1, 6)):
        qi, ti = rnd.sample(range(3), 2)
        (qn, q), (tn, t) = seqs[qi], seqs[ti]
        L = rnd.choice([40, 300, 1000, 1000, 2000, 3000])
        qs = rnd.choice([0, 0, rnd.randrange(0, max(1, len(q) - L)), max(0, len(q) - L)])
        qe = min(len(q), qs + L)
        shift = rnd.choice([0, 0, 3, -5, 40, -60])
        ts = min(max(0, qs + shift), max(0, len(t) - 1)); te = min(len(t), max(ts + 1, ts + (qe - qs) + rnd.choice([0, 0, 7, -9, 50])))
        strand = rnd.choice(["+", "+", "+", "-"])
        clen = rnd.randint(1, 4); cpos = rnd.randint(1, clen)
        ident = rnd.choice(["0.9", "0.95", "1", "0.7521"])
        rows.append("\t".join([qn, str(len(q)), str(qs), str(qe), strand, tn, str(len(t)), str(ts), str(te), "9", str(qe - qs), "20", "id:f:" + ident, "kc:f:1", f"ch:Z:{i + 1}.{clen}.{cpos}"]).encode())
    mp = b"\n".join(rows) + b"\n"
    sam = rnd.random() < 0.2
    ours, st = wb.align_phase(al, mp, seqs, seqs, sam_format=sam, emit_md_tag=sam)
    os.dup2(null, 2)
    ref = pipeutil.reference_align_phase(A, mp, seqs, P, sam_format=sam, emit_md_tag=sam)
    os.dup2(saved, 2)
    if sam: ref = b"".join(ln + b"\n" for ln in ref.split(b"\n") if ln and not ln.startswith(b"@"))
    n += 1; recs += len(rows)
    if sorted(ours.split(b"\n")) != sorted(ref.split(b"\n")):
        bad += 1
        a, b = set(ours.split(b"\n")), set(ref.split(b"\n"))
        print("MISMATCH sam" if sam else "MISMATCH", [r.decode()[:120] for r in rows], "\n only ours:", [x[:200] for x in sorted(a - b)[:2]], "\n only ref:", [x[:200] for x in sorted(b - a)[:2]], flush=True)
print(json.dumps({"cases": n, "rows": recs, "mismatches": bad}))
