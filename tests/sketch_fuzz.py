"""Differential fuzz of the query-fragment sketch kernel body (wfb_sketch_fragments under the host emulation build, WFB_LIB) against the oracle's
sketchSequence restatement: random (k, w, s), overlapping fragments, random / two-letter / periodic / mixed-case-with-N sequences.
TEST INFRASTRUCTURE. python tests/sketch_fuzz.py SEED SECONDS [MAX_FRAGMENTS]"""
import ctypes
import json
import os
import random
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wfmash_b200 as wb
from tests import util
orc = util.load_oracle()
class MM(ctypes.Structure):
    _fields_ = wb.Minmer._fields_
rnd = random.Random(int(sys.argv[1])); T_END = time.time() + float(sys.argv[2])
MAX_N = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 30
n = bad = 0
while time.time() < T_END and bad < 3 and n < MAX_N:
    k = rnd.choice([11, 15, 15, 16, 19, 21, 31, 32]); w = rnd.choice([200, 500, 1000, 1000, 2000]); s = rnd.choice([1, 5, 17, 29, 59, 100])
    nf = rnd.randint(1, 6)
    total = w * nf + rnd.randint(0, 50)
    kind = rnd.random()
    if kind < 0.5: seq = bytes(rnd.choice(b"ACGT") for _ in range(total))
    elif kind < 0.65: seq = bytes(rnd.choice(b"AC") for _ in range(total))
    elif kind < 0.8:
        u = bytes(rnd.choice(b"ACGT") for _ in range(rnd.choice([1, 2, 3, 7, 40]))); seq = (u * (total // len(u) + 1))[:total]
    else:
        b = bytearray(rnd.choice(b"ACGTacgtNn") for _ in range(total)); seq = bytes(b)
    frags = []
    for j in range(nf):
        off = rnd.choice([j * w, rnd.randint(0, total - w)])
        frags.append((off, w, rnd.randint(0, 9)))
    mm, cnt, _ = wb.sketch_fragments(seq, frags, k, s)
    for j, (off, ln, sid) in enumerate(frags):
        b = (MM * max(s, 1))()
        fr = bytes(seq[off: off + ln]).upper()
        fr = bytes(c if c in b"ACGT" else ord("N") for c in fr)
        nb = orc.orc_sketch_fragment(fr, ln, k, s, sid, b)
        got = [(int(x["hash"]), int(x["wpos"]), int(x["wpos_end"]), int(x["seqId"]), int(x["strand"])) for x in mm[j, :cnt[j]]]
        exp = [(x.hash, x.wpos, x.wpos_end, x.seqId, x.strand) for x in b[:nb]]
        n += 1
        if nb != cnt[j] or got != exp:
            bad += 1; print("MISMATCH", k, w, s, off, nb, cnt[j], flush=True)
print(json.dumps({"fragments": n, "mismatches": bad}))
