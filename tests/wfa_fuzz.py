"""Differential fuzz of the biWFA kernels (run under the host emulation build, WFB_LIB) against the oracle's restatement of WFA2-lib's
gap-affine-2-piece biWFA: empty / one-base / base-case-boundary lengths, one- and two-letter alphabets, tandem repeats inserted on one or both
sides, substitutions and indels of 1 - 120 bases at 0 - 50 %. TEST INFRASTRUCTURE. python tests/wfa_fuzz.py SEED SECONDS [MAX_PAIRS]"""
import ctypes
import os
import random
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wfmash_b200 as wb
from tests import util
orc = util.load_oracle()
class Pen(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in "x o1 e1 o2 e2".split()]
rnd = random.Random(int(sys.argv[1]))
T_END = time.time() + float(sys.argv[2])
MAX_PAIRS = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 30
pen = Pen(*wb.WFMASH_PENALTIES)
def rseq(n, alpha=b"ACGT"):
    return bytes(rnd.choice(alpha) for _ in range(n))
def mutate(s, d):
    out = bytearray()
    i = 0
    while i < len(s):
        r = rnd.random()
        if r < d * 0.5: out.append(rnd.choice(b"ACGT")); i += 1
        elif r < d * 0.75: out.extend(rseq(rnd.choice([1, 1, 2, 3, 10, 40, 120]))) 
        elif r < d: i += rnd.choice([1, 1, 2, 3, 10, 40, 120])
        else: out.append(s[i]); i += 1
    return bytes(out)
al = wb.Aligner(0)
n = bad = 0
while time.time() < T_END and n < MAX_PAIRS:
    pairs = []
    for _ in range(24):
        kind = rnd.random()
        L = rnd.choice([0, 1, 2, 5, 30, 99, 100, 101, 200, 300, 600, 1500, 3000])
        alpha = rnd.choice([b"ACGT", b"ACGT", b"AC", b"A"])
        a = rseq(L, alpha)
        if kind < 0.1: b = rseq(rnd.choice([0, 1, 50, L]), alpha)
        else: b = mutate(a, rnd.choice([0.0, 0.01, 0.05, 0.15, 0.3, 0.5]))
        if rnd.random() < 0.3:
            u = rseq(rnd.randint(1, 12)); rep = u * rnd.randint(5, 60)
            p = rnd.randint(0, len(a)); a = a[:p] + rep + a[p:]
            if rnd.random() < 0.7:
                q = rnd.randint(0, len(b)); b = b[:q] + u * rnd.randint(3, 70) + b[q:]
        pairs.append((a, b))
    res = al.align_end2end_batch(pairs)
    for (p, t), r in zip(pairs, res):
        buf = ctypes.create_string_buffer(2 * (len(p) + len(t)) + 16)
        nn, sc = ctypes.c_int(), ctypes.c_int()
        st = orc.orc_biwfa_align(p, len(p), t, len(t), ctypes.byref(pen), buf, len(buf), ctypes.byref(nn), ctypes.byref(sc), None)
        n += 1
        ok = (st == 0) == (r.status == 0) and (st != 0 or (buf.raw[: nn.value] == r.ops and sc.value == r.score))
        if not ok:
            bad += 1
            print("MISMATCH", len(p), len(t), st, r.status, sc.value, getattr(r, 'score', None))
            if bad > 3: break
    if bad > 3: break
import json
print(json.dumps({"pairs": n, "mismatches": bad}))
