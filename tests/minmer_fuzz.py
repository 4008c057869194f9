"""Differential fuzz of the reference-side minmer build against the oracle's exact addMinmers restatement, run in a subprocess by
tests/test_minmer_emu_cpu.py with WFB_LIB pointing at the host emulation build (TEST INFRASTRUCTURE): random (k, w, s), sequence lengths on and
around the window / tile / chunk boundaries, random / two-letter / periodic / homopolymer-mixed / N-sprinkled / lower-case sequences.
    python tests/minmer_fuzz.py SEED SECONDS [MAX_CASES]"""
import os
import random
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wfmash_b200 as wb
from tests import maputil, util
oracle = util.load_oracle()
import ctypes
def orc(seq, k, w, s, sid):
    oracle.orc_add_minmers.restype = ctypes.c_int64
    cl = maputil.clean(seq)
    cap = len(cl) * 6 + 1000
    out = np.zeros(cap, dtype=maputil.MDT)
    n = oracle.orc_add_minmers(cl, ctypes.c_int64(len(cl)), k, w, s, sid, maputil.vp(out), ctypes.c_int64(cap))
    assert n <= cap
    return out[:n]
rnd = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
T_END = time.time() + (float(sys.argv[2]) if len(sys.argv) > 2 else 60)
MAX_CASES = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 30
def mkseq(n):
    kind = rnd.choice(["rand", "rand", "rand", "lowc", "rep", "mix", "nruns"])
    if kind == "rand":
        s = bytes(rnd.choice(b"ACGT") for _ in range(n))
    elif kind == "lowc":
        s = bytes(rnd.choice(b"AC") for _ in range(n))
    elif kind == "rep":
        u = bytes(rnd.choice(b"ACGT") for _ in range(rnd.choice([1, 2, 3, 7, 50, 317])))
        s = (u * (n // len(u) + 1))[:n]
    elif kind == "mix":
        parts = []
        while sum(map(len, parts)) < n:
            m = rnd.randint(1, 3000)
            c = rnd.choice(["r", "r", "a", "n", "u"])
            if c == "r": parts.append(bytes(rnd.choice(b"ACGT") for _ in range(m)))
            elif c == "a": parts.append(bytes([rnd.choice(b"ACGT")]) * m)
            elif c == "n": parts.append(b"N" * rnd.randint(1, 60))
            else:
                u = bytes(rnd.choice(b"ACGT") for _ in range(rnd.randint(2, 40))); parts.append((u * (m // len(u) + 1))[:m])
        s = b"".join(parts)[:n]
    else:
        b = bytearray(rnd.choice(b"ACGT") for _ in range(n))
        for _ in range(rnd.randint(1, 12)):
            p = rnd.randrange(n); b[p:p + rnd.randint(1, 40)] = b"N" * min(rnd.randint(1, 40), n - p)
        s = bytes(b[:n])
    if rnd.random() < 0.2:
        s = s.lower()
    return s
n_cases = n_rec = n_redo = filt = 0
bad = 0
while time.time() < T_END and n_cases < MAX_CASES:
    k = rnd.choice([11, 15, 15, 15, 16, 17, 19, 21, 31, 32])
    w = rnd.choice([200, 250, 500, 1000, 1000, 1000, 2000])
    s = rnd.choice([1, 2, 5, 17, 24, 29, 39, 59, 100])
    if w <= k: continue
    lens = []
    for _ in range(rnd.randint(1, 4)):
        base = rnd.choice([w, w + 1, 2304, 2304 + k - 1, 2 * 2304 + k - 1, 1024 + k - 1, 2048 + k - 1, 4608, rnd.randint(w, 30000)])
        lens.append(max(1, base + rnd.choice([-1, 0, 0, 1, 2])))
    seqs = [mkseq(n) for n in lens]
    ids = [3 + 2 * i for i in range(len(seqs))]
    try:
        got, st = wb.minmers_build(seqs, ids, k, w, s)
    except wb.WfbError as e:
        if "too small" in str(e) or "capacity" in str(e):
            continue
        raise
    exp = [orc(x, k, w, s, sid) for x, sid in zip(seqs, ids) if len(x) >= w]
    exp = np.concatenate(exp) if exp else np.zeros(0, dtype=got.dtype)
    ok = len(got) == len(exp) and all((got[f] == exp[f]).all() for f in ("hash", "wpos", "wpos_end", "seqId", "strand")) and st.stitch_miss == 0
    n_cases += 1; n_rec += len(exp); n_redo += st.redo_chunks; filt += st.filtered
    if not ok:
        print("MISMATCH", k, w, s, lens, st.as_dict(), len(got), len(exp))
        bad += 1
        break
import json
print(json.dumps({"cases": n_cases, "records": int(n_rec), "redo_chunks": int(n_redo), "filtered_builds": int(filt), "mismatches": bad}))
