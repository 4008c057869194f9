#!/usr/bin/env python3
"""Regenerates tests/golden/highdiv_reference.json.gz: BASELINE's C5 regime (full-length mapping records at 10 % and 20 %
divergence, scores of 2-6 x 10^4) aligned by the UNMODIFIED reference (WFA2-lib biWFA, memory mode ultralow, wfmash's penalties
0,5,8,2,24,1 — oracle/_ref/libwfa2ref.so). The records are seeded synthetic ones (wfmash_b200.synth.mapping_records), so only
their digests are stored: sha256 of the operation string, its length and the score. tests/test_gpu_parity.py regenerates the same
records on the GPU box and compares. Run in the build container only (needs oracle/_ref)."""
import ctypes, gzip, hashlib, json, os, sys
from concurrent.futures import ThreadPoolExecutor
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from wfmash_b200 import synth  # noqa: E402

SETS = [dict(n=6, seed=501, len_lo=48000, len_hi=50000, divergences=[0.10]),
        dict(n=6, seed=502, len_lo=48000, len_hi=50000, divergences=[0.20]),
        dict(n=4, seed=503, len_lo=20000, len_hi=30000, divergences=[0.15], pad=5000)]

if __name__ == "__main__":
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libwfa2ref.so"))

    def one(rec):
        p, t, d = rec
        buf = ctypes.create_string_buffer(2 * (len(p) + len(t)) + 16)
        n, sc = ctypes.c_int(), ctypes.c_int()
        st = lib.ref_wfa_end2end(p, len(p), t, len(t), 5, 8, 2, 24, 1, 3, buf, len(buf), ctypes.byref(n), ctypes.byref(sc))
        assert st == 0
        return dict(plen=len(p), tlen=len(t), divergence=d, ops_len=n.value, score=sc.value, sha=hashlib.sha256(buf.raw[: n.value]).hexdigest())

    out = []
    for s in SETS:
        recs = synth.mapping_records(**s)
        with ThreadPoolExecutor(8) as ex:
            out.append(dict(params=s, records=list(ex.map(one, recs))))
    with gzip.GzipFile(os.path.join(HERE, "highdiv_reference.json.gz"), "wb", mtime=0) as f:
        f.write(json.dumps(out).encode())
    for s in out:
        print(s["params"], [(-r["score"], r["ops_len"]) for r in s["records"]])
