"""Shared test helpers. The oracle (oracle/liboracle.so) and the compiled reference (oracle/_ref) are
the CHECKERS; nothing here is imported by the product package."""
import ctypes
import gzip
import os
import random
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
GOLDEN_PEN = (4, 6, 2, 24, 1)   # penalties of the reference's wfa.utest vectors
WFMASH_PEN = (5, 8, 2, 24, 1)   # wflign.cpp:136-144


class Pen(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in "x o1 e1 o2 e2".split()]


class Counters(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int64) for n in "cells extend_matches overlap_tests score_steps".split()]


class MM(ctypes.Structure):
    _fields_ = [("hash", ctypes.c_uint64), ("wpos", ctypes.c_int64), ("wpos_end", ctypes.c_int64),
                ("seqId", ctypes.c_int32), ("strand", ctypes.c_int16), ("pad_", ctypes.c_int16)]


def load_oracle():
    path = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(path):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], check=True)
    lib = ctypes.CDLL(path)
    lib.orc_kmer_hash.restype = ctypes.c_uint64
    return lib


def load_ref(name):
    path = os.path.join(ROOT, "oracle", "_ref", name)
    if not os.path.exists(path):
        return None
    lib = ctypes.CDLL(path)
    if name == "libmapref.so":
        lib.ref_kmer_hash.restype = ctypes.c_uint64
        lib.ref_add_minmers.restype = ctypes.c_int64
    return lib


def orc_biwfa(orc, p, t, pen, counters=None):
    buf = ctypes.create_string_buffer(2 * (len(p) + len(t)) + 16)
    n, sc = ctypes.c_int(), ctypes.c_int()
    P = Pen(*pen)
    st = orc.orc_biwfa_align(p, len(p), t, len(t), ctypes.byref(P), buf, len(buf), ctypes.byref(n), ctypes.byref(sc),
                             ctypes.byref(counters) if counters is not None else None)
    return st, buf.raw[: n.value], sc.value


def orc_wfa(orc, p, t, pen):
    buf = ctypes.create_string_buffer(2 * (len(p) + len(t)) + 16)
    n, sc = ctypes.c_int(), ctypes.c_int()
    P = Pen(*pen)
    st = orc.orc_wfa_align(p, len(p), t, len(t), ctypes.byref(P), buf, len(buf), ctypes.byref(n), ctypes.byref(sc), None)
    return st, buf.raw[: n.value], sc.value


def ref_wfa(ref, p, t, pen, mode=3):
    buf = ctypes.create_string_buffer(2 * (len(p) + len(t)) + 16)
    n, sc = ctypes.c_int(), ctypes.c_int()
    st = ref.ref_wfa_end2end(p, len(p), t, len(t), *pen, mode, buf, len(buf), ctypes.byref(n), ctypes.byref(sc))
    return st, buf.raw[: n.value]


def rle(ops: bytes) -> str:
    out, i = [], 0
    while i < len(ops):
        j = i
        while j < len(ops) and ops[j] == ops[i]:
            j += 1
        out.append(f"{j - i}{chr(ops[i])}")
        i = j
    return "".join(out)


def golden_pairs():
    seqs = [l.rstrip("\n") for l in gzip.open(os.path.join(GOLD, "wfa_utest.seq.gz"), "rt")]
    return [(seqs[i][1:].encode(), seqs[i + 1][1:].encode()) for i in range(0, len(seqs), 2)]


def golden_alg(name):
    return [l.rstrip("\n").split("\t") for l in gzip.open(os.path.join(GOLD, name), "rt")]


def mutate(s, rate, rng):
    out = []
    for ch in s:
        r = rng.random()
        if r < rate * 0.6:
            out.append(rng.choice("ACGT"))
        elif r < rate * 0.8:
            L = min(int(rng.expovariate(1 / 3.0)) + 1, 60)
            out.append(ch)
            out.extend(rng.choice("ACGT") for _ in range(L))
        elif r < rate:
            continue
        else:
            out.append(ch)
    return "".join(out)


def random_pairs(n, seed, lengths=(1, 5, 50, 101, 150, 400, 1000, 3000), rates=(0.0, 0.005, 0.02, 0.08, 0.2, 0.4)):
    """Seeded pairs covering the edge cases the reference handles: empty sides, tiny (<=100, base-case
    path), ragged lengths, unrelated sequences, low-complexity alphabets, N runs."""
    rng = random.Random(seed)
    pairs = []
    for _ in range(n):
        L = rng.choice(lengths)
        rate = rng.choice(rates)
        alpha = "ACGTN"[: rng.choice([4, 4, 4, 5, 2])]
        a = "".join(rng.choice(alpha) for _ in range(L))
        b = mutate(a, rate, rng)
        if rng.random() < 0.1:
            b = b[rng.randrange(0, len(b) + 1):]
        if rng.random() < 0.1:
            a = a[: rng.randrange(0, len(a) + 1)]
        if rng.random() < 0.05:
            b = "".join(rng.choice("ACGT") for _ in range(rng.randrange(0, 300)))
        if rng.random() < 0.03:
            a = ""
        if rng.random() < 0.03:
            b = ""
        if not a and not b:
            a = "A"
        pairs.append((a.encode(), b.encode()))
    return pairs


# ---- a14: record-level do_biwfa_alignment (PAF) ------------------------------------------------------------------

def paf_records():
    """Deterministic mapping records for the do_biwfa_alignment parity tests (shared by tests/golden/make_paf_golden.py
    and the tests, so the committed fixture lines index into this list)."""
    from wfmash_b200 import synth
    recs = []
    seed = 100
    for (n, lo, hi, divs, pad) in [(12, 50, 400, (0.0, 0.01, 0.05, 0.15), 0), (10, 300, 1500, (0.01, 0.05, 0.1), 100),
                                   (6, 1000, 3000, (0.02, 0.1), 300)]:
        for (t, q, _d) in synth.mapping_records(n, seed, lo, hi, divs, pad=pad):
            i = len(recs)
            recs.append(dict(query_name=f"q{i}", target_name=f"t{i}", query=q, target=t, query_total_length=len(q) + 1000 + i,
                             query_offset=17 * i, target_total_length=len(t) + 5000, target_offset=31 * i, query_is_rev=(i % 3 == 0),
                             mashmap_estimated_identity=0.9 + 0.001 * i, chain_id=i, chain_length=(i % 4), chain_pos=1))
        seed += 1

    def mk(t, q):
        i = len(recs)
        recs.append(dict(query_name=f"Q{i}", target_name=f"T{i}", query=q, target=t, query_total_length=len(q) + 10, query_offset=5,
                         target_total_length=len(t) + 9, target_offset=3, query_is_rev=bool(i & 1), mashmap_estimated_identity=0.95,
                         chain_id=3, chain_length=2, chain_pos=1 + (i & 1)))
    rs = np.random.default_rng(99)

    def rseq(n):
        return synth.random_seq(n, rs).tobytes()
    s = rseq(3000)
    mk(s, s)                                        # identical: one long '=' run, the head patch spans the whole record
    mk(s[:200], s[:200])
    mk(b"ACG", b"ACG")                              # <= 3 bases: no patching (wflign.cpp:278)
    mk(b"ACGT", b"ACTT")
    mk(b"A", b"C")
    mk(rseq(500) + s[:600], s[:600])                # leading / trailing deletions and insertions
    mk(s[:600] + rseq(500), s[:600])
    mk(s[:600], rseq(300) + s[:600])
    mk(s[:600], s[:600] + rseq(300))
    mk(rseq(400), rseq(400))                        # unrelated
    mk(rseq(100), rseq(900))
    mk(s[:50] + s[300:1000], s[:1000])              # indels inside the eroded windows
    mk(s[:1000], s[:950] + s[960:1000])
    mk(s[:20] + rseq(6000) + s[20:700], s[:700])    # an indel longer than MAX_ERODE_LENGTH inside the head / tail window
    mk(s[:700], s[:680] + rseq(5000) + s[680:700])
    for _ in range(20):
        L = int(rs.integers(5, 1200))
        q = synth.random_seq(L, rs)
        t = synth.mutate(q, float(rs.choice([0.0, 0.02, 0.1, 0.3])), rs)
        padl, padr = int(rs.integers(0, 200)), int(rs.integers(0, 200))
        t = np.concatenate([synth.random_seq(padl, rs), t, synth.random_seq(padr, rs)]).tobytes()
        mk(t, q.tobytes())
    return recs


PAF_FILTER_SETS = [dict(), dict(disable_chain_patching=True), dict(min_identity=0.9, min_alignment_length=300, min_block_identity=0.8)]


def load_wflign_ref():
    """oracle/_ref/libwflignref.so: the unmodified reference wflign sources (do_biwfa_alignment) compiled in place."""
    path = os.path.join(ROOT, "oracle", "_ref", "libwflignref.so")
    if not os.path.exists(path):
        return None
    R = ctypes.CDLL(path)
    R.ref_do_biwfa_alignment.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int,
                                         ctypes.c_char_p, ctypes.c_char_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                         ctypes.c_uint64, ctypes.c_float, ctypes.c_uint64, ctypes.c_float, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_char_p, ctypes.c_int]
    return R


def ref_paf(R, r, pen=WFMASH_PEN, min_identity=0.0, min_alignment_length=0, min_block_identity=0.0, disable_chain_patching=False):
    cap = len(r["query"]) + len(r["target"]) + 4096
    buf = ctypes.create_string_buffer(cap)
    n = R.ref_do_biwfa_alignment(r["query_name"].encode(), r["query"], r.get("query_total_length", len(r["query"])),
                                 r.get("query_offset", 0), len(r["query"]), 1 if r.get("query_is_rev") else 0,
                                 r["target_name"].encode(), r["target"], r.get("target_total_length", len(r["target"])),
                                 r.get("target_offset", 0), len(r["target"]), pen[0], pen[1], pen[2], pen[3], pen[4],
                                 1 if disable_chain_patching else 0, min_identity, min_alignment_length, min_block_identity, 0,
                                 r.get("mashmap_estimated_identity", 0.0), r.get("chain_id", 0), r.get("chain_length", 0),
                                 r.get("chain_pos", 0), buf, cap)
    assert n >= 0
    return buf.raw[:n]


SAM_SETS = [dict(emit_md_tag=True), dict(emit_md_tag=False, no_seq_in_sam=True, min_alignment_length=32, min_block_identity=0.1),
            dict(emit_md_tag=True, disable_chain_patching=True)]


def ref_sam(R, r, pen=WFMASH_PEN, min_identity=0.0, min_alignment_length=0, min_block_identity=0.0, disable_chain_patching=False, emit_md_tag=False,
            no_seq_in_sam=False):
    """The SAM branch of the unmodified do_biwfa_alignment (oracle/ref_wflign_driver.cpp::ref_do_biwfa_alignment_sam)."""
    R.ref_do_biwfa_alignment_sam.argtypes = R.ref_do_biwfa_alignment.argtypes[:25] + [ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_int]
    cap = 3 * (len(r["query"]) + len(r["target"])) + 4096
    buf = ctypes.create_string_buffer(cap)
    n = R.ref_do_biwfa_alignment_sam(r["query_name"].encode(), r["query"], r.get("query_total_length", len(r["query"])),
                                     r.get("query_offset", 0), len(r["query"]), 1 if r.get("query_is_rev") else 0,
                                     r["target_name"].encode(), r["target"], r.get("target_total_length", len(r["target"])),
                                     r.get("target_offset", 0), len(r["target"]), pen[0], pen[1], pen[2], pen[3], pen[4],
                                     1 if disable_chain_patching else 0, min_identity, min_alignment_length, min_block_identity, 0,
                                     r.get("mashmap_estimated_identity", 0.0), r.get("chain_id", 0), r.get("chain_length", 0),
                                     r.get("chain_pos", 0), int(emit_md_tag), int(no_seq_in_sam), buf, cap)
    assert n >= 0
    return buf.raw[:n]


def sam_golden():
    import json
    with gzip.open(os.path.join(GOLD, "sam_do_biwfa.json.gz"), "rt") as f:
        return json.load(f)


def paf_golden():
    import json
    with gzip.open(os.path.join(GOLD, "paf_do_biwfa.json.gz"), "rt") as f:
        return json.load(f)
