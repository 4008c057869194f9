"""Shared test helpers. The oracle (oracle/liboracle.so) and the compiled reference (oracle/_ref) are
the CHECKERS; nothing here is imported by the product package."""
import ctypes
import gzip
import os
import random
import subprocess

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
