"""Checker side of the ANI auto-identity tests: sequences, the oracle's literal StreamingMinHash restatement
composed per group like map_stats.hpp:617-637, and the reference's UNMODIFIED estimate_identity_for_groups
(oracle/_ref/libstatsref.so). TEST INFRASTRUCTURE ONLY."""
import ctypes
import tempfile

import numpy as np


def case(seed=3, length=40_000, small=False):
    """Five haplotypes in four PanSN groups at 1-12 % divergence, soft-masked stretches, N runs (one inside the first k bases
    of a sequence: the reference's head rule), a satellite array (duplicated hashes) and a sequence shorter than k."""
    from wfmash_b200 import synth
    rng = np.random.default_rng(seed)
    root = synth.random_seq(length, rng)
    unit = synth.random_seq(171, rng)

    def hap(d):
        return synth.mutate(root, d, rng).tobytes()

    a = bytearray(hap(0.0)); a[5000:5400] = b"N" * 400; a[9000:12000] = bytes(a[9000:12000]).lower()
    b = bytearray(hap(0.01)); b[7] = ord("n")
    c = hap(0.05) + np.tile(unit, 60 if small else 400).tobytes()
    d = bytearray(hap(0.12)); d[100:130] = b"RYKM" * 7 + b"XX"
    e = hap(0.03)[: length // 2]
    if not small:
        # a planted k-mer with a very small canonical hash, 3000 copies between random spacers: one hash value that is far
        # more frequent than the candidate capacity of the small-sketch runs (the selection's capacity re-run)
        from tests import util
        orc = util.load_oracle()
        orc.orc_kmer_hash.restype = ctypes.c_uint64
        cands = [synth.random_seq(21, rng).tobytes() for _ in range(4000)]
        comp = bytes.maketrans(b"ACGT", b"TGCA")
        best = min(cands, key=lambda km: min(orc.orc_kmer_hash(km, 21), orc.orc_kmer_hash(km.translate(comp)[::-1], 21)))
        e = e + b"".join(best + synth.random_seq(10, rng).tobytes() for _ in range(3000))
    return [("hapA#1#chr1", bytes(a)), ("hapB#1#chr1", bytes(b)), ("hapC#1#chr1", c), ("hapD#2#chr1", bytes(d)), ("hapC#1#chr2", e), ("hapD#2#tiny", b"ACGTACGTAC")]


def oracle_group_sketches(oracle, seqs, seq_group, n_groups, k=21, s=4096):
    """Per-sequence heaps, merged per group hash by hash, exactly like the reference does it."""
    oracle.orc_ani_add_sequence.argtypes = [ctypes.c_char_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    oracle.orc_ani_add_hash.argtypes = [ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    heaps = [np.zeros(s, dtype=np.uint64) for _ in range(n_groups)]
    sizes = [0] * n_groups
    for seq, g in zip(seqs, seq_group):
        h = np.zeros(s, dtype=np.uint64)
        n = oracle.orc_ani_add_sequence(seq, len(seq), k, s, ctypes.c_void_p(h.ctypes.data), 0)
        for v in np.sort(h[:n]):  # getSketch(): ascending
            sizes[g] = oracle.orc_ani_add_hash(int(v), s, ctypes.c_void_p(heaps[g].ctypes.data), sizes[g])
    sk = np.zeros((n_groups, s), dtype=np.uint64)
    for g in range(n_groups):
        sk[g, : sizes[g]] = np.sort(heaps[g][: sizes[g]])
    return sk, np.array(sizes, dtype=np.int32)


def reference_identity(R, targets, queries, delim="#", percentile=50, adjustment=-2.0, threads=2):
    R.ref_estimate_identity.restype = ctypes.c_double
    same = [n for n, _ in targets] == [n for n, _ in queries]

    def arrs(seqs):
        n = len(seqs)
        return ((ctypes.c_char_p * n)(*[a.encode() for a, _ in seqs]), (ctypes.c_char_p * n)(*[b for _, b in seqs]), (ctypes.c_int64 * n)(*[len(b) for _, b in seqs]), n)

    q, t = arrs(queries), arrs(targets)
    with tempfile.TemporaryDirectory() as d:
        return R.ref_estimate_identity(d.encode(), q[0], q[1], q[2], q[3], t[0], t[1], t[2], t[3], int(same), delim.encode(), percentile,
                                       ctypes.c_float(adjustment), threads)
