"""Shared inputs / helpers of the mapping-path oracle tests and of tests/golden/make_map_golden.py (checker side only)."""
import ctypes
import hashlib
import random
import re

import numpy as np

MDT = np.dtype([("hash", "<u8"), ("wpos", "<i8"), ("wpos_end", "<i8"), ("seqId", "<i4"), ("strand", "<i2"), ("pad_", "<i2")])
IPDT = np.dtype([("pos", "<i8"), ("hash", "<u8"), ("seqId", "<i4"), ("side", "i1"), ("pad", "i1", (3,))])
L1DT = np.dtype([("seqId", "<i4"), ("pad", "<i4"), ("start", "<i8"), ("end", "<i8"), ("isz", "<i4"), ("pad2", "<i4")])
vp = lambda a: ctypes.c_void_p(a.ctypes.data)


def clean(seq: bytes) -> bytes:
    """makeUpperCaseAndValidDNA (commonFunc.hpp:132-142)"""
    return re.sub(rb"[^ACGT]", b"N", seq.upper())


def minmer_cases():
    """(name, seq, k, w, s): small enough for the reference addMinmers to be well defined (it indexes end() of its
    std::map on long i.i.d. sequences) and to run in seconds."""
    from wfmash_b200 import synth
    rng = np.random.default_rng(12)
    r1 = synth.random_seq(60000, rng).tobytes()
    unit = synth.random_seq(317, rng).tobytes()
    rep = synth.mutate(np.frombuffer(unit * 90, dtype=np.uint8), 0.02, rng).tobytes()
    nrun = b"N" * 3000 + r1[:6000] + b"N" * 50 + r1[7000:12000] + b"NNACGTNN" * 40
    lowc = bytes(random.Random(1).choice(b"AC") for _ in range(9000))
    out = []
    for name, sq in [("random", r1), ("tandem", rep), ("acgt", b"ACGT" * 400), ("nruns", nrun), ("lower", r1.lower()[:20000]), ("two-letter", lowc)]:
        out.append((name, sq, 15, 1000, 29))
    out += [("tandem-s59", rep, 15, 1000, 59), ("random-k19", r1[:40000], 19, 500, 17), ("tandem-k19", rep, 19, 500, 17), ("short-w", r1[:5000], 11, 200, 5)]
    return out


def orc_add_minmers(oracle, seq, k, w, s, sid):
    oracle.orc_add_minmers.restype = ctypes.c_int64
    cl = clean(seq)
    cap = len(cl) // 2 + 1000
    out = np.zeros(cap, dtype=MDT)
    n = oracle.orc_add_minmers(cl, ctypes.c_int64(len(cl)), k, w, s, sid, vp(out), ctypes.c_int64(cap))
    assert n <= cap
    return out[:n]


def ref_add_minmers(ref, seq, k, w, s, sid):
    ref.ref_add_minmers.restype = ctypes.c_int64
    buf = ctypes.create_string_buffer(seq, len(seq))  # the reference upper-cases / N-masks in place
    cap = len(seq) // 2 + 1000
    out = np.zeros(cap, dtype=MDT)
    n = ref.ref_add_minmers(buf, ctypes.c_int64(len(seq)), k, w, s, sid, vp(out), ctypes.c_int64(cap))
    assert n <= cap
    return out[:n]


def canonical(a: np.ndarray) -> np.ndarray:
    """addMinmers ends with an UNSTABLE std::sort on (wpos, wpos_end) (commonFunc.hpp:696): the order among records that
    tie on both is whatever libstdc++'s introsort leaves (seen on tandem repeats). Parity is therefore defined up to the
    order inside such tie groups; both sides are put in (wpos, wpos_end, hash) order before comparing. Nothing downstream
    depends on the tie order (Sketch::build groups by hash, mappingCore's lower_bound looks at (seqId, wpos) only)."""
    return a[np.lexsort((a["hash"], a["wpos_end"], a["wpos"]))]


def digest(a: np.ndarray, fields) -> str:
    h = hashlib.sha256()
    for f in fields:
        h.update(np.ascontiguousarray(a[f]).tobytes())
    return h.hexdigest()


def oracle_index(oracle, seqs, ids, k, w, s, F, threads):
    oracle.orc_index_build.restype = ctypes.c_int64
    mi = np.concatenate([orc_add_minmers(oracle, sq, k, w, s, sid) for sq, sid in zip(seqs, ids) if len(sq) >= w]).astype(MDT)
    n = len(mi)
    valid = [sid for sq, sid in zip(seqs, ids) if len(sq) >= w]
    chunk = -(-len(valid) // threads)
    part = np.zeros(max(ids) + 1, dtype=np.int32)
    for j, sid in enumerate(valid):
        part[sid] = j // chunk
    kept = np.zeros(n, dtype=MDT); pts = np.zeros(2 * n + 2, dtype=IPDT)
    uh = np.zeros(n, dtype=np.uint64); us = np.zeros(n, dtype=np.int64); uc = np.zeros(n, dtype=np.int64)
    npnt, nu, thr = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_uint64()
    nk = oracle.orc_index_build(vp(mi), ctypes.c_int64(n), vp(part), ctypes.c_double(F), vp(kept), vp(pts), ctypes.byref(npnt),
                                vp(uh), vp(us), vp(uc), ctypes.byref(nu), ctypes.byref(thr))
    return kept[:nk], pts[: npnt.value], uh[: nu.value], us[: nu.value], uc[: nu.value], thr.value


def l1_case(seed=32, scale=1):
    from wfmash_b200 import synth
    rng = np.random.default_rng(seed)
    root = synth.random_seq(60_000 * scale, rng)
    unit = synth.random_seq(400, rng)
    rep = synth.mutate(np.tile(unit, 60), 0.03, rng)
    seqs = [root.tobytes(), synth.mutate(root, 0.02, rng).tobytes(), synth.mutate(root, 0.08, rng).tobytes(), rep.tobytes(),
            (root[:20000].tobytes() + rep[:12000].tobytes()), synth.mutate(root, 0.15, rng).tobytes()[:35000], b"ACGT" * 50]
    ids = list(range(len(seqs)))
    groups = [0, 0, 1, 2, 2, 3, 4]
    return seqs, ids, groups


L1_MODES = [(1, 1, 0, 3), (0, 0, 0, 2), (0, 1, 1, 5), (0, 0, 1, 12)]  # skip_self, skip_prefix, lower_triangular, minimum_hits


def l1_all_fragments(lib, fn, index, seqs, ids, groups, k, w, s, oracle):
    """Runs lib.<fn> (orc_l1_fragment or ref_l1_fragment: same signature) over every fragment of every sequence and
    every mode of L1_MODES; the fragment sketch comes from the oracle's sketchSequence. Returns a flat int64 array
    (mode, fragment, seqId, start, end, intersectionSize) rows."""
    kept, pts, uh, us, uc, _ = index
    cut = np.array([max(1, int(i * 0.5)) for i in range(1001)], dtype=np.int32)
    grp = np.array(groups, dtype=np.int32)
    rows = []
    f = getattr(lib, fn)
    for m, (ss, sp, lt, mh) in enumerate(L1_MODES):
        fi = 0
        for qi, sq in enumerate(seqs):
            starts = [j * w for j in range(len(sq) // w)]
            if len(sq) >= w and len(sq) % w:
                starts.append(len(sq) - w)  # overlapping tail fragment (computeMap.hpp:602-631)
            for st in starts:
                frag = clean(sq[st:st + w])
                q = np.zeros(s + 8, dtype=MDT)
                qn = oracle.orc_sketch_fragment(frag, w, k, s, ids[qi], vp(q))
                qh = np.ascontiguousarray(q["hash"][:qn])
                o = np.zeros(512, dtype=L1DT)
                n = f(vp(uh), vp(us), vp(uc), ctypes.c_int64(len(uh)), vp(pts), vp(qh), qn, ids[qi], groups[qi], vp(grp), ss, sp, lt, mh, s, w,
                      vp(cut), len(cut), vp(o), 512)
                for j in range(n):
                    rows.append((m, fi, int(o["seqId"][j]), int(o["start"][j]), int(o["end"][j]), int(o["isz"][j])))
                fi += 1
    return np.array(rows, dtype=np.int64).reshape(-1, 6)
