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
    cap = len(cl) * 4 + 1000  # degenerate (two-letter, periodic) sequences emit several records per base
    out = np.zeros(cap, dtype=MDT)
    n = oracle.orc_add_minmers(cl, ctypes.c_int64(len(cl)), k, w, s, sid, vp(out), ctypes.c_int64(cap))
    assert n <= cap
    return out[:n]


def ref_add_minmers(ref, seq, k, w, s, sid):
    ref.ref_add_minmers.restype = ctypes.c_int64
    buf = ctypes.create_string_buffer(seq, len(seq))  # the reference upper-cases / N-masks in place
    cap = len(seq) * 4 + 1000
    out = np.zeros(cap, dtype=MDT)
    n = ref.ref_add_minmers(buf, ctypes.c_int64(len(seq)), k, w, s, sid, vp(out), ctypes.c_int64(cap))
    assert n <= cap
    return out[:n]


def canonical(a: np.ndarray) -> np.ndarray:
    """Records of one sequence in (wpos, wpos_end, hash) order. addMinmers ends with an UNSTABLE std::sort on (wpos, wpos_end)
    (commonFunc.hpp:696): the order among records that tie on both is whatever libstdc++'s introsort leaves. The committed digests
    (tests/golden/map_reference.json.gz) were taken in this canonical order; since the end of round 2 the oracle and the library reproduce
    the reference's ACTUAL tie order (it matters: computeL2MappedRegions evaluates the sketch after every insertion, visible for targets of
    w .. 2w bases), and the live tests compare the exact order."""
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


L2DT = np.dtype([("seqId", "<i4"), ("shared", "<i4"), ("mean", "<i8"), ("start", "<i8"), ("end", "<i8"), ("strand", "<i4"), ("pad", "<i4")])
L2MAPDT = np.dtype([("frag", "<i4"), ("refSeqId", "<i4"), ("refStartPos", "<i8"), ("optimalStart", "<i8"), ("optimalEnd", "<i8"),
                    ("conservedSketches", "<i4"), ("strand", "<i4"), ("nucIdentity", "<f4"), ("kmerComplexity", "<f4")])


def fragments_of(seqs, w):
    """(query index, start) of every fragment as Map::mapQuery cuts them (computeMap.hpp:565-631): len//w full
    fragments plus one overlapping tail fragment."""
    out = []
    for qi, sq in enumerate(seqs):
        starts = [j * w for j in range(len(sq) // w)]
        if len(sq) >= w and len(sq) % w:
            starts.append(len(sq) - w)
        out += [(qi, st) for st in starts]
    return out


def l2_all_loci(lib, kind, index, seqs, ids, groups, k, w, s, oracle, mode=(1, 1, 0, 3)):
    """computeL2MappedRegions over every L1 locus (oracle L1, one filter mode) of every fragment, through `lib`:
    kind = "orc" (orc_l2_locus) or "ref" (ref_l2_locus of oracle/_ref/libl2ref.so). Returns int64 rows
    (fragment, locus, seqId, sharedSketchSize, meanOptimalPos, optimalStart, optimalEnd, strand)."""
    kept, pts, uh, us, uc, _ = index
    kept = np.ascontiguousarray(kept)
    cut = np.array([max(1, int(i * 0.5)) for i in range(1001)], dtype=np.int32)
    grp = np.array(groups, dtype=np.int32)
    ss, sp, lt, mh = mode
    handle = None
    if kind == "ref":
        lib.ref_l2_open.restype = ctypes.c_void_p
        handle = ctypes.c_void_p(lib.ref_l2_open(vp(kept), ctypes.c_int64(len(kept))))
    rows = []
    dup = ctypes.c_int64(0)
    for fi, (qi, st) in enumerate(fragments_of(seqs, w)):
        frag = clean(seqs[qi][st:st + w])
        q = np.zeros(s + 8, dtype=MDT)
        qn = oracle.orc_sketch_fragment(frag, w, k, s, ids[qi], vp(q))
        qh = np.ascontiguousarray(q["hash"][:qn])
        o = np.zeros(512, dtype=L1DT)
        n = oracle.orc_l1_fragment(vp(uh), vp(us), vp(uc), ctypes.c_int64(len(uh)), vp(pts), vp(qh), qn, ids[qi], groups[qi], vp(grp),
                                   ss, sp, lt, mh, s, w, vp(cut), len(cut), vp(o), 512)
        for j in range(n):
            out = np.zeros(256, dtype=L2DT)
            a = (int(o["seqId"][j]), ctypes.c_int64(int(o["start"][j])), ctypes.c_int64(int(o["end"][j])))
            if kind == "ref":
                m = lib.ref_l2_locus(handle, vp(q), qn, w, a[0], a[1], a[2], vp(out), 256)
            else:
                m = lib.orc_l2_locus(vp(kept), ctypes.c_int64(len(kept)), vp(q), qn, w, a[0], a[1], a[2], vp(out), 256, None, ctypes.byref(dup))
            assert m <= 256
            for t in range(m):
                x = out[t]
                rows.append((fi, j, int(x["seqId"]), int(x["shared"]), int(x["mean"]), int(x["start"]), int(x["end"]), int(x["strand"])))
    if kind == "ref":
        lib.ref_l2_close(handle)
    else:
        assert dup.value == 0, "a matching minmer was inserted into an already active SlideMapper slot"
    return np.array(rows, dtype=np.int64).reshape(-1, 8)


def revcomp(b: bytes) -> bytes:
    return b.translate(bytes.maketrans(b"ACGTacgt", b"TGCAtgca"))[::-1]


def l2_case(seed=41, scale=1):
    """l1_case plus reverse-complemented and rearranged copies so that the L2 strand vote takes both signs."""
    from wfmash_b200 import synth
    seqs, ids, groups = l1_case(seed, scale)
    rng = np.random.default_rng(seed + 1000)
    root = np.frombuffer(seqs[0], dtype=np.uint8)
    rc = revcomp(synth.mutate(root[5000:45000], 0.04, rng).tobytes())
    mixed = synth.mutate(root[:15000], 0.03, rng).tobytes() + revcomp(synth.mutate(root[20000:38000], 0.06, rng).tobytes()) + seqs[3][:7000]
    seqs = seqs + [rc, mixed]
    ids = list(range(len(seqs)))
    groups = groups + [5, 6]
    return seqs, ids, groups


def stage1_table(oracle, hg, ani_diff, k, s):
    """wfb_stage1_min_hits' expected output from the oracle's orc_stage1_pass."""
    oracle.orc_stage1_pass.argtypes = [ctypes.c_double, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    t = np.zeros(s + 1, dtype=np.int32)
    for qs in range(1, s + 1):
        v = 0
        while v <= qs and not oracle.orc_stage1_pass(hg, ani_diff, k, qs, v):
            v += 1
        t[qs] = v
    return t


def oracle_map_fragments(oracle, index, seqs, ids, groups, k, w, s, mode=(1, 1, 0, 3), stage1=True, min_shared=None, hg=1.0, ani_diff=0.0, cut=None, ref_group=None):
    """Map::mapSingleQueryFrag's L1 + L2 stages through the oracle for every fragment of every sequence. Returns
    (frag_list, q_all[n, s] MDT, q_count[n], loci rows (frag, seqId, start, end, isz), mappings L2MAPDT sorted by
    (frag, refSeqId, refStartPos))."""
    kept, pts, uh, us, uc, _ = index
    kept = np.ascontiguousarray(kept)
    cut = np.array([max(1, int(i * 0.5)) for i in range(1001)], dtype=np.int32) if cut is None else np.ascontiguousarray(cut, dtype=np.int32)
    # groups[qi] = group of query qi; the table indexed by target seqId is the same list unless the queries are not the targets
    grp = np.array(groups if ref_group is None else ref_group, dtype=np.int32)
    ss, sp, lt, mh = mode
    oracle.orc_l2_fragment.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_float, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    frs = fragments_of(seqs, w)
    q_all = np.zeros((len(frs), s), dtype=MDT)
    q_count = np.zeros(len(frs), dtype=np.int32)
    loci_rows, maps = [], []
    for fi, (qi, st) in enumerate(frs):
        frag = clean(seqs[qi][st:st + w])
        q = np.zeros(s + 8, dtype=MDT)
        qn = oracle.orc_sketch_fragment(frag, w, k, s, ids[qi], vp(q))
        q_all[fi, :qn] = q[:qn]
        q_count[fi] = qn
        if qn == 0:
            continue
        qh = np.ascontiguousarray(q["hash"][:qn])
        o = np.zeros(512, dtype=L1DT)
        n = oracle.orc_l1_fragment(vp(uh), vp(us), vp(uc), ctypes.c_int64(len(uh)), vp(pts), vp(qh), qn, ids[qi], groups[qi], vp(grp),
                                   ss, sp, lt, mh, s, w, vp(cut), len(cut), vp(o), 512)
        for j in range(n):
            loci_rows.append((fi, int(o["seqId"][j]), int(o["start"][j]), int(o["end"][j]), int(o["isz"][j])))
        out = np.zeros(1024, dtype=L2MAPDT)
        kc = np.float32((np.float64(qn) / np.float64(np.longdouble(int(q["hash"][qn - 1])) / np.longdouble(18446744073709551615))) / ((w - k + 1) * 2))
        m = oracle.orc_l2_fragment(vp(kept), len(kept), vp(q), qn, float(kc), k, w, vp(o), n, int(stage1), hg, ani_diff,
                                   int(min_shared[qn]) if min_shared is not None else 0, vp(out), 1024)
        assert m <= 1024
        out["frag"][:m] = fi
        maps.append(out[:m].copy())
    mp = np.concatenate(maps) if maps else np.zeros(0, dtype=L2MAPDT)
    return frs, q_all, q_count, np.array(loci_rows, dtype=np.int64).reshape(-1, 5), mp

L1PUBDT = np.dtype([("seqId", "<i4"), ("intersectionSize", "<i4"), ("rangeStartPos", "<i8"), ("rangeEndPos", "<i8")])  # wfb_l1_locus_t
