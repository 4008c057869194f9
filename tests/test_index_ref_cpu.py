"""SURVEY 8 row a4 pinned by the real code: the oracle's index build (orc_index_build, which the GPU index is compared
with in tests/test_gpu_parity.py::test_index_build_matches_oracle) against the reference's UNMODIFIED skch::Sketch
(src/map/include/winSketch.hpp:175-457: thread-pool sketching, hash frequencies, frequency cut-off with its safety
re-threshold, minmerPosLookupIndex / minmerIndex) and skch::SequenceIdManager (sequenceIds.hpp) compiled in place behind
oracle/ref_sketch_driver.cpp (htslib replaced by oracle/shims/htslib/faidx.h) — live when oracle/_ref is present, and by
the committed fixture tests/golden/index_reference.json.gz generated from it."""
import ctypes
import gzip
import hashlib
import json
import os
import tempfile

import numpy as np
import pytest

from tests import maputil, util

PT = np.dtype([("hash", "<u8"), ("pos", "<i8"), ("seqId", "<i4"), ("side", "<i4")])
# (seed, k, w, s, max_kmer_freq, threads)
CASES = [(32, 15, 1000, 29, 0.0002, 3), (33, 15, 1000, 59, 0.0002, 1), (34, 19, 500, 17, 0.01, 2), (35, 15, 256, 11, 0.0002, 4), (36, 15, 1000, 39, 5.0, 2)]


def names_of(groups):
    return [f"g{g}#1#s{i}" for i, g in enumerate(groups)]


def reference_index(R, seqs, names, k, w, s, F, threads):
    R.ref_sketch_build.restype = ctypes.c_void_p
    n = len(seqs)
    with tempfile.TemporaryDirectory() as d:
        h = ctypes.c_void_p(R.ref_sketch_build(os.path.join(d, "t.fa").encode(), (ctypes.c_char_p * n)(*[x.encode() for x in names]), (ctypes.c_char_p * n)(*seqs),
                                               (ctypes.c_int64 * n)(*[len(x) for x in seqs]), n, k, ctypes.c_int64(w), s, threads, ctypes.c_double(F), b"#"))
        assert h.value
        nm, nh, npt = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        R.ref_sketch_sizes(h, ctypes.byref(nm), ctypes.byref(nh), ctypes.byref(npt))
        mi = np.zeros(nm.value, dtype=maputil.MDT); hs = np.zeros(nh.value, dtype=np.uint64)
        st = np.zeros(nh.value + 1, dtype=np.int64); pts = np.zeros(npt.value, dtype=PT)
        R.ref_sketch_export(h, maputil.vp(mi), maputil.vp(hs), maputil.vp(st), maputil.vp(pts))
        groups = [R.ref_sketch_group(h, i) for i in range(n)]
        R.ref_sketch_free(h)
    return mi, hs, st, pts, groups


def oracle_flat(oracle, seqs, ids, k, w, s, F, threads):
    """The oracle index in the driver's flat layout: minmers, ascending hashes, CSR starts, points (hash, pos, seqId, side)."""
    kept, opts, uh, us, uc, _ = maputil.oracle_index(oracle, [maputil.clean(x) for x in seqs], ids, k, w, s, F, threads)
    order = np.argsort(uh, kind="stable")
    pts = np.zeros(len(opts), dtype=PT)
    st = np.zeros(len(uh) + 1, dtype=np.int64)
    o = 0
    for j, i in enumerate(order):
        seg = opts[int(us[i]): int(us[i]) + int(uc[i])]
        pts["hash"][o: o + len(seg)] = seg["hash"]; pts["pos"][o: o + len(seg)] = seg["pos"]
        pts["seqId"][o: o + len(seg)] = seg["seqId"]; pts["side"][o: o + len(seg)] = seg["side"]
        st[j] = o
        o += len(seg)
    st[len(uh)] = o
    return kept, uh[order], st, pts[:o]


def canonical(mi, hs, st, pts):
    """The reference's thread pool hands the per-sequence minmer lists back in COMPLETION order (winSketch.hpp:222-240), so
    with more than one thread the order of whole sequences inside minmerIndex / inside one hash's point list is not
    reproducible run to run. Inside a sequence the minmers are sorted by (wpos, wpos_end) with an UNSTABLE std::sort
    (commonFunc.hpp:696), so the order of equal (wpos, wpos_end) is unspecified as well. Compare with sequences put back
    in id order and those ties ordered by hash; nothing else is reordered (hashes differ inside a tie, so the per-hash
    point lists do not depend on it)."""
    mi = mi[np.lexsort((mi["strand"], mi["hash"], mi["wpos_end"], mi["wpos"], mi["seqId"]))]
    grp = np.repeat(np.arange(len(hs)), np.diff(st))
    pts = pts[np.lexsort((pts["seqId"], grp))]  # lexsort is stable: hash group, then sequence, original order inside
    return mi, hs, st, pts


def digest(mi, hs, st, pts):
    mi, hs, st, pts = canonical(mi, hs, st, pts)
    h = hashlib.sha256()
    for f in ("hash", "wpos", "wpos_end", "seqId", "strand"):
        h.update(np.ascontiguousarray(mi[f]).tobytes())
    for a in (hs, st, pts["hash"], pts["pos"], pts["seqId"], pts["side"]):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def case_seqs(seed):
    seqs, ids, groups = maputil.l2_case(seed=seed) if seed % 2 else maputil.l1_case(seed=seed)
    return seqs, ids, groups


def test_oracle_index_reproduces_reference_sketch_fixture(oracle):
    with gzip.open(os.path.join(util.GOLD, "index_reference.json.gz"), "rt") as f:
        gold = json.load(f)
    assert len(gold["cases"]) == len(CASES)
    for (seed, k, w, s, F, threads), g in zip(CASES, gold["cases"]):
        seqs, ids, groups = case_seqs(seed)
        mi, hs, st, pts = oracle_flat(oracle, seqs, ids, k, w, s, F, threads)
        assert (len(mi), len(hs), len(pts)) == (g["n_minmers"], g["n_hashes"], g["n_points"]), seed
        assert digest(mi, hs, st, pts) == g["sha"], seed


@pytest.mark.ref
def test_oracle_index_matches_compiled_reference_sketch_live(oracle):
    from wfmash_b200 import pipeline
    R = util.load_ref("libsketchref.so")
    if R is None:
        pytest.skip("oracle/_ref not built (reference sources absent)")
    for seed, k, w, s, F, threads in [(c[0] + 50,) + c[1:] for c in CASES]:
        seqs, ids, groups = case_seqs(seed)
        names = names_of(groups)
        r_mi, r_hs, r_st, r_pts, r_groups = reference_index(R, seqs, names, k, w, s, F, threads)
        r_mi, r_hs, r_st, r_pts = canonical(r_mi, r_hs, r_st, r_pts)
        mi, hs, st, pts = canonical(*oracle_flat(oracle, seqs, ids, k, w, s, F, threads))
        assert len(mi) == len(r_mi) > 1000 and all((mi[f] == r_mi[f]).all() for f in ("hash", "wpos", "wpos_end", "seqId", "strand")), seed
        assert (hs == r_hs).all() and (st == r_st).all(), seed
        assert all((pts[f] == r_pts[f]).all() for f in ("hash", "pos", "seqId", "side")), seed
        # the host mirror of SequenceIdManager::buildRefGroups (sequenceIds.hpp:284-330)
        mirror = pipeline.SequenceIds([(n, b"") for n in names], [], "#")
        assert mirror.group == r_groups


def _flat_to_export(mi, hs, st, pts):
    """The driver's flat layout -> the Index.export() layout (packed points seqId << 41 | pos << 1 | open)."""
    import wfmash_b200 as wb
    m = np.zeros(len(mi), dtype=wb.MINMER_DTYPE)
    for f in ("hash", "wpos", "wpos_end", "seqId", "strand"):
        m[f] = mi[f]
    packed = (pts["seqId"].astype(np.uint64) << np.uint64(41)) | (pts["pos"].astype(np.uint64) << np.uint64(1)) | (pts["side"] == 1).astype(np.uint64)
    return m, hs.astype(np.uint64), st[:-1].astype(np.uint32), np.diff(st).astype(np.uint32), packed


def _same_export(a, b):
    return (all((a[0][f] == b[0][f]).all() for f in ("hash", "wpos", "wpos_end", "seqId", "strand")) and len(a[0]) == len(b[0])
            and all(len(x) == len(y) and (x == y).all() for x, y in zip(a[1:], b[1:])))


@pytest.mark.ref
def test_index_file_is_interchangeable_with_the_reference(oracle, tmp_path):
    """a `-W` file written by the reference's UNMODIFIED Sketch::writeIndex loads here, and a file written here
    loads in its Sketch::readIndex (both compiled in place in libsketchref.so) — same minmerIndex and per-hash postings."""
    import wfmash_b200 as wb
    R = util.load_ref("libsketchref.so")
    if R is None:
        pytest.skip("oracle/_ref not built (reference sources absent)")
    R.ref_sketch_build.restype = ctypes.c_void_p
    R.ref_sketch_read_index.restype = ctypes.c_void_p
    seed, k, w, s, F, threads = CASES[0]
    seqs, ids, groups = case_seqs(seed)
    names = names_of(groups)
    n = len(seqs)
    cn = (ctypes.c_char_p * n)(*[x.encode() for x in names]); cs = (ctypes.c_char_p * n)(*seqs); cl = (ctypes.c_int64 * n)(*[len(x) for x in seqs])
    fa = str(tmp_path / "t.fa").encode()
    h = ctypes.c_void_p(R.ref_sketch_build(fa, cn, cs, cl, n, k, ctypes.c_int64(w), s, 1, ctypes.c_double(F), b"#"))
    ref_file = str(tmp_path / "ref.idx")
    R.ref_sketch_write_index(h, ref_file.encode())
    R.ref_sketch_free(h)
    # reference -> here
    hdr, data, nxt = wb.index_file_read(ref_file)
    assert (hdr["kmer_size"], hdr["window_length"], hdr["sketch_size"], hdr["batch_idx"], hdr["total_batches"]) == (k, w, s, 0, 1)
    assert hdr["target_names"] == names and hdr["id_map"] == {nm: i for i, nm in enumerate(names)} and hdr["next_id"] == n
    assert nxt == os.path.getsize(ref_file)
    want = _flat_to_export(*oracle_flat(oracle, seqs, ids, k, w, s, F, 1))   # = the reference's index (test above), threads = 1: no reordering
    assert _same_export(data, want)
    # here -> reference
    our_file = str(tmp_path / "ours.idx")
    wb.index_file_write(our_file, data, k, w, s, names, hdr["id_map"])
    assert os.path.getsize(our_file) == os.path.getsize(ref_file)
    h2 = ctypes.c_void_p(R.ref_sketch_read_index(fa, our_file.encode(), cn, cs, cl, n, k, ctypes.c_int64(w), s, b"#"))
    assert h2.value
    nm_, nh_, np_ = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
    R.ref_sketch_sizes(h2, ctypes.byref(nm_), ctypes.byref(nh_), ctypes.byref(np_))
    mi = np.zeros(nm_.value, dtype=maputil.MDT); hs = np.zeros(nh_.value, dtype=np.uint64); st = np.zeros(nh_.value + 1, dtype=np.int64); pts = np.zeros(np_.value, dtype=PT)
    R.ref_sketch_export(h2, maputil.vp(mi), maputil.vp(hs), maputil.vp(st), maputil.vp(pts))
    R.ref_sketch_free(h2)
    assert _same_export(_flat_to_export(mi, hs, st, pts), want)
    # two subsets in one file (-b): append, then read them back one after the other
    wb.index_file_write(our_file, data, k, w, s, names[:3], hdr["id_map"], append=True, batch_idx=1, total_batches=2)
    h1, d1, o1 = wb.index_file_read(our_file)
    h2_, d2, o2 = wb.index_file_read(our_file, o1)
    assert o2 == os.path.getsize(our_file) and (h2_["batch_idx"], h2_["total_batches"], h2_["target_names"]) == (1, 2, names[:3]) and _same_export(d1, d2)
    with pytest.raises(wb.WfbError):
        wb.index_file_read(our_file, 8)          # not a subset boundary: bad magic number
    with pytest.raises(wb.WfbError):
        wb.index_file_read(str(tmp_path / "missing.idx"))


def test_index_file_reader_rejects_truncated_and_forged_files(tmp_path):
    """wfb_index_file_read trusts nothing in the file: every truncation point and every count forged to a huge value must
    come back as an error — no crash, no multi-GB allocation, no exception through the C boundary."""
    import struct
    import wfmash_b200 as wb
    rng = np.random.default_rng(5)
    n_mi, n_u = 40, 12
    mi = np.zeros(n_mi, dtype=wb.MINMER_DTYPE)
    mi["hash"] = rng.integers(1, 2**62, n_mi, dtype=np.uint64); mi["wpos"] = np.arange(n_mi) * 7; mi["wpos_end"] = mi["wpos"] + 5
    mi["seqId"] = rng.integers(0, 3, n_mi); mi["strand"] = 1
    uhash = np.sort(rng.choice(2**61, n_u, replace=False).astype(np.uint64))
    ucount = rng.integers(1, 5, n_u).astype(np.uint32)
    ustart = np.concatenate([[0], np.cumsum(ucount)[:-1]]).astype(np.uint32)
    pts = ((rng.integers(0, 3, int(ucount.sum())).astype(np.uint64) << np.uint64(41)) | (rng.integers(0, 10**6, int(ucount.sum())).astype(np.uint64) << np.uint64(1)))
    names = ["g#1#a", "g#1#b", "h#1#c"]
    path = str(tmp_path / "ok.idx")
    wb.index_file_write(path, (mi, uhash, ustart, ucount, pts), 15, 1000, 24, names, {nm: i for i, nm in enumerate(names)})
    hdr, data, nxt = wb.index_file_read(path)
    assert hdr["target_names"] == names and len(data[0]) == n_mi and len(data[1]) == n_u and nxt == os.path.getsize(path)
    blob = open(path, "rb").read()
    # 1) every truncation
    bad = str(tmp_path / "bad.idx")
    for cut in list(range(0, 200, 3)) + list(range(200, len(blob), 37)):
        open(bad, "wb").write(blob[:cut])
        with pytest.raises(wb.WfbError):
            wb.index_file_read(bad)
    # 2) forged counts: find the 8-byte fields holding the minmer count and the hash count and blow them up
    pos_mi = blob.index(struct.pack("<q", n_mi) + mi.tobytes()[:16]) if struct.pack("<q", n_mi) + mi.tobytes()[:16] in blob else None
    forged = 0
    for off in range(0, len(blob) - 8):                      # (fields after the variable-length names are not 8-byte aligned)
        v = struct.unpack_from("<Q", blob, off)[0]
        if v in (n_mi, n_u, len(names)) or (0 < v < 6):      # counts / lengths of this small file
            for huge in (1 << 39, 1 << 62, 0xFFFFFFFFFFFFFFFF):
                open(bad, "wb").write(blob[:off] + struct.pack("<Q", huge) + blob[off + 8:])
                try:
                    wb.index_file_read(bad)                   # a forged field that happens not to be a count may still parse
                except wb.WfbError:
                    forged += 1
    assert forged >= 30          # the sequence count, name lengths, the minmer count, the hash count, postings counts
    assert pos_mi is None or pos_mi > 0
