"""GPU parity tests (-m gpu): every call goes through the C ABI of libwfmash_b200.so (ctypes) and is
compared bit-exactly with the committed golden fixtures and with the oracle on the same seeded inputs."""
import ctypes
import os

import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def wb():
    import wfmash_b200 as w
    assert w.device_count() >= 1, "GPU tests need a CUDA device"
    return w


def _check_batch_against_oracle(wb, oracle, pairs, pen):
    al = wb.Aligner(0, penalties=pen)
    res = al.align_end2end_batch(pairs)
    al.close()
    for (p, t), r in zip(pairs, res):
        st, ops, score = util.orc_biwfa(oracle, p, t, pen)
        assert r.status == st == 0
        assert r.ops == ops, (len(p), len(t), util.rle(r.ops)[:80], util.rle(ops)[:80])
        assert r.score == score
    return res


def test_biwfa_reference_golden_vectors_bit_exact(wb):
    # the reference's own known-answer file (penalties 0,4,6,2,24,1): score \t CIGAR per pair
    pairs = util.golden_pairs()
    gold = util.golden_alg("wfa_utest.biwfa.affine2p.alg.gz")
    al = wb.Aligner(0, penalties=util.GOLDEN_PEN)
    res = al.align_end2end_batch(pairs)
    for r, (gscore, gcigar) in zip(res, gold):
        assert r.status == 0
        assert util.rle(r.ops) == gcigar
        assert r.score == int(gscore)


def test_wavefront_align_shim_matches_reference_golden_vectors(wb):
    # the wavefront_align-shaped single-pair entry reproduces the same known-answer file
    pairs = util.golden_pairs()[:40]
    gold = util.golden_alg("wfa_utest.biwfa.affine2p.alg.gz")[:40]
    al = wb.Aligner(0, penalties=util.GOLDEN_PEN)
    for (p, t), (gscore, gcigar) in zip(pairs, gold):
        st, ops, score = al.wavefront_align(p, t)
        assert st == 0 and util.rle(ops) == gcigar and score == int(gscore)


def test_biwfa_wfmash_penalties_match_reference_fixture(wb):
    pairs = util.golden_pairs()
    rows = util.golden_alg("wfa_wfmash_pen.tsv.gz")
    al = wb.Aligner(0, penalties=util.WFMASH_PEN)
    res = al.align_end2end_batch(pairs)
    for r, row in zip(res, rows):
        assert r.status == 0 and r.ops.decode() == row[0]


def test_biwfa_edge_cases_match_oracle(wb, oracle):
    # empty sides, <=100 bp (base-case-only path), ragged lengths, unrelated and low-complexity pairs
    pairs = util.random_pairs(250, seed=11)
    pairs += [(b"A", b"A"), (b"A", b"C"), (b"", b"ACGT"), (b"ACGT", b""), (b"ACGT" * 30, b"ACGT" * 30),
              (b"A" * 101, b"A" * 101), (b"A" * 500, b"T" * 480), (b"ACGTN" * 60, b"ACGTN" * 61)]
    _check_batch_against_oracle(wb, oracle, pairs, util.WFMASH_PEN)


def test_biwfa_medium_pairs_match_oracle(wb, oracle):
    from wfmash_b200 import synth
    recs = synth.mapping_records(24, seed=5, len_lo=2000, len_hi=20000, divergences=[0.001, 0.01, 0.05, 0.1, 0.2], pad=0)
    recs += synth.mapping_records(8, seed=6, len_lo=3000, len_hi=9000, divergences=[0.02, 0.1], pad=1000)
    _check_batch_against_oracle(wb, oracle, [(p, t) for p, t, _ in recs], util.WFMASH_PEN)


def test_biwfa_full_size_records(wb, oracle):
    # -P50k sized records: a few compared bit-exactly with the oracle, the rest through
    # size-independent properties (valid transcript, consumes both sequences, score consistent,
    # deterministic across runs / batch compositions)
    from wfmash_b200 import synth
    recs = synth.mapping_records(24, seed=9, len_lo=30000, len_hi=50000, divergences=[0.005, 0.02, 0.05], pad=1000)
    pairs = [(p, t) for p, t, _ in recs]
    al = wb.Aligner(0)
    res = al.align_end2end_batch(pairs)
    for i, ((p, t), r) in enumerate(zip(pairs, res)):
        assert r.status == 0
        assert oracle.orc_cigar_check(p, len(p), t, len(t), r.ops, len(r.ops)) == 1
        P = util.Pen(*util.WFMASH_PEN)
        assert -oracle.orc_cigar_score(r.ops, len(r.ops), ctypes.byref(P)) == r.score
        if i < 4:
            st, ops, score = util.orc_biwfa(oracle, p, t, util.WFMASH_PEN)
            assert st == 0 and ops == r.ops and score == r.score
    res2 = al.align_end2end_batch(pairs[::-1])[::-1]
    assert [r.ops for r in res] == [r.ops for r in res2]
    st = al.last_stats
    assert st.cells > 0 and st.break_tasks >= len(pairs) and st.kernel_ms > 0


def test_biwfa_high_divergence_full_length_records_match_reference_digests(wb):
    """BASELINE's C5 regime: 50 kb records at 10 % and 20 % divergence and padded 20-30 kb records at 15 % (scores 3-6 x 10^4,
    wavefronts of tens of thousands of diagonals, team mode, every fallback size) against digests of the UNMODIFIED reference's
    operation strings (tests/golden/make_highdiv_golden.py ran oracle/_ref/libwfa2ref.so in the build container). None of the
    device capacities (task queue, base-case score cap) may fire."""
    import gzip, hashlib, json
    from wfmash_b200 import synth
    gold = json.loads(gzip.open(os.path.join(util.ROOT, "tests", "golden", "highdiv_reference.json.gz")).read())
    al = wb.Aligner(0)
    for s in gold:
        recs = synth.mapping_records(**s["params"])
        res = al.align_end2end_batch([(p, t) for p, t, _ in recs])
        for (p, t, _), r, g in zip(recs, res, s["records"]):
            assert (len(p), len(t)) == (g["plen"], g["tlen"])
            assert r.status == 0
            assert r.score == g["score"] and len(r.ops) == g["ops_len"]
            assert hashlib.sha256(r.ops).hexdigest() == g["sha"]
    al.close()


def test_biwfa_work_counts_equal_the_reference_algorithms(wb, oracle):
    """The roofline's numerator: the cells (C), extended matches (E) and score steps the kernels count must be the ones the
    reference's algorithm performs (wavefront_compute_affine2p.c:355-361, wavefront_extend_kernels.c:68-92,
    wavefront_bialign.c:1010-1079), counted by the oracle's instrumented restatement on the same records — over-computing would
    inflate roofline.frac. Overlap tests (O) are the exception by design: block maxima of the rows let the scan skip diagonal
    blocks that cannot hold a breakpoint, so the device examines FEWER diagonals than wavefront_bialign_overlap (:877-955) does."""
    from wfmash_b200 import synth
    recs = synth.mapping_records(12, seed=77, len_lo=2000, len_hi=16000, divergences=[0.01, 0.03, 0.08, 0.15], pad=0)
    recs += synth.mapping_records(4, seed=78, len_lo=3000, len_hi=6000, divergences=[0.05], pad=1000)
    pairs = [(p, t) for p, t, _ in recs]
    al = wb.Aligner(0)
    res = al.align_end2end_batch(pairs)
    st = al.last_stats
    al.close()
    C = util.Counters()
    for (p, t), r in zip(pairs, res):
        s_, ops, score = util.orc_biwfa(oracle, p, t, util.WFMASH_PEN, C)
        assert s_ == 0 and ops == r.ops
    assert st.cells + st.base_cells == C.cells, (st.cells, st.base_cells, C.cells)
    # (a speculative reverse step that a TEAM ran and phase 1 then dropped keeps its matches counted: never fewer, at most a trace more)
    E = st.extend_matches + st.base_extend_matches
    assert C.extend_matches <= E <= C.extend_matches + C.extend_matches // 50, (st.extend_matches, st.base_extend_matches, C.extend_matches)
    assert st.score_steps + st.base_score_steps == C.score_steps, (st.score_steps, st.base_score_steps, C.score_steps)
    assert 0 < st.overlap_tests <= C.overlap_tests, (st.overlap_tests, C.overlap_tests)


def test_biwfa_device_resident_entry_point_matches_host_entry_point(wb):
    from wfmash_b200 import synth
    recs = synth.mapping_records(16, seed=21, len_lo=500, len_hi=6000, divergences=[0.01, 0.08])
    pairs = [(p, t) for p, t, _ in recs]
    al = wb.Aligner(0)
    host = al.align_end2end_batch(pairs)
    L = wb.lib()
    blob = b"".join(p + t for p, t in pairs)
    d = L.wfb_device_malloc(0, len(blob) + 64)
    assert d
    assert L.wfb_memcpy_h2d(0, d, blob, len(blob)) == 0
    n = len(pairs)
    poff = np.zeros(n, dtype=np.int64); toff = np.zeros(n, dtype=np.int64)
    plen = np.zeros(n, dtype=np.int32); tlen = np.zeros(n, dtype=np.int32)
    o = 0
    for i, (p, t) in enumerate(pairs):
        poff[i] = o; plen[i] = len(p); o += len(p)
        toff[i] = o; tlen[i] = len(t); o += len(t)
    cap = int(plen.sum() + tlen.sum()) + 16
    ops = ctypes.create_string_buffer(cap)
    res = (wb._Res * n)()
    rc = L.wfb_align_batch_device(al._h, d, poff.ctypes.data, plen.ctypes.data, toff.ctypes.data, tlen.ctypes.data, n,
                                  ops, cap, res, None)
    assert rc == 0, L.wfb_last_error()
    for h, r in zip(host, res):
        assert r.status == 0 and ops.raw[r.ops_offset: r.ops_offset + r.ops_len] == h.ops
    L.wfb_device_free(0, d)


def test_sketch_matches_oracle(wb, oracle):
    import random
    rng = random.Random(5)

    def clean(s):
        return "".join(c.upper() if c.upper() in "ACGT" else "N" for c in s)

    for it in range(30):
        k = rng.choice([15, 15, 21, 11, 16, 17, 31, 32, 8])
        s = rng.choice([29, 39, 59, 5, 300])
        alpha = rng.choice(["ACGT", "ACGT", "acgtACGT", "ACGTN", "AC", "ACGTRYn"])
        total = rng.choice([5000, 20000])
        seq = "".join(rng.choice(alpha) for _ in range(total))
        if rng.random() < 0.3:
            u = seq[: rng.randrange(20, 300)]
            seq = (u * (total // len(u) + 1))[:total]
        frs = []
        for j in range(40):
            L = rng.choice([1000, 1000, 500, k, k - 1, k + 1, 3000, 250])
            frs.append((rng.randrange(0, total - L), L, j))
        mm, cnt, _ = wb.sketch_fragments(seq.encode(), frs, k, s)
        for j, (off, L, sid) in enumerate(frs):
            frag = clean(seq[off: off + L]).encode()
            b = (util.MM * max(s, 1))()
            nb = oracle.orc_sketch_fragment(frag, L, k, s, sid, b)
            assert nb == cnt[j]
            got = [(int(x["hash"]), int(x["wpos"]), int(x["wpos_end"]), int(x["seqId"]), int(x["strand"])) for x in mm[j, :nb]]
            exp = [(x.hash, x.wpos, x.wpos_end, x.seqId, x.strand) for x in b[:nb]]
            assert got == exp


def test_sketch_full_size_properties(wb, oracle):
    # C3-sized fragment count (scerevisiae8: 96.3 k fragments of w = 1000 at s = 29): sortedness,
    # distinctness and bounds for all, bit-exact spot checks against the oracle
    from wfmash_b200 import synth
    rng = np.random.default_rng(4)
    nfr, w, k, s = 96000, 1000, 15, 29
    seq = synth.random_seq(nfr * w, rng).tobytes()
    frs = np.zeros(nfr, dtype=wb.FRAG_DTYPE)
    frs["seq_offset"] = np.arange(nfr, dtype=np.int64) * w
    frs["len"] = w
    frs["seq_id"] = np.arange(nfr) % 136
    mm, cnt, ms = wb.sketch_fragments(seq, frs, k, s)
    assert (cnt == s).all()
    h = mm["hash"]
    assert (h[:, 1:] > h[:, :-1]).all()
    assert (mm["wpos"] >= 0).all() and (mm["wpos_end"] <= w - k).all() and (mm["wpos"] <= mm["wpos_end"]).all()
    for j in (0, 1, nfr // 2, nfr - 1):
        b = (util.MM * s)()
        nb = oracle.orc_sketch_fragment(seq[j * w: (j + 1) * w], w, k, s, int(frs["seq_id"][j]), b)
        got = [(int(x["hash"]), int(x["wpos"]), int(x["wpos_end"]), int(x["seqId"]), int(x["strand"])) for x in mm[j, :nb]]
        assert got == [(x.hash, x.wpos, x.wpos_end, x.seqId, x.strand) for x in b[:nb]]


def _orc_add_minmers(oracle, seq, k, w, s, sid):
    import re
    oracle.orc_add_minmers.restype = ctypes.c_int64
    cl = re.sub(rb"[^ACGT]", b"N", seq.upper())
    cap = len(cl) // 2 + 1000
    out = np.zeros(cap, dtype=np.dtype([("hash", "<u8"), ("wpos", "<i8"), ("wpos_end", "<i8"), ("seqId", "<i4"), ("strand", "<i2"), ("pad_", "<i2")]))
    n = oracle.orc_add_minmers(cl, ctypes.c_int64(len(cl)), k, w, s, sid, ctypes.c_void_p(out.ctypes.data), ctypes.c_int64(cap))
    assert n <= cap
    return out[:n]


def test_reference_minmers_match_oracle(wb, oracle):
    # addMinmers parity incl. order: random, tandem repeats (stale-heap / tally-split stress), N runs,
    # lower case, sequences shorter than w (skipped), several (k, w, s)
    import random
    from wfmash_b200 import synth
    rng = np.random.default_rng(12)
    r1 = synth.random_seq(60000, rng).tobytes()
    unit = synth.random_seq(317, rng).tobytes()
    rep = synth.mutate(np.frombuffer(unit * 90, dtype=np.uint8), 0.02, rng).tobytes()
    nrun = b"N" * 3000 + r1[:6000] + b"N" * 50 + r1[7000:12000] + b"NNACGTNN" * 40
    lowc = bytes(random.Random(1).choice(b"AC") for _ in range(9000))
    cases = [([r1, rep, b"ACGT" * 100, nrun, r1.lower()[:20000], lowc], 15, 1000, 29),
             ([rep, r1[:30000]], 15, 1000, 59), ([r1[:40000], rep], 19, 500, 17), ([r1[:5000]], 11, 200, 5)]
    for seqs, k, w, s in cases:
        ids = [7 + 3 * i for i in range(len(seqs))]
        got, st = wb.minmers_build(seqs, ids, k, w, s)
        exp = [_orc_add_minmers(oracle, sq, k, w, s, sid) for sq, sid in zip(seqs, ids) if len(sq) >= w]
        exp = np.concatenate(exp)
        assert st.stitch_miss == 0
        assert len(got) == len(exp)
        for f in ("hash", "wpos", "wpos_end", "seqId", "strand"):
            assert (got[f] == exp[f]).all(), f


def test_reference_minmers_full_size_properties(wb, oracle):
    # C3-sized input (8 x 12 Mbp): global properties + bit-exact spot check of two sequences
    from wfmash_b200 import synth
    rng = np.random.default_rng(3)
    root = synth.random_seq(12_000_000, rng)
    seqs = [root.tobytes()] + [synth.mutate(root, 0.03, rng).tobytes() for _ in range(7)]
    k, w, s = 15, 1000, 29
    got, st = wb.minmers_build(seqs, list(range(8)), k, w, s)
    assert st.stitch_miss == 0 and st.bases == sum(len(x) for x in seqs)
    assert (got["wpos"] >= 0).all() and (got["wpos_end"] > got["wpos"]).all() and (got["wpos_end"] - got["wpos"] <= w).all()
    key = got["seqId"].astype(np.int64) * (1 << 40) + got["wpos"]
    assert (key[1:] >= key[:-1]).all()
    dens = len(got) / st.bases
    assert 0.04 < dens < 0.12  # ~0.0027*s windows per base
    for sid in (0, 7):
        sub = got[got["seqId"] == sid]
        exp = _orc_add_minmers(oracle, seqs[sid][:600_000], k, w, s, sid)
        a = sub[sub["wpos_end"] < 590_000]  # away from the truncation point
        b = exp[exp["wpos_end"] < 590_000]
        assert len(a) == len(b)
        for f in ("hash", "wpos", "wpos_end", "strand"):
            assert (a[f] == b[f]).all(), f


def _oracle_index(oracle, seqs, ids, k, w, s, F, threads):
    oracle.orc_index_build.restype = ctypes.c_int64
    mdt = np.dtype([("hash", "<u8"), ("wpos", "<i8"), ("wpos_end", "<i8"), ("seqId", "<i4"), ("strand", "<i2"), ("pad_", "<i2")])
    ipdt = np.dtype([("pos", "<i8"), ("hash", "<u8"), ("seqId", "<i4"), ("side", "i1"), ("pad", "i1", (3,))])
    mi = np.concatenate([_orc_add_minmers(oracle, sq, k, w, s, sid) for sq, sid in zip(seqs, ids) if len(sq) >= w]).astype(mdt)
    n = len(mi)
    valid = [sid for sq, sid in zip(seqs, ids) if len(sq) >= w]
    chunk = -(-len(valid) // threads)
    part = np.zeros(max(ids) + 1, dtype=np.int32)
    for j, sid in enumerate(valid):
        part[sid] = j // chunk
    kept = np.zeros(n, dtype=mdt); pts = np.zeros(2 * n + 2, dtype=ipdt)
    uh = np.zeros(n, dtype=np.uint64); us = np.zeros(n, dtype=np.int64); uc = np.zeros(n, dtype=np.int64)
    npnt, nu, thr = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_uint64()
    vp = lambda a: ctypes.c_void_p(a.ctypes.data)
    nk = oracle.orc_index_build(vp(mi), ctypes.c_int64(n), vp(part), ctypes.c_double(F), vp(kept), vp(pts), ctypes.byref(npnt),
                                vp(uh), vp(us), vp(uc), ctypes.byref(nu), ctypes.byref(thr))
    return kept[:nk], pts[: npnt.value], uh[: nu.value], us[: nu.value], uc[: nu.value], thr.value


def _index_case(seed):
    from wfmash_b200 import synth
    rng = np.random.default_rng(seed)
    root = synth.random_seq(150_000, rng)
    unit = synth.random_seq(400, rng)
    rep = synth.mutate(np.tile(unit, 150), 0.03, rng)
    seqs = [root.tobytes(), synth.mutate(root, 0.02, rng).tobytes(), synth.mutate(root, 0.08, rng).tobytes(), rep.tobytes(),
            (root[:30000].tobytes() + rep[:20000].tobytes()), synth.mutate(root, 0.15, rng).tobytes()[:90000], b"ACGT" * 50]
    ids = list(range(len(seqs)))
    groups = [0, 0, 1, 2, 2, 3, 4]
    return seqs, ids, groups


def test_index_build_matches_oracle(wb, oracle):
    seqs, ids, groups = _index_case(31)
    for (k, w, s, F, thr) in [(15, 1000, 29, 0.0002, 4), (15, 1000, 59, 0.0002, 1), (19, 500, 17, 0.01, 2)]:
        ix = wb.Index(seqs, ids, k, w, s, max_kmer_freq=F, index_threads=thr)
        mi, uh, us, uc, pts = ix.export()
        e_kept, e_pts, e_uh, e_us, e_uc, e_thr = _oracle_index(oracle, seqs, ids, k, w, s, F, thr)
        assert ix.stats.count_threshold == e_thr
        assert len(mi) == len(e_kept)
        for f in ("hash", "wpos", "wpos_end", "seqId", "strand"):
            assert (mi[f] == e_kept[f]).all(), f
        assert (uh == e_uh).all() and (us == e_us).all() and (uc == e_uc).all()
        e_packed = (e_pts["seqId"].astype(np.uint64) << np.uint64(41)) | (e_pts["pos"].astype(np.uint64) << np.uint64(1)) | (e_pts["side"] == 1).astype(np.uint64)
        assert (pts == e_packed).all()
        ix.close()


def test_l1_hit_counts_match_oracle(wb, oracle):
    # L1 loci incl. intersectionSize (the L1 hit count) for every fragment, several filter modes
    seqs, ids, groups = _index_case(32)
    k, w, s = 15, 1000, 29
    ix = wb.Index(seqs, ids, k, w, s, index_threads=3)
    e_kept, e_pts, e_uh, e_us, e_uc, _ = _oracle_index(oracle, seqs, ids, k, w, s, 0.0002, 3)
    cut = np.array([max(1, int(i * 0.5)) for i in range(1001)], dtype=np.int32)
    grp = np.array(groups, dtype=np.int32)
    blob = b"".join(seqs)
    offs = np.cumsum([0] + [len(x) for x in seqs])
    frags, fqs = [], []
    for qi, sq in enumerate(seqs):
        for j in range(len(sq) // w):
            frags.append((int(offs[qi]) + j * w, w, ids[qi]))
            fqs.append((ids[qi], groups[qi]))
        if len(sq) >= w and len(sq) % w:
            frags.append((int(offs[qi]) + len(sq) - w, w, ids[qi]))  # the overlapping tail fragment (computeMap.hpp:602-631)
            fqs.append((ids[qi], groups[qi]))
    l1dt = np.dtype([("seqId", "<i4"), ("pad", "<i4"), ("start", "<i8"), ("end", "<i8"), ("isz", "<i4"), ("pad2", "<i4")])
    vp = lambda a: ctypes.c_void_p(a.ctypes.data)
    nloci = 0
    for (ss, sp, lt, mh) in [(1, 1, 0, 3), (0, 0, 0, 2), (0, 1, 1, 5), (0, 0, 1, 12)]:
        r = ix.l1(blob, frags, fqs, mh, cut, grp, skip_self=ss, skip_prefix=sp, lower_triangular=lt)
        assert (r["status"] == 0).all()
        for f, (fr, fq) in enumerate(zip(frags, fqs)):
            qn = int(r["q_count"][f])
            qh = np.ascontiguousarray(r["q_minmers"]["hash"][f, :qn])
            o = np.zeros(512, dtype=l1dt)
            n2 = oracle.orc_l1_fragment(vp(e_uh), vp(e_us), vp(e_uc), ctypes.c_int64(len(e_uh)), vp(e_pts), vp(qh), qn, fq[0], fq[1],
                                        vp(grp), ss, sp, lt, mh, s, w, vp(cut), len(cut), vp(o), 512)
            got = r["loci"][int(r["offset"][f]): int(r["offset"][f]) + int(r["count"][f])]
            assert n2 == len(got), (f, n2, len(got))
            assert (got["seqId"] == o["seqId"][:n2]).all() and (got["rangeStartPos"] == o["start"][:n2]).all()
            assert (got["rangeEndPos"] == o["end"][:n2]).all() and (got["intersectionSize"] == o["isz"][:n2]).all()
            nloci += n2
    assert nloci > 1000
    ix.close()


def test_l2_mappings_match_oracle(wb, oracle):
    # L1 + L2 fused on the device (wfb_map_fragments_batch) against the oracle's mapSingleQueryFrag restatement
    # (oracle L2 pinned to the reference's slidingMap.hpp / mappingCore.hpp by tests/test_l2_emu_cpu.py): every fragment
    # mapping (refSeqId, refStartPos, optimalStart/End, conservedSketches, strand, nucIdentity bits), both strands,
    # with and without the stage-1 / identity filters, several (k, w, s).
    from tests import maputil
    from tests.test_l2_emu_cpu import L2_CASES, same_mappings
    total = 0
    for seed, mode, (k, w, s) in L2_CASES:
        seqs, ids, groups = maputil.l2_case(seed=seed)
        index = maputil.oracle_index(oracle, seqs, ids, k, w, s, 0.0002, 3)
        ix = wb.Index(seqs, ids, k, w, s, index_threads=3)
        kept = ix.export()[0]
        assert len(kept) == len(index[0]) and (kept["hash"] == index[0]["hash"]).all() and (kept["wpos"] == index[0]["wpos"]).all()
        cut = np.array([max(1, int(i * 0.5)) for i in range(1001)], dtype=np.int32)
        grp = np.array(groups, dtype=np.int32)
        blob = b"".join(seqs)
        offs = np.cumsum([0] + [len(x) for x in seqs])
        frs = maputil.fragments_of(seqs, w)
        frags = [(int(offs[qi]) + st, w, ids[qi]) for qi, st in frs]
        fqs = [(ids[qi], groups[qi]) for qi, _ in frs]
        ss, sp, lt, mh = mode
        s1 = wb.stage1_min_hits(k, s)
        assert (s1 == maputil.stage1_table(oracle, 1.0, 0.0, k, s)).all()
        ms = wb.l2_min_shared(0.85, k, s)
        for stage1 in (True, False):
            _, q_all, q_count, loci, want = maputil.oracle_map_fragments(oracle, index, seqs, ids, groups, k, w, s, mode, stage1=stage1,
                                                                         min_shared=None if stage1 else ms)
            r = ix.map_fragments(blob, frags, fqs, mh, cut, grp, skip_self=ss, skip_prefix=sp, lower_triangular=lt,
                                 stage1_min_hits=s1 if stage1 else None, l2_min_shared=None if stage1 else ms, with_l1=True)
            assert (r["status"] == 0).all()
            assert (r["l1"]["q_count"] == q_count).all()
            assert int(r["n_l1_loci"]) == len(loci)
            got = r["mappings"]
            assert same_mappings(got, want), (seed, stage1, len(got), len(want))
            assert (got["kmerComplexity"] == want["kmerComplexity"]).all()
            off = r["offset"]
            assert off[0] == 0 and off[-1] == len(got) and (np.diff(off) == np.bincount(got["frag"], minlength=len(frags))).all()
            assert r["l2_steps"] > r["l2_loci"] > 0
            total += len(got)
        # the fast path (no L1 dump) gives the same mappings
        r2 = ix.map_fragments(blob, frags, fqs, mh, cut, grp, skip_self=ss, skip_prefix=sp, lower_triangular=lt, l2_min_shared=ms)
        assert same_mappings(r2["mappings"], want)
        ix.close()
    assert total > 2000


def _patch_cases(seed, n):
    import random
    rng = random.Random(seed)
    cases = []
    for _ in range(n):
        L = rng.choice([5, 20, 130, 200, 500, 1500, 4000])
        x = "".join(rng.choice("ACGT") for _ in range(L))
        y = util.mutate(x, rng.choice([0, 0.02, 0.1, 0.3]), rng)
        if rng.random() < 0.3:
            y = "".join(rng.choice("ACGT") for _ in range(rng.randrange(1, 60))) + y
        if rng.random() < 0.3:
            x = x + "".join(rng.choice("ACGT") for _ in range(rng.randrange(1, 60)))
        if not x or not y:
            continue
        p, tt = x.encode(), y.encode()
        cases.append((p, len(p), 0, tt, len(tt), 0))   # head patch: begin-free (wflign.cpp:300-305)
        cases.append((p, 0, len(p), tt, 0, len(tt)))   # tail patch: end-free  (wflign.cpp:392-397)
    return cases


def test_ends_free_patch_alignments_match_oracle_and_reference(wb, oracle):
    # a13: ends-free WFA of the head/tail patches; the oracle is pinned to the AVX2-built reference (MemoryMed)
    cases = _patch_cases(77, 200)
    al = wb.Aligner(0)
    ref = util.load_ref("libwfa2ref.so")
    for G in (8, 1, 16):
        res = al.align_ends_free_batch(cases, term_group=G)
        for (p, pbf, pef, t, tbf, tef), r in zip(cases, res):
            buf = ctypes.create_string_buffer(2 * (len(p) + len(t)) + 16)
            n, sc = ctypes.c_int(), ctypes.c_int()
            P = util.Pen(*util.WFMASH_PEN)
            st = oracle.orc_wfa_endsfree(p, len(p), pbf, pef, t, len(t), tbf, tef, ctypes.byref(P), G, buf, len(buf), ctypes.byref(n), ctypes.byref(sc))
            assert st == r.status == 0
            assert buf.raw[: n.value] == r.ops
            if G == 8 and ref is not None:   # oracle/_ref is compiled with -march=x86-64-v3 (AVX2 kernels)
                st2 = ref.ref_wfa_endsfree(p, len(p), pbf, pef, t, len(t), tbf, tef, *util.WFMASH_PEN, 1, buf, len(buf), ctypes.byref(n), ctypes.byref(sc))
                assert st2 == 0 and buf.raw[: n.value] == r.ops


def test_do_biwfa_alignment_paf_lines_match_reference(wb):
    # a14: whole-record parity. The fixture holds the PAF text written by the unmodified reference do_biwfa_alignment
    # (AVX2 build -> term_group 8) for tests.util.paf_records(); main biWFA, head / tail patches, swizzles, trimming,
    # identity metrics and the filters all have to agree for the lines to be byte-identical.
    recs = util.paf_records()
    gold = util.paf_golden()
    al = wb.Aligner(0, penalties=tuple(gold["penalties"]))
    R = util.load_wflign_ref()
    for kw, lines_ref in zip(gold["filter_sets"], gold["lines"]):
        lines, status = al.biwfa_paf_batch(recs, term_group=gold["term_group"], **kw)
        assert wb.REC_PATCH_CAP not in status
        for i, (got, want) in enumerate(zip(lines, lines_ref)):
            assert got.decode() == want, (i, kw, got[:200], want[:200])
            assert (status[i] == wb.REC_WRITTEN) == bool(want)
        if R is not None:  # live differential against the compiled reference when it travelled with the snapshot
            for r, got in zip(recs[:20], lines[:20]):
                assert got == util.ref_paf(R, r, **kw)
    # the SAM branch of the same function (write_alignment_sam + MD tag, wflign_patch.cpp:2397-2609)
    import hashlib
    sgold = util.sam_golden()
    for kw, want in zip(sgold["sets"], sgold["lines"]):
        lines, status = al.biwfa_paf_batch(recs, term_group=sgold["term_group"], sam_format=True, **kw)
        for i, (got, w) in enumerate(zip(lines, want)):
            assert hashlib.sha256(got).hexdigest() == w["sha"], (i, kw, got[:120], w["head"])
        if R is not None:
            for r, got in zip(recs[40:52], lines[40:52]):
                assert got == util.ref_sam(R, r, **kw)
    al.close()


def test_ani_auto_identity_matches_reference(wb, oracle):
    # group MinHash sketches from ani_hash_kernel (threshold filter + segmented radix sort) against the oracle's
    # literal StreamingMinHash restatement, and the adopted identity against the doubles the reference's UNMODIFIED
    # estimate_identity_for_groups returned (committed fixture; live when oracle/_ref travelled with the snapshot).
    import gzip, hashlib, json, os
    from tests import aniutil
    from tests.test_ani_cpu import SETS
    from wfmash_b200 import pipeline
    with gzip.open(os.path.join(util.GOLD, "ani_reference.json.gz"), "rt") as f:
        gold = json.load(f)
    seqs = aniutil.case()
    ids = pipeline.SequenceIds(seqs, seqs, "#")
    gids = sorted(set(ids.group)); dense = {g: i for i, g in enumerate(gids)}
    grp = [dense[ids.group[ids.id_of[n]]] for n, _ in seqs]
    passes = []
    for s in (4096, 64, 16):
        sk, cnt, st = wb.ani_group_sketches([x for _, x in seqs], grp, len(gids), 21, s)
        osk, ocnt = aniutil.oracle_group_sketches(oracle, [x for _, x in seqs], grp, len(gids), 21, s)
        assert (cnt == ocnt).all() and (sk == osk).all(), s
        passes.append(st.passes)
        if s == 4096:
            assert hashlib.sha256(sk.tobytes() + cnt.tobytes()).hexdigest() == gold["sketch_sha"]
            assert st.candidates < 0.6 * st.valid_kmers and st.hash_kernel_ms > 0
    assert passes[0] == 1 and max(passes[1:]) >= 2
    R = util.load_ref("libstatsref.so")
    for name, (q0, q1), (t0, t1), pct, adj in SETS:
        P = pipeline.Params(percentage_identity=None, ani_percentile=pct, ani_adjustment=adj)
        got = pipeline.estimate_identity(seqs[t0:t1], seqs[q0:q1], P)[0]
        assert float(got).hex() == gold["identity"][name], name
        if R is not None:
            assert got == aniutil.reference_identity(R, seqs[t0:t1], seqs[q0:q1], "#", pct, adj)
    # a genome-sized group: every k-mer of 24 Mbp through the threshold filter; size-independent properties
    import numpy as np
    from wfmash_b200 import synth
    rng = np.random.default_rng(77)
    big = [synth.random_seq(6_000_000, rng).tobytes() for _ in range(4)]
    sk, cnt, st = wb.ani_group_sketches(big, [0, 0, 1, 1], 2)
    assert (cnt == 4096).all() and st.passes == 1 and st.valid_kmers > 23_900_000 and st.candidates < 200_000
    assert (np.diff(sk.astype(np.float64), axis=1) >= 0).all()
    # the reference's merge identity: a group's sketch = the 4096 smallest of its sequences' own sketches taken together
    parts = [wb.ani_group_sketches([b], [0], 1)[0][0] for b in big[:2]]
    assert (np.sort(np.concatenate(parts))[:4096] == sk[0]).all()


def test_index_import_and_file_round_trip_map_identically(wb, tmp_path):
    # wfb_index_import of an exported index, and of the same index after a trip through the reference's file
    # format, must map exactly like the index built from the sequences
    import numpy as np
    from tests import maputil
    seqs, ids, groups = maputil.l2_case(seed=41)
    k, w, s = 15, 1000, 29
    ix = wb.Index(seqs, ids, k, w, s, index_threads=2)
    blob = b"".join(seqs)
    offs = np.cumsum([0] + [len(x) for x in seqs])
    frags = np.array([(int(offs[q]) + j * w, w, ids[q]) for q in range(len(seqs)) for j in range(len(seqs[q]) // w)], dtype=wb.FRAG_DTYPE)
    fqs = np.array([(ids[q], groups[q]) for q in range(len(seqs)) for j in range(len(seqs[q]) // w)], dtype=wb.FRAG_QUERY_DTYPE)
    cut, s1 = wb.sketch_cutoffs(s, k), wb.stage1_min_hits(k, s)
    grp = np.array(groups, dtype=np.int32)

    def run(index):
        r = index.map_fragments(blob, frags, fqs, 3, cut, grp, stage1_min_hits=s1, l2_min_shared=wb.l2_min_shared_relaxed(0.85, k, s))
        return r["mappings"].tobytes(), r["offset"].tobytes()

    base = run(ix)
    exported = ix.export()
    assert len(exported[0]) > 1000
    ix2 = wb.Index.from_export(exported, k, w, s)
    assert run(ix2) == base
    path = str(tmp_path / "index.bin")
    wb.index_file_write(path, exported, k, w, s, [f"s{i}" for i in ids], {f"s{i}": i for i in ids})
    hdr, data, _ = wb.index_file_read(path)
    ix3 = wb.Index.from_export(data, hdr["kmer_size"], hdr["window_length"], hdr["sketch_size"])
    assert run(ix3) == base and all((a == b).all() for a, b in zip(ix3.export()[1:], exported[1:]))
    for i in (ix, ix2, ix3):
        i.close()



def test_pipeline_map_align_matches_reference_pieces(wb, oracle):
    # SURVEY 8 f2 + b3 + both hot paths chained: sequences -> index -> L1/L2 kernels -> host chain merge + filters -> mapping
    # PAF -> padded records -> biWFA kernels + patches -> alignment PAF, against the run composed from the reference-side
    # pieces (tests/pipeutil.py): committed fixture always, live when oracle/_ref travelled with the snapshot.
    import gzip, json, os
    from tests import pipeutil
    from wfmash_b200 import pipeline
    with gzip.open(os.path.join(util.GOLD, "pipeline_reference.json.gz"), "rt") as f:
        gold = json.load(f)
    fref, wref = util.load_ref("libfilterref.so"), util.load_wflign_ref()
    assert [c["name"] for c in gold["cases"]] == [c[0] for c in pipeutil.PIPELINE_CASES]
    for (name, gen, prm), g in zip(pipeutil.PIPELINE_CASES, gold["cases"]):
        seqs = pipeutil.case(**gen)
        P = pipeutil.params(prm)
        paf, st = pipeline.wfmash(seqs, seqs, P)
        assert sorted(st["mapping_paf"].decode().splitlines()) == sorted(g["mapping_paf"].splitlines()), name
        # the alignment lines follow the order of the mapping PAF (grouped by query, sorted by query start with the
        # reference's unstable sort): compare as multisets of (first 12 columns, digest of the whole line)
        got = sorted((d["head"], d["sha"]) for d in map(pipeutil.line_digest, [ln + b"\n" for ln in paf.split(b"\n") if ln]))
        want = sorted((d["head"], d["sha"]) for d in g["lines"] if d["head"])
        assert got == want, name
        assert st["written"] == len(want) and st["aligned_bp"] > 0
        # the same run through the two one-call C-ABI phases (wfmash_b200/csrc/phases_host.cu)
        mp_c, mst = wb.map_phase(seqs, seqs, wb.MapPhaseParams(filter=P.filter, window_length=P.window_length, percentage_identity=P.percentage_identity))
        assert mp_c == st["mapping_paf"], name
        al = wb.Aligner(0)
        paf_c, ast = wb.align_phase(al, mp_c, seqs, seqs, window_length=P.window_length, disable_chain_patching=P.disable_chain_patching)
        al.close()
        assert paf_c == paf and ast.records == st["records"] and ast.aligned_bp == st["aligned_bp"], name
        if fref is not None and wref is not None and name == "defaults_p90":
            mp, lines = pipeutil.expected(seqs, P, oracle, fref, wref)
            assert st["mapping_paf"] == mp and paf == b"".join(lines)
            A = util.load_ref("libalignref.so")   # the reference's whole alignment phase (align::Aligner::compute) on our mapping PAF
            if A is not None:
                assert pipeutil.reference_align_phase(A, mp_c, seqs, P) == paf_c


def test_pipeline_auto_identity_then_map(wb):
    # the CLI default (-p ani50-2): estimate, adopt, re-derive the sketch size, map with it
    from tests import pipeutil
    from wfmash_b200 import pipeline
    seqs = pipeutil.case(seed=5, length=60_000)
    P = pipeline.auto_identity(seqs, seqs, pipeline.Params(percentage_identity=None))
    assert 0.85 < P.percentage_identity < 0.97 and P.sketch_size == wb.sketch_size(P.percentage_identity, 1000, 15)
    m = pipeline.map(seqs, seqs, P)
    assert m.stats["mappings"] >= 6 and m.paf.count(b"\n") == m.stats["mappings"]
    mp_c, mst = wb.map_phase(seqs, seqs, wb.MapPhaseParams())   # percentage_identity <= 0: the C phase estimates it itself
    assert abs(mst.percentage_identity - P.percentage_identity) < 1e-6 and mst.sketch_size == P.sketch_size and mp_c == m.paf