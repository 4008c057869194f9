"""GPU parity tests (-m gpu): every call goes through the C ABI of libwfmash_b200.so (ctypes) and is
compared bit-exactly with the committed golden fixtures and with the oracle on the same seeded inputs."""
import ctypes

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
        m = exp["wpos_end"] < 590_000  # away from the truncation point
        assert (sub[: m.sum()][["hash", "wpos", "wpos_end", "strand"]] == exp[m][["hash", "wpos", "wpos_end", "strand"]]).all()
