"""Host stage between the chain merge and the aligner (second part, and b3): wfb_filter_mappings_batch,
wfb_filter_by_group, wfb_one_to_one_filter, wfb_mapping_paf_format against the reference's UNMODIFIED mappingFilter.hpp /
filter.hpp / mappingOutput.hpp compiled in place (oracle/_ref/libfilterref.so) and against the committed fixture generated
from it (tests/golden/filter_reference.json.gz); wfb_mapping_paf_parse against the rules of Aligner::parseMashmapRow.
These entry points are host C++ inside the product library: no GPU is needed to run them."""
import ctypes
import gzip
import hashlib
import json
import os

import numpy as np
import pytest

from tests import chainutil, util

# (seed, generator arguments, filter parameters, PanSN groups of the 4 target sequences or None)
CASES = [
    (21, dict(w=1000), dict(), None),                                                                   # the CLI defaults
    (22, dict(w=1000), dict(num_mappings_for_segment=1, overlap_threshold=0.5, skip_prefix=1), [0, 0, 1, 2]),
    (23, dict(w=1000), dict(num_mappings_for_segment=3, block_length=3000, scaffold_gap=0), None),
    (24, dict(w=1000), dict(merge_mappings=0, filter_mode=3), None),                                     # -M -f
    (25, dict(w=1000), dict(split=0, num_mappings_for_segment=2, scaffold_min_length=500), None),                                # -N
    (26, dict(w=1000), dict(filter_mode=3), None),
    (27, dict(w=1000), dict(drop_rand=1, num_mappings_for_segment=2, overlap_threshold=1.0), None),
    (28, dict(w=500, qlen=120_000), dict(scaffold_gap=5000, scaffold_max_deviation=3000, scaffold_min_length=5000, max_mapping_length=10000), None),
    (29, dict(w=1000), dict(sparsity_hash_threshold=2**63, filter_mode=2, num_mappings_for_scaffold=2), None),
    (30, dict(w=1000, nref=1), dict(merge_mappings=0, num_mappings_for_segment=1, scaffold_min_length=1000), None),
]
REF_LEN = np.full(4, 400_000, dtype=np.int64)
QLEN = 300_000


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def params(gen, prm):
    import wfmash_b200 as wb
    return wb.FilterParams(window_length=gen.get("w", 1000), **prm)


def ours(seed, gen, prm, groups):
    import wfmash_b200 as wb
    m, off = chainutil.batch(seed, **gen)
    qlen = np.full(len(off) - 1, gen.get("qlen", QLEN), dtype=np.int64)
    out, info, oo = wb.filter_mappings_batch(params(gen, prm), m, off, qlen, REF_LEN, ref_group=groups, host_threads=3)
    return (out, info, oo), m, off


def reference(ref, gen, prm, groups, m, off):
    import wfmash_b200 as wb
    ref.ref_filter_subset.restype = ctypes.c_int64
    P = params(gen, prm)
    g = np.ascontiguousarray(groups, dtype=np.int32) if groups is not None else None
    outs, infos, oo = [], [], [0]
    for q in range(len(off) - 1):
        a = np.ascontiguousarray(m[off[q]: off[q + 1]])
        o = np.zeros(len(a) + 4, dtype=wb.MAPPING_DTYPE)
        c = np.zeros(len(a) + 4, dtype=wb.CHAIN_INFO_DTYPE)
        n = ref.ref_filter_subset(ctypes.byref(P), ctypes.c_void_p(a.ctypes.data), ctypes.c_int64(len(a)), q, ctypes.c_int64(gen.get("qlen", QLEN)),
                                  ctypes.c_void_p(g.ctypes.data) if g is not None else None, ctypes.c_void_p(REF_LEN.ctypes.data),
                                  ctypes.c_void_p(o.ctypes.data), ctypes.c_void_p(c.ctypes.data), ctypes.c_int64(len(o)))
        assert 0 <= n <= len(o)
        outs.append(o[:n]); infos.append(c[:n]); oo.append(oo[-1] + n)
    return np.concatenate(outs), np.concatenate(infos), np.array(oo, dtype=np.int64)


def paf_ours(gen, prm, out, info, oo):
    import wfmash_b200 as wb
    P = params(gen, prm)
    names = [f"s{i}" for i in range(len(REF_LEN))]
    return b"".join(wb.mapping_paf_format(P, out[oo[q]: oo[q + 1]], info[oo[q]: oo[q + 1]], f"q{q}", gen.get("qlen", QLEN), names, REF_LEN)
                    for q in range(len(oo) - 1))


def paf_reference(ref, gen, prm, out, info, oo):
    ref.ref_report_mappings.restype = ctypes.c_int64
    P = params(gen, prm)
    txt = []
    for q in range(len(oo) - 1):
        a, c = np.ascontiguousarray(out[oo[q]: oo[q + 1]]), np.ascontiguousarray(info[oo[q]: oo[q + 1]])
        buf = ctypes.create_string_buffer(300 * len(a) + 64)
        n = ref.ref_report_mappings(ctypes.byref(P), ctypes.c_void_p(a.ctypes.data), ctypes.c_void_p(c.ctypes.data), ctypes.c_int64(len(a)), f"q{q}".encode(),
                                    ctypes.c_int64(gen.get("qlen", QLEN)), ctypes.c_void_p(REF_LEN.ctypes.data), buf, ctypes.c_int64(len(buf)))
        assert n <= len(buf)
        txt.append(buf.raw[:n])
    return b"".join(txt)


def test_filters_reproduce_reference_fixture():
    with gzip.open(os.path.join(util.GOLD, "filter_reference.json.gz"), "rt") as f:
        gold = json.load(f)
    assert len(gold["cases"]) == len(CASES)
    for (seed, gen, prm, groups), g in zip(CASES, gold["cases"]):
        (out, info, oo), m, off = ours(seed, gen, prm, groups)
        assert len(m) == g["n_in"] and len(out) == g["n_out"], (seed, len(out), g["n_out"])
        assert oo.tolist() == g["out_offset"], seed
        assert digest(out) == g["sha_out"], seed
        assert digest(info) == g["sha_chain_info"], seed
        txt = paf_ours(gen, prm, out, info, oo)
        assert hashlib.sha256(txt).hexdigest() == g["sha_paf"], seed
        assert txt.decode().splitlines()[:3] == g["paf_head"], seed


@pytest.mark.ref
def test_filters_match_compiled_reference_live():
    ref = util.load_ref("libfilterref.so")
    if ref is None:
        pytest.skip("oracle/_ref not built (reference sources absent)")
    kept = dropped = 0
    for seed, gen, prm, groups in [(s + 100, g, p, grp) for s, g, p, grp in CASES]:
        (out, info, oo), m, off = ours(seed, gen, prm, groups)
        r_out, r_info, r_oo = reference(ref, gen, prm, groups, m, off)
        assert (oo == r_oo).all(), seed
        assert out.tobytes() == r_out.tobytes() and info.tobytes() == r_info.tobytes(), seed
        assert paf_ours(gen, prm, out, info, oo) == paf_reference(ref, gen, prm, out, info, oo), seed
        kept += len(out); dropped += len(m) - len(out)
    assert kept > 300 and dropped > 300

@pytest.mark.ref
def test_filters_match_compiled_reference_on_random_option_sets():
    """Random combinations of the chain / filter options (the CASES above pin ten hand-picked ones) on random mapping sets, against the
    unmodified mappingFilter.hpp + filter.hpp + mappingOutput.hpp: surviving mappings, ChainInfo and the mapping PAF must be byte-identical.
    14 300 combinations of the same generator ran clean at the end of round 2; the suite runs 80."""
    import random
    ref = util.load_ref("libfilterref.so")
    if ref is None:
        pytest.skip("oracle/_ref not built (reference sources absent)")
    rnd = random.Random(2024)
    space = [("num_mappings_for_segment", [1, 2, 3, 5], 0.35), ("overlap_threshold", [0.0, 0.5, 0.95, 1.0], 0.35), ("skip_prefix", [0, 1], 0.35),
             ("block_length", [0, 1000, 3000, 10000], 0.35), ("scaffold_gap", [0, 5000, 100000], 0.35), ("scaffold_max_deviation", [0, 3000, 100000], 0.35),
             ("scaffold_min_length", [0, 500, 5000, 50000], 0.35), ("merge_mappings", [0, 1], 0.35), ("filter_mode", [1, 2, 3], 0.35), ("split", [0, 1], 0.35),
             ("drop_rand", [0, 1], 0.15), ("max_mapping_length", [5000, 10000, 50000, 1 << 30], 0.35),
             ("sparsity_hash_threshold", [2**63, 2**62, 2**64 - 1], 0.15), ("num_mappings_for_scaffold", [1, 2, 3], 0.2)]
    for _ in range(80):
        gen = dict(w=rnd.choice([500, 1000, 1000, 2000]))
        if rnd.random() < 0.3:
            gen["qlen"] = rnd.choice([50_000, 120_000, 300_000])
        if rnd.random() < 0.2:
            gen["nref"] = rnd.choice([1, 2, 3])
        prm = {k: rnd.choice(v) for k, v, p in space if rnd.random() < p}
        groups = rnd.choice([None, None, [0, 0, 1, 2], [0, 1, 1, 1], [0, 0, 0, 0]])
        if prm.get("skip_prefix") and groups is None:
            groups = [0, 0, 1, 2]
        seed = rnd.randrange(1 << 30)
        (out, info, oo), m, off = ours(seed, gen, prm, groups)
        r_out, r_info, r_oo = reference(ref, gen, prm, groups, m, off)
        what = (seed, gen, prm, groups)
        assert (oo == r_oo).all(), what
        assert out.tobytes() == r_out.tobytes() and info.tobytes() == r_info.tobytes(), what
        assert paf_ours(gen, prm, out, info, oo) == paf_reference(ref, gen, prm, out, info, oo), what


@pytest.mark.ref
def test_filter_by_group_both_axes_match_compiled_reference_live():
    import wfmash_b200 as wb
    ref = util.load_ref("libfilterref.so")
    if ref is None:
        pytest.skip("oracle/_ref not built (reference sources absent)")
    ref.ref_filter_by_group.restype = ctypes.c_int64
    groups = np.array([0, 1, 1, 2], dtype=np.int32)
    n_checked = 0
    for seed in range(40, 46):
        m, off = chainutil.batch(seed, nq=3)
        _, merged, _, mo = wb.chain_mappings_batch(m, off, 1000)
        for axis in (0, 1):
            for n_map, prm in [(0, dict()), (1, dict(overlap_threshold=0.3)), (-2, dict(skip_prefix=1)), (2, dict(drop_rand=1))]:
                P = wb.FilterParams(window_length=1000, **prm)
                for a in (np.ascontiguousarray(m[off[0]: off[1]]), np.ascontiguousarray(merged[mo[1]: mo[2]])):
                    mine_in, mine = wb.filter_by_group(P, a, n_map, bool(axis), REF_LEN, ref_group=groups)
                    r_in = a.copy()
                    r_out = np.zeros(len(a) + 1, dtype=wb.MAPPING_DTYPE)
                    n = ref.ref_filter_by_group(ctypes.byref(P), ctypes.c_void_p(r_in.ctypes.data), ctypes.c_int64(len(a)), n_map, axis,
                                                ctypes.c_void_p(groups.ctypes.data), ctypes.c_void_p(REF_LEN.ctypes.data), ctypes.c_void_p(r_out.ctypes.data),
                                                ctypes.c_int64(len(r_out)))
                    assert n == len(mine) and mine.tobytes() == r_out[:n].tobytes(), (seed, axis, n_map)
                    assert mine_in.tobytes() == r_in.tobytes()
                    n_checked += 1
    assert n_checked == 96


@pytest.mark.ref
def test_one_to_one_final_pass_equals_reference_pieces_composed():
    """computeMap.hpp:788-850 lives inside Map (not compilable here): the reference-axis sweep it calls is the compiled
    reference, the regrouping loops around it are restated in this test."""
    import wfmash_b200 as wb
    ref = util.load_ref("libfilterref.so")
    if ref is None:
        pytest.skip("oracle/_ref not built (reference sources absent)")
    ref.ref_filter_by_group.restype = ctypes.c_int64
    P = wb.FilterParams(window_length=1000, filter_mode=wb.FILTER_ONETOONE, num_mappings_for_segment=1)
    m, off = chainutil.batch(77, nq=5)
    out, info, oo = wb.filter_mappings_batch(P, m, off, np.full(5, QLEN, dtype=np.int64), REF_LEN)
    out = np.concatenate([out, out[oo[1]: oo[1] + 2]])  # two queries sharing (refSeqId, refStartPos, queryStartPos) triples
    oo = np.append(oo, oo[-1] + 2)
    got, owner = wb.one_to_one_filter(P, out, oo, REF_LEN)
    exp = []
    for t in sorted(set(out["refSeqId"].tolist())):
        a = np.ascontiguousarray(out[out["refSeqId"] == t])
        r = np.zeros(len(a) + 1, dtype=wb.MAPPING_DTYPE)
        n = ref.ref_filter_by_group(ctypes.byref(P), ctypes.c_void_p(a.ctypes.data), ctypes.c_int64(len(a)), 0, 1, None, ctypes.c_void_p(REF_LEN.ctypes.data),
                                    ctypes.c_void_p(r.ctypes.data), ctypes.c_int64(len(r)))
        for k in r[:n]:
            for q in range(len(oo) - 1):
                sub = out[oo[q]: oo[q + 1]]
                if ((sub["refSeqId"] == k["refSeqId"]) & (sub["refStartPos"] == k["refStartPos"]) & (sub["queryStartPos"] == k["queryStartPos"])).any():
                    exp.append((q, k.tobytes()))
    assert sorted(exp) == sorted((int(q), k.tobytes()) for q, k in zip(owner, got))
    assert 0 < len(got) < len(out) and (np.diff(owner) >= 0).all()


def test_mapping_paf_round_trip_and_parse_rules():
    """parseMashmapRow + createSeqRecord (computeAlignments.hpp:195-303,611-624) on lines written by wfb_mapping_paf_format."""
    import wfmash_b200 as wb
    P = wb.FilterParams(window_length=1000)
    m = np.zeros(3, dtype=wb.MAPPING_DTYPE)
    m[0] = (1, 500, 0, 30_000, 30, 420, 9512, 0, 97)
    m[1] = (0, 200_000, 30_000, 41_000, 41, 600, 10000, 1, 88)
    m[2] = (1, 70_000, 80_000, 19_000, 19, 250, 8000, 0, 100)
    info = np.array([(0, 1, 30), (1, 1, 41), (1, 41, 41)], dtype=wb.CHAIN_INFO_DTYPE)
    names, rlen = ["tgt#1#a", "tgt#1#b"], np.array([250_000, 90_000], dtype=np.int64)
    txt = wb.mapping_paf_format(P, m, info, "qry#1#x", 100_000, names, rlen)
    lines = txt.split(b"\n")[:-1]
    assert lines[0] == b"qry#1#x\t100000\t0\t30000\t+\ttgt#1#b\t90000\t500\t30500\t420\t30000\t13\tid:f:0.9512\tkc:f:0.97\tch:Z:0.1.30"
    assert lines[1] == b"qry#1#x\t100000\t30000\t71000\t-\ttgt#1#a\t250000\t200000\t241000\t600\t41000\t255\tid:f:1\tkc:f:0.88\tch:Z:1.1.41"
    r0, qn, tn = wb.mapping_paf_parse(lines[0], 1000, 1000, 128_000)
    assert (qn, tn) == ("qry#1#x", "tgt#1#b")
    # first piece, not the last: target padded and clamped at 0, the query padding is computed but only stored for the last piece
    assert (r0.r_start, r0.r_end, r0.q_start, r0.q_end, r0.strand) == (0, 31_500, 0, 30_000, 1)
    assert (r0.chain_id, r0.chain_pos, r0.chain_length) == (0, 1, 30) and abs(r0.mashmap_estimated_identity - 0.9512) < 1e-7
    assert (r0.ref_fetch_start, r0.ref_fetch_len) == (0, 90_000)  # 128 kb of patch room on both sides, clamped to the sequence
    r1, _, _ = wb.mapping_paf_parse(lines[1], 1000, 1000, 5_000)
    assert (r1.r_start, r1.r_end, r1.strand) == (199_000, 242_000, -1) and (r1.ref_fetch_start, r1.ref_fetch_len) == (194_000, 53_000)
    r2, _, _ = wb.mapping_paf_parse(lines[2], 5000, 700, 1000)
    assert (r2.q_start, r2.q_end) == (80_000, 99_700)   # last piece of its chain (41 of 41): the end is padded; the start only for piece 1
    assert (r2.r_start, r2.r_end, r2.ref_fetch_start, r2.ref_fetch_len) == (65_000, 90_000, 64_000, 26_000)
    # 13 tokens suffice (no kc / ch tags): defaults chain -1 / 1 / 1; a non-numeric identity falls back to 0.70
    r3, _, _ = wb.mapping_paf_parse(b"q 1000 0 900 + t 5000 100 1000 9 900 20 id:f:x", 0, 50, 10)
    assert (r3.chain_id, r3.chain_length, r3.chain_pos) == (-1, 1, 1) and abs(r3.mashmap_estimated_identity - 0.70) < 1e-7
    assert (r3.q_start, r3.q_end) == (0, 950)
    for bad in (b"q 1000 0 900 + t 5000 100 1000 9 900 20", b"q 1000 0 900 + t 5000 5000 5100 9 900 20 id:f:0.9", b"q 1000 zero 900 + t 5000 1 2 9 900 20 id:f:0.9"):
        with pytest.raises(wb.WfbError):
            wb.mapping_paf_parse(bad, 0)
    # -M output carries jc:f:0 instead of the chain tag; legacy output is blank separated with inclusive ends
    assert wb.mapping_paf_format(wb.FilterParams(merge_mappings=0), m[:1], None, "q", 100_000, names, rlen).endswith(b"\tkc:f:0.97\tjc:f:0\n")
    assert wb.mapping_paf_format(wb.FilterParams(legacy_output=1), m[:1], None, "q", 100_000, names, rlen) == b"q 100000 0 29999 + tgt#1#b 90000 500 30499 951200\n"


@pytest.mark.ref
def test_filter_edge_cases_match_compiled_reference_live():
    """Empty batch, queries without mappings, one and two mappings, a mapping ending beyond the query, zero-length blocks."""
    import wfmash_b200 as wb
    ref = util.load_ref("libfilterref.so")
    if ref is None:
        pytest.skip("oracle/_ref not built (reference sources absent)")
    ref.ref_filter_subset.restype = ctypes.c_int64
    out, info, oo = wb.filter_mappings_batch(wb.FilterParams(), np.zeros(0, dtype=wb.MAPPING_DTYPE), [0], [], REF_LEN)
    assert len(out) == 0 and oo.tolist() == [0]
    one = np.zeros(1, dtype=wb.MAPPING_DTYPE); one[0] = (2, 399_500, 299_000, 1000, 1, 9, 9100, 1, 80)     # ends at the sequence ends
    two = np.zeros(2, dtype=wb.MAPPING_DTYPE); two[0] = (1, 5000, 0, 1000, 1, 20, 9900, 0, 90); two[1] = (1, 6000, 1000, 1000, 1, 22, 9800, 0, 95)
    zero = np.zeros(3, dtype=wb.MAPPING_DTYPE); zero[0] = (0, 100, 100, 0, 1, 0, 0, 0, 0); zero[1] = (0, 100, 100, 0, 1, 0, 0, 0, 0); zero[2] = (0, 7000, 7000, 1000, 1, 5, 8000, 0, 50)
    many = np.concatenate([two] * 3 + [one])                                                                  # exact duplicates
    for prm in (dict(), dict(scaffold_gap=0), dict(merge_mappings=0), dict(num_mappings_for_segment=1, scaffold_min_length=500), dict(filter_mode=3, block_length=5000)):
        P = wb.FilterParams(window_length=1000, **prm)
        batch = [one, np.zeros(0, dtype=wb.MAPPING_DTYPE), two, zero, many]
        off = np.cumsum([0] + [len(b) for b in batch])
        got, ginfo, goo = wb.filter_mappings_batch(P, np.concatenate(batch), off, [QLEN] * len(batch), REF_LEN, host_threads=2)
        for q, b in enumerate(batch):
            a = np.ascontiguousarray(b)
            o = np.zeros(len(a) + 4, dtype=wb.MAPPING_DTYPE); c = np.zeros(len(a) + 4, dtype=wb.CHAIN_INFO_DTYPE)
            n = ref.ref_filter_subset(ctypes.byref(P), ctypes.c_void_p(a.ctypes.data) if len(a) else None, ctypes.c_int64(len(a)), q, ctypes.c_int64(QLEN), None,
                                      ctypes.c_void_p(REF_LEN.ctypes.data), ctypes.c_void_p(o.ctypes.data), ctypes.c_void_p(c.ctypes.data), ctypes.c_int64(len(o)))
            assert n == goo[q + 1] - goo[q], (prm, q)
            assert got[goo[q]: goo[q + 1]].tobytes() == o[:n].tobytes() and ginfo[goo[q]: goo[q + 1]].tobytes() == c[:n].tobytes(), (prm, q)
    with pytest.raises(wb.WfbError):
        wb.filter_mappings_batch(wb.FilterParams(window_length=0), one, [0, 1], [QLEN], REF_LEN)
    with pytest.raises(wb.WfbError):
        wb.filter_mappings_batch(wb.FilterParams(skip_prefix=1), one, [0, 1], [QLEN], REF_LEN)     # skip_prefix needs the groups


def test_phase_entry_points_validate_arguments_before_touching_the_device():
    import wfmash_b200 as wb
    seqs = [("a#1#c", b"ACGT" * 300)]
    for bad in (dict(kmer_size=0), dict(window_length=0)):
        with pytest.raises(wb.WfbError):
            wb.map_phase(seqs, seqs, wb.MapPhaseParams(percentage_identity=0.9, **bad))
    with pytest.raises(wb.WfbError):
        wb.map_phase([], seqs, wb.MapPhaseParams(percentage_identity=0.9))   # no targets


@pytest.mark.ref
def test_mapping_paf_parse_matches_compiled_reference_live():
    """wfb_mapping_paf_parse against the reference's UNMODIFIED Aligner::parseMashmapRow (computeAlignments.hpp:195-303, compiled in
    place in oracle/_ref/libalignref.so) on rows written by wfb_mapping_paf_format and on malformed rows: same fields, same rejections."""
    import wfmash_b200 as wb
    A = util.load_ref("libalignref.so")
    if A is None:
        pytest.skip("oracle/_ref not built (reference sources absent)")
    (out, info, oo), m, off = ours(21, dict(w=1000), dict(), None)
    names = [f"s{i}" for i in range(len(REF_LEN))]
    rows = []
    for q in range(len(oo) - 1):
        rows += wb.mapping_paf_format(params(dict(w=1000), dict()), out[oo[q]: oo[q + 1]], info[oo[q]: oo[q + 1]], f"q{q}", QLEN, names, REF_LEN).split(b"\n")[:-1]
    rows += [b"q 1000 0 900 + t 5000 100 1000 9 900 20 id:f:x", b"q 1000 0 900 + t 5000 100 1000 9 900 20 id:f:0.93 kc:f:1 ch:Z:7.1.1",
             b"q 1000 10 990 - t 5000 4100 4990 9 900 20 id:f:0.5 kc:f:1 ch:Z:3.2.2", b"q 1000 0 900 + t 5000 100 1000 9 900 20 id:f:0.9 kc:f:1 ch:Z:bad",
             b"q 1000 0 900 + t 5000 100 1000 9 900 20", b"q 1000 0 900 + t 5000 5000 5100 9 900 20 id:f:0.9", b"q 1000 zero 900 + t 5000 1 2 9 900 20 id:f:0.9",
             b"q\t1000\t0\t900\t+\tt\t5000\t100\t6000\t9\t900\t20\tid:f:0.9", b""]
    assert len(rows) > 100
    n_ok = n_bad = 0
    for tp, qp in ((0, 0), (1000, 1000), (5000, 700), (77, 100000)):
        for line in rows:
            q0, q1, r0, r1 = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
            strand, cid, clen, cpos, ident = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32(), ctypes.c_float()
            bad = A.ref_parse_mashmap_row(line, ctypes.c_uint64(tp), ctypes.c_uint64(qp), ctypes.byref(q0), ctypes.byref(q1), ctypes.byref(r0), ctypes.byref(r1),
                                          ctypes.byref(strand), ctypes.byref(ident), ctypes.byref(cid), ctypes.byref(clen), ctypes.byref(cpos))
            try:
                row, _, _ = wb.mapping_paf_parse(line, tp, qp, 128_000)
            except wb.WfbError:
                assert bad == 1, (line, tp, qp)
                n_bad += 1
                continue
            assert bad == 0, (line, tp, qp)
            assert (row.q_start, row.q_end, row.r_start, row.r_end, row.strand, row.chain_id, row.chain_length, row.chain_pos) == \
                   (q0.value, q1.value, r0.value, r1.value, strand.value, cid.value, clen.value, cpos.value), (line, tp, qp)
            assert np.float32(row.mashmap_estimated_identity) == np.float32(ident.value)
            n_ok += 1
    assert n_ok > 400 and n_bad >= 12


@pytest.mark.ref
def test_mapping_paf_parse_matches_compiled_reference_on_mutated_rows():
    """Rows written by wfb_mapping_paf_format with fields replaced / deleted / inserted / corrupted (signs, blanks, trailing letters, overflowing
    numbers, malformed id:f: and ch:Z: tags, spaces for tabs): accept / reject and every parsed field must equal the unmodified parseMashmapRow.
    1.09 M such rows ran clean at the end of round 2; 3 000 here."""
    import random
    import wfmash_b200 as wb
    A = util.load_ref("libalignref.so")
    if A is None:
        pytest.skip("oracle/_ref not built (reference sources absent)")
    (out, info, oo), m, off = ours(21, dict(w=1000), dict(), None)
    names = [f"s{i}" for i in range(len(REF_LEN))]
    rows = []
    for q in range(len(oo) - 1):
        rows += wb.mapping_paf_format(params(dict(w=1000), dict()), out[oo[q]: oo[q + 1]], info[oo[q]: oo[q + 1]], f"q{q}", QLEN, names, REF_LEN).split(b"\n")[:-1]
    weird = [b"", b"-1", b"0", b"00012", b"+7", b" 5", b"5 ", b"12abc", b"abc", b"1e3", b"0x10", b"99999999999", b"9223372036854775807", b"18446744073709551616",
             b"-", b"+", b".", b"1.5", b"id:f:", b"id:f:nan", b"id:f:1e-3", b"id:f:0.9x", b"id:f:-0.5", b"id:f:2", b"ch:Z:1.2", b"ch:Z:1.2.3.4", b"ch:Z:a.b.c",
             b"ch:Z:-1.2.3", b"ch:Z:1..3", b"ch:Z:", b"kc:f:1", b"xx:i:3", b":", b"id:f:0.9:0.8"]
    rnd = random.Random(5)
    n_ok = n_bad = 0
    saved = os.dup(2)  # the reference logs every rejected row to stderr
    null = os.open(os.devnull, os.O_WRONLY)
    os.dup2(null, 2)
    try:
        for _ in range(3000):
            f = rnd.choice(rows).split(b"\t")
            for _ in range(rnd.choice([0, 1, 1, 1, 2, 3])):
                op = rnd.random()
                if op < 0.55:
                    f[rnd.randrange(len(f))] = rnd.choice(weird)
                elif op < 0.7 and len(f) > 1:
                    del f[rnd.randrange(len(f))]
                elif op < 0.85:
                    f.insert(rnd.randrange(len(f) + 1), rnd.choice(weird))
                else:
                    i = rnd.randrange(len(f))
                    tok = bytearray(f[i])
                    if tok:
                        tok[rnd.randrange(len(tok))] = rnd.choice(b"0123456789-+. :xZ\t")
                    f[i] = bytes(tok)
            line = rnd.choice([b"\t", b"\t", b"\t", b" "]).join(f)
            tp, qp = rnd.choice([(0, 0), (1000, 1000), (5000, 700), (77, 100000)])
            q0, q1, r0, r1 = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
            strand, cid, clen, cpos, ident = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32(), ctypes.c_float()
            bad = A.ref_parse_mashmap_row(line, ctypes.c_uint64(tp), ctypes.c_uint64(qp), ctypes.byref(q0), ctypes.byref(q1), ctypes.byref(r0), ctypes.byref(r1),
                                          ctypes.byref(strand), ctypes.byref(ident), ctypes.byref(cid), ctypes.byref(clen), ctypes.byref(cpos))
            try:
                row, _, _ = wb.mapping_paf_parse(line, tp, qp, 128_000)
            except wb.WfbError:
                assert bad == 1, (line, tp, qp)
                n_bad += 1
                continue
            assert bad == 0, (line, tp, qp)
            assert (row.q_start, row.q_end, row.r_start, row.r_end, row.strand, row.chain_id, row.chain_length, row.chain_pos) == \
                   (q0.value, q1.value, r0.value, r1.value, strand.value, cid.value, clen.value, cpos.value), (line, tp, qp)
            a, b = np.float32(row.mashmap_estimated_identity), np.float32(ident.value)
            assert a == b or (np.isnan(a) and np.isnan(b)), (line, a, b)
            n_ok += 1
    finally:
        os.dup2(saved, 2)
        os.close(saved)
        os.close(null)
    assert n_ok > 800 and n_bad > 500
