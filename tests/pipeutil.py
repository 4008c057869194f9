"""End-to-end checker of wfmash_b200.pipeline (map -> mapping PAF -> align -> alignment PAF): the same run composed from
the reference-side pieces — the oracle's mapping restatement (pinned to the compiled reference by its own tests), the
reference's UNMODIFIED filter / output code (oracle/_ref/libfilterref.so) and the reference's UNMODIFIED
do_biwfa_alignment (oracle/_ref/libwflignref.so). TEST INFRASTRUCTURE ONLY."""
import ctypes

import numpy as np

from tests import maputil, util


def case(seed=5, length=30_000):
    """Three haplotypes of one chromosome in PanSN naming: 4 % and 8 % diverged copies of a root, the second with an
    inverted segment (reverse-strand mappings) and a 1.2 kb deletion (a chain break)."""
    from wfmash_b200 import synth
    rng = np.random.default_rng(seed)
    root = synth.random_seq(length, rng)
    b = synth.mutate(root, 0.04, rng)
    c = synth.mutate(root, 0.08, rng)
    i0, i1 = length // 3, length // 3 + 6000
    c = np.concatenate([c[:i0], np.frombuffer(maputil.revcomp(c[i0:i1].tobytes()), dtype=np.uint8), c[i1: 2 * length // 3], c[2 * length // 3 + 1200:]])
    low = root.tobytes()[:length // 2].lower() + root.tobytes()[length // 2:]  # soft-masked half: the kernels upper-case
    return [("a#1#chr1", low), ("b#1#chr1", b.tobytes()), ("c#1#chr1", c.tobytes()), ("c#1#tiny", b"ACGT" * 100)]


def expected(seqs, P, oracle, fref, wref):
    """(mapping PAF, alignment PAF lines) the reference-side pieces produce for an all-vs-all run over `seqs`."""
    import wfmash_b200 as wb
    from wfmash_b200 import pipeline
    P = P.resolved()
    k, w, s = P.kmer_size, P.window_length, P.sketch_size
    ids = pipeline.SequenceIds(seqs, seqs, P.prefix_delim if P.skip_prefix else "")
    raw = [sq for _, sq in seqs]
    sid = [ids.id_of[n] for n, _ in seqs]
    index = maputil.oracle_index(oracle, [maputil.clean(x) for x in raw], sid, k, w, s, P.max_kmer_freq, P.index_threads)
    min_hits = max(P.minimum_hits, wb.estimate_minimum_hits_relaxed(s, k, P.percentage_identity))
    shared = wb.l2_min_shared_relaxed(P.percentage_identity, k, s) if P.keep_low_pct_id else wb.l2_min_shared(P.percentage_identity, k, s)
    frs, _, _, _, mp = maputil.oracle_map_fragments(oracle, index, raw, sid, ids.group, k, w, s, mode=(int(P.skip_self), int(P.skip_prefix), int(P.lower_triangular), min_hits),
                                                    stage1=P.stage1_top_ani_filter, min_shared=shared, hg=P.hg_numerator, ani_diff=P.ani_diff,
                                                    cut=wb.sketch_cutoffs(s, k, P.ani_diff, P.ani_diff_conf, P.stage1_top_ani_filter))
    ref_len = np.array(ids.lengths, dtype=np.int64)
    groups = np.array(ids.group, dtype=np.int32)
    frag_q = np.array([qi for qi, _ in frs], dtype=np.int64)
    frag_index, seen = [], {}
    for qi, _ in frs:
        frag_index.append(seen.get(qi, 0)); seen[qi] = seen.get(qi, 0) + 1
    fi = np.array(frag_index, dtype=np.int32)
    fref.ref_filter_subset.restype = ctypes.c_int64
    fref.ref_report_mappings.restype = ctypes.c_int64
    text = []
    for qi, (name, sq) in enumerate(seqs):
        if len(sq) < w:
            continue
        l2 = mp[frag_q[mp["frag"]] == qi]
        m = np.ascontiguousarray(wb.l2_to_query_mappings(l2, fi, w, len(sq), ref_len))
        o = np.zeros(len(m) + 4, dtype=wb.MAPPING_DTYPE); c = np.zeros(len(m) + 4, dtype=wb.CHAIN_INFO_DTYPE)
        n = fref.ref_filter_subset(ctypes.byref(P.filter), ctypes.c_void_p(m.ctypes.data), ctypes.c_int64(len(m)), ids.id_of[name], ctypes.c_int64(len(sq)),
                                   ctypes.c_void_p(groups.ctypes.data), ctypes.c_void_p(ref_len.ctypes.data), ctypes.c_void_p(o.ctypes.data),
                                   ctypes.c_void_p(c.ctypes.data), ctypes.c_int64(len(o)))
        buf = ctypes.create_string_buffer(400 * n + 64)
        tn = fref.ref_report_mappings(ctypes.byref(P.filter), ctypes.c_void_p(o.ctypes.data), ctypes.c_void_p(c.ctypes.data), ctypes.c_int64(n), name.encode(),
                                      ctypes.c_int64(len(sq)), ctypes.c_void_p(ref_len.ctypes.data), buf, ctypes.c_int64(len(buf)))
        line = buf.raw[:tn]
        for i, nm in enumerate(ids.names):  # the mock id manager of the reference driver prints s<id>
            line = line.replace(b"\ts%d\t" % i, b"\t" + nm.encode() + b"\t")
        text.append(line)
    map_paf = b"".join(text)
    recs = pipeline.records_from_paf(map_paf, seqs, seqs, P)
    # do_biwfa_alignment's own text, then the re-emission of Aligner::processMappingRecord (computeAlignments.hpp:486-516: fields
    # joined by single tabs, i.e. without the trailing tab). tests/test_pipeline_emu_cpu.py also runs the reference's whole
    # align::Aligner (oracle/_ref/libalignref.so) on the same mapping PAF, which includes that step for real.
    lines = [pipeline._phase_line(util.ref_paf(wref, r, min_identity=P.min_identity, min_alignment_length=P.min_alignment_length,
                                               min_block_identity=P.min_block_identity, disable_chain_patching=P.disable_chain_patching)) for r in recs]
    return map_paf, lines


def reference_align_phase(A, mapping_paf, seqs, P, sam_format=False, emit_md_tag=False, no_seq_in_sam=False):
    """The reference's UNMODIFIED align::Aligner::compute() (oracle/ref_align_driver.cpp -> libalignref.so) on a mapping PAF."""
    import ctypes, tempfile
    P = P.resolved()
    A.ref_align_phase.restype = ctypes.c_int64
    n = len(seqs)
    names = (ctypes.c_char_p * n)(*[a.encode() for a, _ in seqs]); sq = (ctypes.c_char_p * n)(*[b for _, b in seqs]); ln = (ctypes.c_int64 * n)(*[len(b) for _, b in seqs])
    buf = ctypes.create_string_buffer(4 * sum(len(b) for _, b in seqs) * 4 + (1 << 20))
    with tempfile.TemporaryDirectory() as d:
        k = A.ref_align_phase(d.encode(), names, sq, ln, n, names, sq, ln, n, mapping_paf, ctypes.c_int64(len(mapping_paf)), ctypes.c_uint64(P.target_padding),
                              ctypes.c_uint64(P.query_padding), ctypes.c_uint64(P.window_length * 128), ctypes.c_float(P.min_identity),
                              ctypes.c_uint64(P.min_alignment_length), ctypes.c_float(P.min_block_identity), int(P.disable_chain_patching), int(sam_format),
                              int(emit_md_tag), int(no_seq_in_sam), buf, ctypes.c_int64(len(buf)))
    assert k >= 0
    return buf.raw[:k]


# (name, arguments of case(), pipeline.Params overrides; "filter" holds FilterParams overrides)
PIPELINE_CASES = [
    ("defaults_p90", dict(seed=5, length=60_000), dict()),
    # (the one-to-one mode is checked at the unit level: its line order and chain ids follow an unordered_map in the reference)
    ("w500_p85_n1", dict(seed=9, length=30_000), dict(window_length=500, percentage_identity=0.85,
                                                       filter=dict(num_mappings_for_segment=1, overlap_threshold=0.5, chain_gap=1000, scaffold_min_length=3000))),
    ("no_merge_no_filter", dict(seed=11, length=16_000), dict(filter=dict(merge_mappings=0, filter_mode=3, scaffold_gap=0), disable_chain_patching=True)),
]


def params(prm):
    import wfmash_b200 as wb
    from wfmash_b200 import pipeline
    prm = dict(prm)
    f = prm.pop("filter", None)
    P = pipeline.Params(**prm)
    if f is not None:
        P.filter = wb.FilterParams(window_length=P.window_length, percentage_identity=P.percentage_identity, skip_prefix=int(P.skip_prefix), **f)
    return P


def line_digest(line: bytes):
    import hashlib
    return {"head": b"\t".join(line.split(b"\t")[:12]).decode(), "sha": hashlib.sha256(line).hexdigest()}
