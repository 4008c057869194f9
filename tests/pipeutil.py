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


def expected(seqs, P, oracle, fref, wref, fragment_order="reverse", align=True):
    """(mapping PAF, alignment PAF lines) the reference-side pieces produce for an all-vs-all run over `seqs`.
    fragment_order: the order in which the fragments' results reach Map::filterSubsetMappings. The reference appends them as its
    fragment tasks finish (computeMap.hpp:590-597), so it is schedule dependent there, and it decides the ch:Z: tags (chain ids are
    the ranks of the smallest ORIGINAL index in each chain, mappingFilter.hpp:401-404,498-520). "reverse" = the order a one-thread
    taskflow executor runs the subflow (last emplaced fragment first): the reference's only reproducible schedule, the one the library
    follows, and the one under which this composition equals the reference's real `wfmash -m -t 1` text (reference_map_phase below);
    "forward" = fragment 0 first."""
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
        if fragment_order == "reverse":   # whole fragments in reverse order, each fragment's own results in their order
            fr = sorted(set(l2["frag"].tolist()), reverse=True)
            l2 = np.concatenate([l2[l2["frag"] == f] for f in fr]) if fr else l2
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
    if not align:
        return map_paf, []
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


class _RefMapParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in "kmer_size sketch_size threads filter_mode skip_self skip_prefix lower_triangular merge_mappings split minimum_hits".split()] + \
               [(n, ctypes.c_int64) for n in "window_length block_length chain_gap scaffold_gap scaffold_max_deviation scaffold_min_length".split()] + \
               [("max_mapping_length", ctypes.c_uint64), ("num_mappings_for_segment", ctypes.c_uint32), ("num_mappings_for_scaffold", ctypes.c_uint32),
                ("percentage_identity", ctypes.c_float), ("prefix_delim", ctypes.c_int32), ("overlap_threshold", ctypes.c_double),
                ("scaffold_overlap_threshold", ctypes.c_double), ("max_kmer_freq", ctypes.c_double)]


def reference_map_phase(M, seqs, P, threads=1, queries=None):
    """The reference's UNMODIFIED skch::Map (oracle/ref_mapper_driver.cpp -> libmapperref.so): the whole `wfmash -m` run; all-vs-all over
    `seqs`, or `queries` (a second FASTA) against `seqs`."""
    import ctypes, tempfile
    P = P.resolved()
    F = P.filter
    M.ref_map_phase.restype = ctypes.c_int64
    prm = _RefMapParams(P.kmer_size, P.sketch_size, threads, F.filter_mode, int(P.skip_self), int(P.skip_prefix), int(P.lower_triangular), F.merge_mappings, F.split,
                        P.minimum_hits, P.window_length, F.block_length, F.chain_gap, F.scaffold_gap, F.scaffold_max_deviation, F.scaffold_min_length,
                        F.max_mapping_length, F.num_mappings_for_segment, F.num_mappings_for_scaffold, P.percentage_identity, ord(P.prefix_delim),
                        F.overlap_threshold, F.scaffold_overlap_threshold, P.max_kmer_freq)
    n = len(seqs)
    names = (ctypes.c_char_p * n)(*[a.encode() for a, _ in seqs]); sq = (ctypes.c_char_p * n)(*[b for _, b in seqs]); ln = (ctypes.c_int64 * n)(*[len(b) for _, b in seqs])
    buf = ctypes.create_string_buffer(64 << 20)
    qn, qs, ql, nq, same = names, sq, ln, n, 1
    if queries is not None:
        nq, same = len(queries), 0
        qn = (ctypes.c_char_p * nq)(*[a.encode() for a, _ in queries]); qs = (ctypes.c_char_p * nq)(*[b for _, b in queries]); ql = (ctypes.c_int64 * nq)(*[len(b) for _, b in queries])
    with tempfile.TemporaryDirectory() as d:
        k = M.ref_map_phase(d.encode(), ctypes.byref(prm), names, sq, ln, n, qn, qs, ql, nq, same, buf, ctypes.c_int64(len(buf)))
    assert k >= 0
    return buf.raw[:k]


class OracleIndex:
    """Stands in for wb.Index in CPU tests of pipeline.map's HOST half: map_fragments answers from the oracle's mapping restatement
    (the GPU kernels are compared with that same restatement in tests/test_gpu_parity.py)."""

    def __init__(self, oracle, seqs, ids, groups, k, w, s, F, threads, queries=None):
        """seqs / ids: the targets; groups: group of every sequence id; queries: [(seq, id)] when they are not the targets."""
        self.o, self.seqs, self.ids, self.groups, self.k, self.w, self.s = oracle, seqs, ids, groups, k, w, s
        self.q = queries if queries is not None else list(zip(seqs, ids))
        self.index = maputil.oracle_index(oracle, [maputil.clean(x) for x in seqs], ids, k, w, s, F, threads)

    def map_fragments(self, blob, frags, fqs, minimum_hits, cutoffs, grp, skip_self=True, skip_prefix=True, lower_triangular=False, stage1_min_hits=None,
                      l2_min_shared=None, **kw):
        import wfmash_b200 as wb
        qs = [x for x, _ in self.q if len(x) >= self.w]
        qi = [i for x, i in self.q if len(x) >= self.w]
        frs, _, _, _, mp = maputil.oracle_map_fragments(self.o, self.index, qs, qi, [self.groups[i] for i in qi], self.k, self.w, self.s,
                                                        mode=(int(skip_self), int(skip_prefix), int(lower_triangular), minimum_hits),
                                                        stage1=stage1_min_hits is not None, min_shared=l2_min_shared, cut=cutoffs, ref_group=self.groups)
        assert len(frs) == len(frags)   # the same fragments in the same order (every sequence is both target and query)
        out = np.zeros(len(mp), dtype=wb.L2_MAPPING_DTYPE)
        for f in ("frag", "refSeqId", "refStartPos", "optimalStart", "optimalEnd", "conservedSketches", "strand", "nucIdentity", "kmerComplexity"):
            out[f] = mp[f]
        off = np.zeros(len(frs) + 1, dtype=np.int64)
        np.add.at(off, out["frag"] + 1, 1)
        return {"mappings": out, "offset": np.cumsum(off), "status": np.zeros(len(frs), dtype=np.int32), "l1_kernel_ms": 0.0, "l2_kernel_ms": 0.0,
                "n_l1_loci": 0}

    def close(self):
        pass
