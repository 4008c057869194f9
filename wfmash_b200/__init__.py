"""wfmash_b200 — host-side mirror of the reference's call sites over the C ABI of libwfmash_b200.so.

The library is hand-written CUDA for sm_100a (see wfmash_b200/csrc). There is NO CPU path: loading
fails loudly if the shared object is missing, and every call fails with WFB_ENODEV without a GPU.
Python here is plumbing only (ctypes); torch is used by bench.py / multi-GPU sharding, not here.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# WFB_LIB lets tuning sweeps point at a differently-compiled build of the same sources
LIB_PATH = os.environ.get("WFB_LIB") or os.path.join(_HERE, "libwfmash_b200.so")

# wflign_penalties_t defaults of the CLI (src/interface/parse_args.hpp -> align::Parameters;
# do_biwfa_alignment is called with mismatch 5, gap1 (8,2), gap2 (24,1); src/align/include/computeAlignments.hpp:684-690)
WFMASH_PENALTIES = (5, 8, 2, 24, 1)


class WfbError(RuntimeError):
    pass


class _Pen(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("mismatch", "gap_opening1", "gap_extension1", "gap_opening2", "gap_extension2")]


class _Pair(ctypes.Structure):
    _fields_ = [("pattern", ctypes.c_char_p), ("pattern_len", ctypes.c_int32), ("text", ctypes.c_char_p), ("text_len", ctypes.c_int32)]


class _Res(ctypes.Structure):
    _fields_ = [("status", ctypes.c_int32), ("score", ctypes.c_int32), ("ops_offset", ctypes.c_int64),
                ("ops_len", ctypes.c_int32), ("reserved_", ctypes.c_int32)]


class _EfPair(ctypes.Structure):
    _fields_ = [("pattern", ctypes.c_char_p), ("pattern_len", ctypes.c_int32), ("text", ctypes.c_char_p), ("text_len", ctypes.c_int32),
                ("pattern_begin_free", ctypes.c_int32), ("pattern_end_free", ctypes.c_int32), ("text_begin_free", ctypes.c_int32),
                ("text_end_free", ctypes.c_int32)]


class _Record(ctypes.Structure):
    """wfb_record_t: the parameter list of do_biwfa_alignment (src/common/wflign/src/wflign.cpp:108-133)."""
    _fields_ = [("query_name", ctypes.c_char_p), ("query", ctypes.c_char_p), ("query_total_length", ctypes.c_uint64),
                ("query_offset", ctypes.c_uint64), ("query_length", ctypes.c_uint64), ("query_is_rev", ctypes.c_int32),
                ("chain_id", ctypes.c_int32), ("target_name", ctypes.c_char_p), ("target", ctypes.c_char_p),
                ("target_total_length", ctypes.c_uint64), ("target_offset", ctypes.c_uint64), ("target_length", ctypes.c_uint64),
                ("chain_length", ctypes.c_int32), ("chain_pos", ctypes.c_int32), ("mashmap_estimated_identity", ctypes.c_float),
                ("reserved_", ctypes.c_int32)]


class _PafParams(ctypes.Structure):
    _fields_ = [("disable_chain_patching", ctypes.c_int32), ("term_group", ctypes.c_int32), ("min_identity", ctypes.c_float),
                ("min_block_identity", ctypes.c_float), ("min_alignment_length", ctypes.c_uint64), ("sam_format", ctypes.c_int32),
                ("emit_md_tag", ctypes.c_int32), ("no_seq_in_sam", ctypes.c_int32), ("reserved_", ctypes.c_int32)]


REC_WRITTEN, REC_FILTERED, REC_UNALIGNED, REC_PATCH_CAP = 0, 1, 2, -1


class AlignStats(ctypes.Structure):
    _fields_ = [("cells", ctypes.c_uint64), ("extend_matches", ctypes.c_uint64), ("overlap_tests", ctypes.c_uint64),
                ("score_steps", ctypes.c_uint64), ("break_tasks", ctypes.c_uint64), ("base_tasks", ctypes.c_uint64),
                ("base_cells", ctypes.c_uint64), ("base_extend_matches", ctypes.c_uint64), ("base_score_steps", ctypes.c_uint64),
                ("levels", ctypes.c_uint64), ("kernel_ms", ctypes.c_double), ("break_kernel_ms", ctypes.c_double), ("patch_kernel_ms", ctypes.c_double),
                ("h2d_bytes", ctypes.c_uint64), ("d2h_bytes", ctypes.c_uint64), ("patch_cap_kept_main", ctypes.c_uint64),
                ("main_device_cap", ctypes.c_uint64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class Minmer(ctypes.Structure):
    """skch::MinmerInfo, 32-byte reference layout (src/map/include/base_types.hpp:28-60)."""
    _fields_ = [("hash", ctypes.c_uint64), ("wpos", ctypes.c_int64), ("wpos_end", ctypes.c_int64),
                ("seqId", ctypes.c_int32), ("strand", ctypes.c_int16), ("pad_", ctypes.c_int16)]


MINMER_DTYPE = np.dtype([("hash", "<u8"), ("wpos", "<i8"), ("wpos_end", "<i8"), ("seqId", "<i4"), ("strand", "<i2"), ("pad_", "<i2")])
FRAG_DTYPE = np.dtype([("seq_offset", "<i8"), ("len", "<i4"), ("seq_id", "<i4")])

_lib = None


def lib():
    """Load libwfmash_b200.so; fail loudly when it is missing (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise WfbError(f"{LIB_PATH} is missing: build it with `python -m wfmash_b200.build` "
                       "(nvcc, sm_100a). wfmash_b200 has no CPU implementation.")
    L = ctypes.CDLL(LIB_PATH)
    L.wfb_last_error.restype = ctypes.c_char_p
    L.wfb_version.restype = ctypes.c_char_p
    L.wfb_launch_count.restype = ctypes.c_uint64
    L.wfb_aligner_create.restype = ctypes.c_void_p
    L.wfb_aligner_create.argtypes = [ctypes.c_int, ctypes.POINTER(_Pen), ctypes.c_uint64]
    L.wfb_aligner_destroy.argtypes = [ctypes.c_void_p]
    L.wfb_align_batch.argtypes = [ctypes.c_void_p, ctypes.POINTER(_Pair), ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64,
                                  ctypes.POINTER(_Res), ctypes.POINTER(AlignStats)]
    L.wfb_align_batch_device.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64,
                                         ctypes.POINTER(_Res), ctypes.POINTER(AlignStats)]
    L.wfb_device_malloc.restype = ctypes.c_void_p
    L.wfb_device_malloc.argtypes = [ctypes.c_int, ctypes.c_uint64]
    L.wfb_device_free.argtypes = [ctypes.c_int, ctypes.c_void_p]
    L.wfb_memcpy_h2d.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64]
    if hasattr(L, "wfb_sketch_fragments"):
        L.wfb_sketch_fragments.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int32,
                                           ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.POINTER(ctypes.c_double)]
    _lib = L
    return L


def _err(rc):
    return WfbError(f"wfmash_b200 error {rc}: {lib().wfb_last_error().decode()}")


def device_count():
    return lib().wfb_device_count()


def launch_count():
    return lib().wfb_launch_count()


class AlignResult:
    __slots__ = ("status", "score", "ops")

    def __init__(self, status, score, ops):
        self.status, self.score, self.ops = status, score, ops

    def cigar(self):
        """Run-length string with M -> '=' like wfa_edit_cigar_to_string (wflign_swizzle.cpp:359-383)."""
        return ops_to_cigar(self.ops)


def ops_to_cigar(ops: bytes, match_char: str = "=") -> str:
    out = []
    i, n = 0, len(ops)
    while i < n:
        j = i
        while j < n and ops[j] == ops[i]:
            j += 1
        ch = chr(ops[i])
        out.append(f"{j - i}{match_char if ch == 'M' else ch}")
        i = j
    return "".join(out)


class Aligner:
    """Batched drop-in for the aligner object of do_biwfa_alignment
    (wfa::WFAlignerGapAffine2Pieces(0,x,o1,e1,o2,e2,Alignment,MemoryUltralow) + setHeuristicNone,
    src/common/wflign/src/wflign.cpp:136-148): align_end2end(pattern=target, text=query)."""

    def __init__(self, device=0, penalties=WFMASH_PENALTIES, workspace_bytes=0):
        self._L = lib()
        self.device = device
        pen = _Pen(*penalties)
        self._h = self._L.wfb_aligner_create(device, ctypes.byref(pen), workspace_bytes)
        if not self._h:
            raise WfbError(self._L.wfb_last_error().decode())
        self.last_stats = None

    def close(self):
        if getattr(self, "_h", None):
            self._L.wfb_aligner_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def align_end2end_batch(self, pairs):
        """pairs: sequence of (pattern: bytes, text: bytes). Returns [AlignResult]."""
        n = len(pairs)
        if n == 0:
            return []
        arr = (_Pair * n)(*[_Pair(p, len(p), t, len(t)) for p, t in pairs])
        cap = sum(len(p) + len(t) for p, t in pairs) + 16
        ops = ctypes.create_string_buffer(cap)
        res = (_Res * n)()
        stats = AlignStats()
        rc = self._L.wfb_align_batch(self._h, arr, n, ops, cap, res, ctypes.byref(stats))
        if rc != 0:
            raise _err(rc)
        self.last_stats = stats
        raw = ops.raw
        return [AlignResult(r.status, r.score, raw[r.ops_offset:r.ops_offset + r.ops_len]) for r in res]

    def align_end2end(self, pattern: bytes, text: bytes):
        return self.align_end2end_batch([(pattern, text)])[0]

    def align_ends_free_batch(self, items, term_group=8):
        """Head / tail patch alignments of do_biwfa_alignment (wflign.cpp:280-305, 368-397):
        items = [(pattern, pattern_begin_free, pattern_end_free, text, text_begin_free, text_end_free)], the argument
        order of wfa::WFAligner::alignEndsFree. term_group: 1 / 8 / 16 = scalar / AVX2 / AVX-512 build of the reference."""
        n = len(items)
        if n == 0:
            return []
        arr = (_EfPair * n)(*[_EfPair(p, len(p), t, len(t), pbf, pef, tbf, tef) for p, pbf, pef, t, tbf, tef in items])
        cap = sum(len(it[0]) + len(it[3]) for it in items) + 16
        ops = ctypes.create_string_buffer(cap)
        res = (_Res * n)()
        rc = self._L.wfb_align_endsfree_batch(ctypes.c_void_p(self._h), arr, n, term_group, ops, ctypes.c_int64(cap), res)
        if rc != 0:
            raise _err(rc)
        raw = ops.raw
        return [AlignResult(r.status, r.score, raw[r.ops_offset:r.ops_offset + r.ops_len]) for r in res]


    def wavefront_align(self, pattern: bytes, text: bytes):
        """wavefront_align-shaped single-pair call (deps/WFA2-lib/wavefront/wavefront_align.c:212):
        returns (status, cigar operations, score) with the reference's status codes (0 completed, -300 unattainable)."""
        cap = len(pattern) + len(text) + 16
        buf = ctypes.create_string_buffer(cap)
        n, sc = ctypes.c_int32(0), ctypes.c_int32(0)
        st = self._L.wfb_wavefront_align(ctypes.c_void_p(self._h), pattern, len(pattern), text, len(text), buf, cap, ctypes.byref(n), ctypes.byref(sc))
        if -100 < st < 0:
            raise _err(st)
        return st, buf.raw[: n.value], sc.value

    def biwfa_paf_batch(self, records, min_identity=0.0, min_alignment_length=0, min_block_identity=0.0,
                        disable_chain_patching=False, term_group=8, sam_format=False, emit_md_tag=False, no_seq_in_sam=False):
        """Batched wflign::wavefront::do_biwfa_alignment, PAF branch (src/common/wflign/src/wflign.cpp:108-483).
        records: dicts with the reference's parameter names: query_name, query, query_total_length, query_offset,
        query_is_rev, target_name, target, target_total_length, target_offset, mashmap_estimated_identity, chain_id,
        chain_length, chain_pos (query_length / target_length are the slice lengths).
        sam_format / emit_md_tag / no_seq_in_sam select the SAM branch (write_alignment_sam, wflign_patch.cpp:2480-2609).
        Returns (lines, status): lines[i] is the PAF line / SAM record (b"" when none), status[i] one of REC_*."""
        n = len(records)
        if n == 0:
            return [], []
        arr = (_Record * n)()
        keep = []
        for i, r in enumerate(records):
            q, t = r["query"], r["target"]
            qn, tn = r["query_name"].encode() if isinstance(r["query_name"], str) else r["query_name"], \
                r["target_name"].encode() if isinstance(r["target_name"], str) else r["target_name"]
            keep.append((q, t, qn, tn))
            arr[i] = _Record(qn, q, r.get("query_total_length", len(q)), r.get("query_offset", 0), len(q),
                             1 if r.get("query_is_rev", False) else 0, r.get("chain_id", 0), tn, t,
                             r.get("target_total_length", len(t)), r.get("target_offset", 0), len(t),
                             r.get("chain_length", 0), r.get("chain_pos", 0), r.get("mashmap_estimated_identity", 0.0), 0)
        pp = _PafParams(1 if disable_chain_patching else 0, term_group, min_identity, min_block_identity, min_alignment_length,
                        int(sam_format), int(emit_md_tag), int(no_seq_in_sam), 0)
        # PAF: the CIGAR is a fraction of the sequences; SAM carries SEQ (the query) + CIGAR (+ MD): size the first buffer for it, so the
        # WFB_ECAP retry (which repeats the alignment) stays the exception
        if sam_format:
            cap = sum((0 if no_seq_in_sam else len(k[0])) + (len(k[0]) + len(k[1])) // 2 + (len(k[1]) // 2 if emit_md_tag else 0) for k in keep) + 1024 * n + 4096
        else:
            cap = sum(len(k[0]) + len(k[1]) for k in keep) // 2 + 512 * n + 4096
        offs = (ctypes.c_int64 * (n + 1))()
        st = (ctypes.c_int32 * n)()
        stats = AlignStats()
        out_len = ctypes.c_int64(0)
        for _ in range(2):
            out = ctypes.create_string_buffer(cap)
            rc = self._L.wfb_biwfa_paf_batch(ctypes.c_void_p(self._h), arr, n, ctypes.byref(pp), out, ctypes.c_int64(cap),
                                             ctypes.byref(out_len), offs, st, ctypes.byref(stats))
            if rc == -5 and out_len.value > cap:  # WFB_ECAP: retry once with the size the library asked for
                cap = out_len.value + 16
                continue
            break
        if rc != 0:
            raise _err(rc)
        self.last_stats = stats
        raw = out.raw
        return [raw[offs[i]:offs[i + 1]] for i in range(n)], list(st)


class MinmerStats(ctypes.Structure):
    _fields_ = [("stream_kernel_ms", ctypes.c_double), ("total_kernel_ms", ctypes.c_double), ("bases", ctypes.c_uint64),
                ("raw_records", ctypes.c_uint64), ("chunks", ctypes.c_uint64), ("stale_absorbed", ctypes.c_uint64),
                ("stitch_miss", ctypes.c_uint64), ("filtered", ctypes.c_uint64), ("candidates", ctypes.c_uint64),
                ("redo_chunks", ctypes.c_uint64), ("cand_kernel_ms", ctypes.c_double), ("filtered_stream_ms", ctypes.c_double),
                ("redo_ms", ctypes.c_double), ("tie_sequences", ctypes.c_uint64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def minmers_build(seqs, seq_ids, kmer_size: int, window_size: int, sketch_size: int, device: int = 0):
    """Batched CommonFunc::addMinmers over target sequences (src/map/include/commonFunc.hpp:439-708), output in
    Sketch::build's order (winSketch.hpp:424-429). seqs: list of bytes. Returns (minmers, MinmerStats)."""
    L = lib()
    n = len(seqs)
    ptrs = (ctypes.c_char_p * max(n, 1))(*seqs)
    lens = (ctypes.c_int64 * max(n, 1))(*[len(s) for s in seqs])
    ids = (ctypes.c_int32 * max(n, 1))(*seq_ids)
    cap = int(sum(len(s) for s in seqs) * (0.01 * sketch_size + 0.05)) + 4096
    out = np.zeros(cap, dtype=MINMER_DTYPE)
    cnt = ctypes.c_int64(0)
    st = MinmerStats()
    rc = L.wfb_minmers_build(device, ptrs, lens, ids, n, kmer_size, window_size, sketch_size,
                             ctypes.c_void_p(out.ctypes.data), ctypes.c_int64(cap), ctypes.byref(cnt), ctypes.byref(st))
    if rc != 0:
        raise _err(rc)
    return out[: cnt.value], st


class _IndexParams(ctypes.Structure):
    _fields_ = [("kmer_size", ctypes.c_int32), ("window_size", ctypes.c_int32), ("sketch_size", ctypes.c_int32),
                ("index_threads", ctypes.c_int32), ("max_kmer_freq", ctypes.c_double)]


class IndexStats(ctypes.Structure):
    _fields_ = [("minmer", MinmerStats), ("index_kernel_ms", ctypes.c_double), ("total_windows", ctypes.c_uint64),
                ("kept_minmers", ctypes.c_uint64), ("interval_points", ctypes.c_uint64), ("unique_hashes", ctypes.c_uint64),
                ("count_threshold", ctypes.c_uint64), ("table_buckets", ctypes.c_uint64)]


class _L1Params(ctypes.Structure):
    _fields_ = [("minimum_hits", ctypes.c_int32), ("sketch_cutoffs", ctypes.c_void_p), ("n_cutoffs", ctypes.c_int32),
                ("ref_group", ctypes.c_void_p), ("n_ref_group", ctypes.c_int32), ("skip_self", ctypes.c_int32),
                ("skip_prefix", ctypes.c_int32), ("lower_triangular", ctypes.c_int32), ("kmer_complexity_threshold", ctypes.c_float)]


class _L1Out(ctypes.Structure):
    _fields_ = [("q_minmers", ctypes.c_void_p), ("q_count", ctypes.c_void_p), ("q_complexity", ctypes.c_void_p),
                ("loci", ctypes.c_void_p), ("loci_cap", ctypes.c_int64), ("frag_loci_offset", ctypes.c_void_p),
                ("frag_loci_count", ctypes.c_void_p), ("frag_status", ctypes.c_void_p), ("n_loci", ctypes.c_int64),
                ("kernel_ms", ctypes.c_double)]


L1_LOCUS_DTYPE = np.dtype([("seqId", "<i4"), ("intersectionSize", "<i4"), ("rangeStartPos", "<i8"), ("rangeEndPos", "<i8")])
FRAG_QUERY_DTYPE = np.dtype([("q_seq_id", "<i4"), ("q_group", "<i4")])


class Index:
    """GPU-resident drop-in for skch::Sketch (src/map/include/winSketch.hpp:63-154,175-457): build() =
    Sketch::build, export() = the public members minmerIndex / minmerPosLookupIndex, l1() =
    Map::doL1Mapping (src/map/include/computeMap.hpp:945-983) for a batch of fragments."""

    def __init__(self, seqs, seq_ids, kmer_size, window_size, sketch_size, max_kmer_freq=0.0002, index_threads=1, device=0):
        L = lib()
        L.wfb_index_build.restype = ctypes.c_void_p
        L.wfb_index_free.argtypes = [ctypes.c_void_p]
        self._L = L
        n = len(seqs)
        prm = _IndexParams(kmer_size, window_size, sketch_size, index_threads, max_kmer_freq)
        ptrs = (ctypes.c_char_p * n)(*seqs)
        lens = (ctypes.c_int64 * n)(*[len(s) for s in seqs])
        ids = (ctypes.c_int32 * n)(*seq_ids)
        self.stats = IndexStats()
        self.k, self.w, self.s = kmer_size, window_size, sketch_size
        self._h = L.wfb_index_build(device, ctypes.byref(prm), ptrs, lens, ids, n, ctypes.byref(self.stats))
        if not self._h:
            raise WfbError(L.wfb_last_error().decode())

    def close(self):
        if getattr(self, "_h", None):
            self._L.wfb_index_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def export(self):
        """(kept minmers, uhash, ustart, ucount, packed points)."""
        st = self.stats
        mi = np.zeros(max(int(st.kept_minmers), 1), dtype=MINMER_DTYPE)
        uh = np.zeros(max(int(st.unique_hashes), 1), dtype=np.uint64)
        us = np.zeros_like(uh, dtype=np.uint32)
        uc = np.zeros_like(uh, dtype=np.uint32)
        pts = np.zeros(max(int(st.interval_points), 1), dtype=np.uint64)
        rc = self._L.wfb_index_export(ctypes.c_void_p(self._h), ctypes.c_void_p(mi.ctypes.data), ctypes.c_int64(len(mi)),
                                      ctypes.c_void_p(uh.ctypes.data), ctypes.c_void_p(us.ctypes.data), ctypes.c_void_p(uc.ctypes.data),
                                      ctypes.c_int64(len(uh)), ctypes.c_void_p(pts.ctypes.data), ctypes.c_int64(len(pts)))
        if rc != 0:
            raise _err(rc)
        return (mi[: int(st.kept_minmers)], uh[: int(st.unique_hashes)], us[: int(st.unique_hashes)], uc[: int(st.unique_hashes)],
                pts[: int(st.interval_points)])

    def l1(self, seq: bytes, frags, frag_queries, minimum_hits, sketch_cutoffs, ref_group, skip_self=True, skip_prefix=True,
           lower_triangular=False, loci_cap=None):
        fr = np.ascontiguousarray(frags, dtype=FRAG_DTYPE)
        fq = np.ascontiguousarray(frag_queries, dtype=FRAG_QUERY_DTYPE)
        n = int(fr.shape[0])
        cut = np.ascontiguousarray(sketch_cutoffs, dtype=np.int32)
        grp = np.ascontiguousarray(ref_group, dtype=np.int32)
        lp = _L1Params(minimum_hits, cut.ctypes.data, len(cut), grp.ctypes.data, len(grp), int(skip_self), int(skip_prefix),
                       int(lower_triangular), 0.0)
        loci_cap = loci_cap or (64 * n + 1024)
        qm = np.zeros((max(n, 1), self.s), dtype=MINMER_DTYPE)
        qn = np.zeros(max(n, 1), dtype=np.int32)
        kc = np.zeros(max(n, 1), dtype=np.float32)
        loci = np.zeros(loci_cap, dtype=L1_LOCUS_DTYPE)
        off = np.zeros(max(n, 1), dtype=np.int64)
        cnt = np.zeros(max(n, 1), dtype=np.int32)
        stt = np.zeros(max(n, 1), dtype=np.int32)
        out = _L1Out(qm.ctypes.data, qn.ctypes.data, kc.ctypes.data, loci.ctypes.data, loci_cap, off.ctypes.data, cnt.ctypes.data,
                     stt.ctypes.data, 0, 0.0)
        rc = self._L.wfb_l1_batch(ctypes.c_void_p(self._h), ctypes.byref(lp), seq, ctypes.c_int64(len(seq)),
                                  ctypes.c_void_p(fr.ctypes.data), ctypes.c_void_p(fq.ctypes.data), n, ctypes.byref(out))
        if rc != 0:
            raise _err(rc)
        return {"q_minmers": qm[:n], "q_count": qn[:n], "q_complexity": kc[:n], "loci": loci[: out.n_loci], "offset": off[:n],
                "count": cnt[:n], "status": stt[:n], "kernel_ms": out.kernel_ms}


    def map_fragments(self, seq: bytes, frags, frag_queries, minimum_hits, sketch_cutoffs, ref_group, skip_self=True, skip_prefix=True,
                      lower_triangular=False, stage1_min_hits=None, l2_min_shared=None, with_l1=False, mappings_cap=None):
        """Map::mapSingleQueryFrag's L1 + L2 stages for a batch of fragments (src/map/include/computeMap.hpp:879-921,
        945-1061; MappingCore::computeL2MappedRegions, mappingCore.hpp:306-442). The L1 loci stay in device memory;
        with_l1=True also returns them. stage1_min_hits / l2_min_shared: tables indexed by Q.sketchSize (see
        stage1_min_hits(), l2_min_shared()) or None (filter off)."""
        fr = np.ascontiguousarray(frags, dtype=FRAG_DTYPE)
        fq = np.ascontiguousarray(frag_queries, dtype=FRAG_QUERY_DTYPE)
        n = int(fr.shape[0])
        cut = np.ascontiguousarray(sketch_cutoffs, dtype=np.int32)
        grp = np.ascontiguousarray(ref_group, dtype=np.int32)
        lp = _L1Params(minimum_hits, cut.ctypes.data, len(cut), grp.ctypes.data, len(grp), int(skip_self), int(skip_prefix),
                       int(lower_triangular), 0.0)
        s1 = None if stage1_min_hits is None else np.ascontiguousarray(stage1_min_hits, dtype=np.int32)
        ms = None if l2_min_shared is None else np.ascontiguousarray(l2_min_shared, dtype=np.int32)
        l2p = _L2Params(s1.ctypes.data if s1 is not None else None, len(s1) if s1 is not None else 0,
                        ms.ctypes.data if ms is not None else None, len(ms) if ms is not None else 0)
        cap = mappings_cap or (32 * n + 1024)
        maps = np.zeros(cap, dtype=L2_MAPPING_DTYPE)
        moff = np.zeros(n + 1, dtype=np.int64)
        stt = np.zeros(max(n, 1), dtype=np.int32)
        l1 = None
        keep = []
        if with_l1:
            loci_cap = 64 * n + 1024
            qm = np.zeros((max(n, 1), self.s), dtype=MINMER_DTYPE)
            qn = np.zeros(max(n, 1), dtype=np.int32)
            kc = np.zeros(max(n, 1), dtype=np.float32)
            loci = np.zeros(loci_cap, dtype=L1_LOCUS_DTYPE)
            off = np.zeros(max(n, 1), dtype=np.int64)
            cnt = np.zeros(max(n, 1), dtype=np.int32)
            st1 = np.zeros(max(n, 1), dtype=np.int32)
            l1 = _L1Out(qm.ctypes.data, qn.ctypes.data, kc.ctypes.data, loci.ctypes.data, loci_cap, off.ctypes.data, cnt.ctypes.data,
                        st1.ctypes.data, 0, 0.0)
            keep = [qm, qn, kc, loci, off, cnt, st1]
        out = _MapOut(maps.ctypes.data, cap, moff.ctypes.data, stt.ctypes.data, ctypes.addressof(l1) if l1 is not None else None,
                      0, 0.0, 0.0, 0.0, 0, 0, 0)
        rc = self._L.wfb_map_fragments_batch(ctypes.c_void_p(self._h), ctypes.byref(lp), ctypes.byref(l2p), seq, ctypes.c_int64(len(seq)),
                                             ctypes.c_void_p(fr.ctypes.data), ctypes.c_void_p(fq.ctypes.data), n, ctypes.byref(out))
        if rc != 0:
            raise _err(rc)
        res = {"mappings": maps[: out.n_mappings], "offset": moff, "status": stt[:n], "l1_kernel_ms": out.l1_kernel_ms,
               "l2_kernel_ms": out.l2_kernel_ms, "sort_kernel_ms": out.sort_kernel_ms, "n_l1_loci": out.n_l1_loci, "l2_loci": out.l2_loci,
               "l2_steps": out.l2_steps}
        if with_l1:
            qm, qn, kc, loci, off, cnt, st1 = keep
            res["l1"] = {"q_minmers": qm[:n], "q_count": qn[:n], "q_complexity": kc[:n], "loci": loci[: l1.n_loci], "offset": off[:n],
                         "count": cnt[:n], "status": st1[:n]}
        return res


class _L2Params(ctypes.Structure):
    _fields_ = [("stage1_min_hits", ctypes.c_void_p), ("n_stage1_min_hits", ctypes.c_int32), ("l2_min_shared", ctypes.c_void_p),
                ("n_l2_min_shared", ctypes.c_int32)]


class _MapOut(ctypes.Structure):
    _fields_ = [("mappings", ctypes.c_void_p), ("mappings_cap", ctypes.c_int64), ("frag_map_offset", ctypes.c_void_p),
                ("frag_status", ctypes.c_void_p), ("l1", ctypes.c_void_p), ("n_mappings", ctypes.c_int64), ("l1_kernel_ms", ctypes.c_double),
                ("l2_kernel_ms", ctypes.c_double), ("sort_kernel_ms", ctypes.c_double), ("n_l1_loci", ctypes.c_uint64),
                ("l2_loci", ctypes.c_uint64), ("l2_steps", ctypes.c_uint64)]


L2_MAPPING_DTYPE = np.dtype([("frag", "<i4"), ("refSeqId", "<i4"), ("refStartPos", "<i8"), ("optimalStart", "<i8"), ("optimalEnd", "<i8"),
                             ("conservedSketches", "<i4"), ("strand", "<i4"), ("nucIdentity", "<f4"), ("kmerComplexity", "<f4")])


def stage1_min_hits(kmer_size: int, sketch_size: int, hg_numerator: float = 1.0, ani_diff: float = 0.0):
    """Table of the stage-1 top-ANI test of Map::doL2Mapping (computeMap.hpp:999-1012), indexed by Q.sketchSize."""
    out = np.zeros(sketch_size + 1, dtype=np.int32)
    rc = lib().wfb_stage1_min_hits(ctypes.c_double(hg_numerator), ctypes.c_float(ani_diff), kmer_size, sketch_size, ctypes.c_void_p(out.ctypes.data))
    if rc != 0:
        raise _err(rc)
    return out


def l2_min_shared(percentage_identity: float, kmer_size: int, sketch_size: int):
    """Table of the identity test of Map::doL2Mapping (computeMap.hpp:1018-1024) for keep_low_pct_id == false."""
    out = np.zeros(sketch_size + 1, dtype=np.int32)
    rc = lib().wfb_l2_min_shared(ctypes.c_float(percentage_identity), kmer_size, sketch_size, ctypes.c_void_p(out.ctypes.data))
    if rc != 0:
        raise _err(rc)
    return out


def sketch_fragments(seq: bytes, frags, kmer_size: int, sketch_size: int, device: int = 0):
    """Batched CommonFunc::sketchSequence (src/map/include/commonFunc.hpp:217-323).
    frags: array-like of (seq_offset, len, seq_id). Returns (minmers[n, sketch_size], counts[n], kernel_ms)."""
    L = lib()
    fr = np.ascontiguousarray(np.array(frags, dtype=FRAG_DTYPE) if not isinstance(frags, np.ndarray) else frags)
    n = int(fr.shape[0])
    out = np.zeros((max(n, 1), sketch_size), dtype=MINMER_DTYPE)
    cnt = np.zeros(max(n, 1), dtype=np.int32)
    ms = ctypes.c_double(0.0)
    rc = L.wfb_sketch_fragments(device, seq, len(seq), fr.ctypes.data, n, kmer_size, sketch_size,
                                out.ctypes.data, cnt.ctypes.data, ctypes.byref(ms))
    if rc != 0:
        raise _err(rc)
    return out[:n], cnt[:n], ms.value


# ---- path 1 -> path 2 hand-over (host stage) ---------------------------------------------------------
MAPPING_DTYPE = np.dtype([("refSeqId", "<u4"), ("refStartPos", "<u4"), ("queryStartPos", "<u4"), ("blockLength", "<u4"), ("n_merged", "<u4"),
                          ("conservedSketches", "<u4"), ("nucIdentity", "<u2"), ("flags", "u1"), ("kmerComplexity", "u1")])  # skch::MappingResult
CHAIN_INFO_DTYPE = np.dtype([("chainId", "<u4"), ("chainPos", "<u2"), ("chainLen", "<u2")])  # skch::ChainInfo


class _ChainParams(ctypes.Structure):
    _fields_ = [("split", ctypes.c_int32), ("reserved_", ctypes.c_int32), ("chain_gap", ctypes.c_int64), ("window_length", ctypes.c_int64),
                ("max_mapping_length", ctypes.c_uint64)]


def l2_to_query_mappings(l2, frag_index, window_length: int, query_len: int, ref_seq_len):
    """Map::processFragment + mappingBoundarySanityCheck for the L2 mappings of ONE query (computeMap.hpp:121-127,
    1029-1046; mappingOutput.hpp:31-69): returns skch::MappingResult records."""
    l2 = np.ascontiguousarray(l2, dtype=L2_MAPPING_DTYPE)
    fi = np.ascontiguousarray(frag_index, dtype=np.int32)
    rl = np.ascontiguousarray(ref_seq_len, dtype=np.int64)
    out = np.zeros(max(len(l2), 1), dtype=MAPPING_DTYPE)
    rc = lib().wfb_l2_to_query_mappings(ctypes.c_void_p(l2.ctypes.data), ctypes.c_int64(len(l2)), ctypes.c_void_p(fi.ctypes.data),
                                        ctypes.c_int64(window_length), ctypes.c_int64(query_len), ctypes.c_void_p(rl.ctypes.data),
                                        ctypes.c_void_p(out.ctypes.data))
    if rc != 0:
        raise _err(rc)
    return out[: len(l2)]


def chain_mappings_batch(mappings, query_offset, window_length: int, chain_gap: int = 2000, max_mapping_length: int = 50000, split: bool = True,
                         host_threads: int = 0):
    """MappingFilterUtils::mergeMappingsInRangeWithChains (mappingFilter.hpp:381-571) for a batch of queries.
    Returns (mappings reordered like the reference's readMappings, merged mappings, chain info, merged_offset)."""
    m = np.array(mappings, dtype=MAPPING_DTYPE, copy=True)
    qo = np.ascontiguousarray(query_offset, dtype=np.int64)
    nq = len(qo) - 1
    cap = len(m) + 16
    merged = np.zeros(cap, dtype=MAPPING_DTYPE)
    info = np.zeros(cap, dtype=CHAIN_INFO_DTYPE)
    mo = np.zeros(nq + 1, dtype=np.int64)
    prm = _ChainParams(int(split), 0, chain_gap, window_length, max_mapping_length)
    rc = lib().wfb_chain_mappings_batch(ctypes.byref(prm), ctypes.c_void_p(m.ctypes.data), ctypes.c_void_p(qo.ctypes.data), nq,
                                        ctypes.c_void_p(merged.ctypes.data), ctypes.c_void_p(info.ctypes.data), ctypes.c_int64(cap),
                                        ctypes.c_void_p(mo.ctypes.data), host_threads)
    if rc != 0:
        raise _err(rc)
    return m, merged[: mo[-1]], info[: mo[-1]], mo


# ---- SURVEY 8 f2 (second part) + b3: filters, mapping PAF writer / reader ----------------------------------------------
FILTER_MAP, FILTER_ONETOONE, FILTER_NONE = 1, 2, 3


class FilterParams(ctypes.Structure):
    """wfb_filter_params_t; the defaults are the reference's CLI defaults (src/interface/parse_args.hpp)."""
    _fields_ = [("split", ctypes.c_int32), ("merge_mappings", ctypes.c_int32), ("filter_mode", ctypes.c_int32), ("skip_prefix", ctypes.c_int32),
                ("filter_length_mismatches", ctypes.c_int32), ("drop_rand", ctypes.c_int32), ("threads", ctypes.c_int32), ("legacy_output", ctypes.c_int32),
                ("chain_gap", ctypes.c_int64), ("window_length", ctypes.c_int64), ("block_length", ctypes.c_int64),
                ("max_mapping_length", ctypes.c_uint64), ("sparsity_hash_threshold", ctypes.c_uint64),
                ("num_mappings_for_segment", ctypes.c_uint32), ("num_mappings_for_scaffold", ctypes.c_uint32),
                ("overlap_threshold", ctypes.c_double), ("scaffold_overlap_threshold", ctypes.c_double),
                ("scaffold_gap", ctypes.c_int64), ("scaffold_max_deviation", ctypes.c_int64), ("scaffold_min_length", ctypes.c_int64),
                ("percentage_identity", ctypes.c_float), ("reserved_", ctypes.c_int32)]

    def __init__(self, window_length=1000, **kw):
        d = dict(split=1, merge_mappings=1, filter_mode=FILTER_MAP, skip_prefix=0, filter_length_mismatches=1, drop_rand=0, threads=2, legacy_output=0,
                 chain_gap=2000, window_length=window_length, block_length=0, max_mapping_length=50000, sparsity_hash_threshold=2**64 - 1,
                 num_mappings_for_segment=2**32 - 1, num_mappings_for_scaffold=1, overlap_threshold=0.95, scaffold_overlap_threshold=0.5,
                 scaffold_gap=100000, scaffold_max_deviation=100000, scaffold_min_length=10000, percentage_identity=0.7, reserved_=0)
        unknown = set(kw) - set(d)
        if unknown:
            raise TypeError(f"unknown filter parameter(s): {sorted(unknown)}")
        d.update(kw)
        super().__init__(**d)


class MappingRow(ctypes.Structure):
    """wfb_mapping_row_t: align::MappingBoundaryRow + the target fetch range of createSeqRecord."""
    _fields_ = [(n, ctypes.c_int64) for n in ("q_start", "q_end", "r_start", "r_end", "ref_fetch_start", "ref_fetch_len", "query_len", "ref_len",
                                               "chain_id", "chain_length", "chain_pos")] + \
               [("strand", ctypes.c_int32), ("mashmap_estimated_identity", ctypes.c_float), ("q_name_off", ctypes.c_int32), ("q_name_len", ctypes.c_int32),
                ("r_name_off", ctypes.c_int32), ("r_name_len", ctypes.c_int32)]


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else None


def filter_mappings_batch(params: FilterParams, mappings, query_offset, query_len, ref_seq_len, ref_group=None, host_threads: int = 0):
    """Map::filterSubsetMappings (computeMap.hpp:1076-1165) for a batch of queries -> (mappings, chain info, out_offset)."""
    m = np.ascontiguousarray(mappings, dtype=MAPPING_DTYPE)
    qo, ql, rl = _i64(query_offset), _i64(query_len), _i64(ref_seq_len)
    rg = np.ascontiguousarray(ref_group, dtype=np.int32) if ref_group is not None else None
    nq = len(qo) - 1
    cap = len(m) + 16
    out = np.zeros(cap, dtype=MAPPING_DTYPE)
    info = np.zeros(cap, dtype=CHAIN_INFO_DTYPE)
    oo = np.zeros(nq + 1, dtype=np.int64)
    rc = lib().wfb_filter_mappings_batch(ctypes.byref(params), _ptr(m), _ptr(qo), _ptr(ql), nq, _ptr(rg), _ptr(rl), _ptr(out), _ptr(info),
                                         ctypes.c_int64(cap), _ptr(oo), host_threads)
    if rc != 0:
        raise _err(rc)
    return out[: oo[-1]], info[: oo[-1]], oo


def filter_by_group(params: FilterParams, mappings, n_mappings: int, filter_ref: bool, ref_seq_len, ref_group=None):
    """MappingFilterUtils::filterByGroup (mappingFilter.hpp:220-293) -> (input as the reference reorders it, survivors)."""
    m = np.array(mappings, dtype=MAPPING_DTYPE, copy=True)
    rl = _i64(ref_seq_len)
    rg = np.ascontiguousarray(ref_group, dtype=np.int32) if ref_group is not None else None
    out = np.zeros(len(m) + 1, dtype=MAPPING_DTYPE)
    L = lib()
    L.wfb_filter_by_group.restype = ctypes.c_int64
    n = L.wfb_filter_by_group(ctypes.byref(params), _ptr(m), ctypes.c_int64(len(m)), ctypes.c_int32(n_mappings), int(filter_ref), _ptr(rg), _ptr(rl),
                              _ptr(out), ctypes.c_int64(len(out)))
    if n < 0:
        raise _err(n)
    return m, out[:n]


def one_to_one_filter(params: FilterParams, mappings, query_offset, ref_seq_len, ref_group=None):
    """Final reference-axis pass of the one-to-one mode (computeMap.hpp:788-850) -> (mappings grouped by query, owner query of each)."""
    m = np.ascontiguousarray(mappings, dtype=MAPPING_DTYPE)
    qo, rl = _i64(query_offset), _i64(ref_seq_len)
    rg = np.ascontiguousarray(ref_group, dtype=np.int32) if ref_group is not None else None
    cap = 4 * len(m) + 16
    out = np.zeros(cap, dtype=MAPPING_DTYPE)
    owner = np.zeros(cap, dtype=np.int32)
    L = lib()
    L.wfb_one_to_one_filter.restype = ctypes.c_int64
    n = L.wfb_one_to_one_filter(ctypes.byref(params), _ptr(m), _ptr(qo), len(qo) - 1, _ptr(rg), _ptr(rl), _ptr(out), _ptr(owner), ctypes.c_int64(cap))
    if n < 0:
        raise _err(n)
    return out[:n], owner[:n]


def mapping_paf_format(params: FilterParams, mappings, chain, query_name: str, query_len: int, ref_names, ref_seq_len) -> bytes:
    """OutputHandler::reportReadMappings (mappingOutput.hpp:74-139): the `wfmash -m` lines of one query."""
    m = np.ascontiguousarray(mappings, dtype=MAPPING_DTYPE)
    c = np.ascontiguousarray(chain, dtype=CHAIN_INFO_DTYPE) if chain is not None else None
    rl = _i64(ref_seq_len)
    names = (ctypes.c_char_p * len(ref_names))(*[s.encode() for s in ref_names])
    L = lib()
    L.wfb_mapping_paf_format.restype = ctypes.c_int64
    cap = 256 * len(m) + 256
    while True:
        buf = ctypes.create_string_buffer(cap)
        need = ctypes.c_int64(0)
        n = L.wfb_mapping_paf_format(ctypes.byref(params), _ptr(m), _ptr(c), ctypes.c_int64(len(m)), query_name.encode(), ctypes.c_int64(query_len), names,
                                     _ptr(rl), buf, ctypes.c_int64(cap), ctypes.byref(need))
        if n == -5 and need.value > cap:
            cap = need.value
            continue
        if n < 0:
            raise _err(n)
        return buf.raw[:n]


def mapping_paf_parse(line: bytes, target_padding: int, query_padding: int = 0, wflign_max_len_minor: int = 128000):
    """Aligner::parseMashmapRow + the fetch range of createSeqRecord (computeAlignments.hpp:195-303,611-624).
    Returns (MappingRow, query name, target name); raises WfbError where the reference throws."""
    row = MappingRow()
    rc = lib().wfb_mapping_paf_parse(line, ctypes.c_int64(len(line)), ctypes.c_uint64(target_padding), ctypes.c_uint64(query_padding),
                                     ctypes.c_uint64(wflign_max_len_minor), ctypes.byref(row))
    if rc != 0:
        raise _err(rc)
    return row, line[row.q_name_off: row.q_name_off + row.q_name_len].decode(), line[row.r_name_off: row.r_name_off + row.r_name_len].decode()


# ---- run-level constants of the mapping path (host) ---------------------------------------------------------------------
def sketch_size(percentage_identity: float, window_length: int, kmer_size: int) -> int:
    """param.sketchSize when -s is not given (parse_args.hpp:642-644)."""
    return int(lib().wfb_sketch_size(ctypes.c_float(percentage_identity), ctypes.c_int64(window_length), kmer_size))


def estimate_minimum_hits_relaxed(sketch_size: int, kmer_size: int, percentage_identity: float, confidence_interval: float = 0.95) -> int:
    """skch::Stat::estimateMinimumHitsRelaxed (map_stats.hpp:159-180)."""
    r = lib().wfb_estimate_minimum_hits_relaxed(sketch_size, kmer_size, ctypes.c_float(percentage_identity), ctypes.c_float(confidence_interval))
    if r < 0:
        raise _err(r)
    return int(r)


def sketch_cutoffs(sketch_size: int, kmer_size: int, ani_diff: float = 0.0, ani_diff_conf: float = 0.999, stage1_top_ani_filter: bool = True):
    """Map::sketchCutoffs (computeMap.hpp:150,234-293)."""
    out = np.zeros(min(sketch_size, 1000) + 1, dtype=np.int32)
    rc = lib().wfb_sketch_cutoffs(sketch_size, kmer_size, ctypes.c_float(ani_diff), ctypes.c_float(ani_diff_conf), int(stage1_top_ani_filter),
                                  _ptr(out), len(out))
    if rc != 0:
        raise _err(rc)
    return out


def l2_min_shared_relaxed(percentage_identity: float, kmer_size: int, sketch_size: int, confidence_interval: float = 0.95):
    """Identity test of Map::doL2Mapping with keep_low_pct_id (computeMap.hpp:1016-1024), as a table over Q.sketchSize."""
    out = np.zeros(sketch_size + 1, dtype=np.int32)
    rc = lib().wfb_l2_min_shared_relaxed(ctypes.c_float(percentage_identity), kmer_size, sketch_size, ctypes.c_float(confidence_interval), _ptr(out))
    if rc != 0:
        raise _err(rc)
    return out


# ---- ANI auto-identity ---------------------------------------------------------------------------------------
class AniStats(ctypes.Structure):
    _fields_ = [("hash_kernel_ms", ctypes.c_double), ("sort_kernel_ms", ctypes.c_double), ("bases", ctypes.c_uint64), ("valid_kmers", ctypes.c_uint64),
                ("candidates", ctypes.c_uint64), ("tiles", ctypes.c_uint64), ("passes", ctypes.c_int32), ("reserved_", ctypes.c_int32)]


def ani_group_sketches(seqs, seq_group, n_groups: int, kmer_size: int = 21, sketch_size: int = 4096, device: int = 0):
    """Per-group bottom-`sketch_size` multiset MinHash on the GPU (Stat::estimate_identity_for_groups, map_stats.hpp:563-637).
    seq_group[i] = dense group index of seqs[i]. Returns (sketches[n_groups, sketch_size] ascending, counts[n_groups], AniStats)."""
    n = len(seqs)
    ptrs = (ctypes.c_char_p * max(n, 1))(*seqs)
    lens = (ctypes.c_int64 * max(n, 1))(*[len(x) for x in seqs])
    grp = (ctypes.c_int32 * max(n, 1))(*seq_group)
    sk = np.zeros((n_groups, sketch_size), dtype=np.uint64)
    cnt = np.zeros(n_groups, dtype=np.int32)
    st = AniStats()
    rc = lib().wfb_ani_group_sketches(device, ptrs, lens, grp, n, n_groups, kmer_size, sketch_size, _ptr(sk), _ptr(cnt), ctypes.byref(st))
    if rc != 0:
        raise _err(rc)
    return sk, cnt, st


def ani_estimate_identity(q_sketch, q_count, q_group, t_sketch, t_count, t_group, kmer_size: int = 21, ani_percentile: int = 50,
                          ani_adjustment: float = -2.0):
    """Host part of the ANI estimate (map_stats.hpp:690-800) -> (identity the CLI adopts, number of group comparisons)."""
    qs, ts = np.ascontiguousarray(q_sketch, dtype=np.uint64), np.ascontiguousarray(t_sketch, dtype=np.uint64)
    qc, tc = np.ascontiguousarray(q_count, dtype=np.int32), np.ascontiguousarray(t_count, dtype=np.int32)
    qg, tg = np.ascontiguousarray(q_group, dtype=np.int32), np.ascontiguousarray(t_group, dtype=np.int32)
    assert qs.ndim == 2 and ts.ndim == 2 and qs.shape[1] == ts.shape[1]
    L = lib()
    L.wfb_ani_estimate_identity.restype = ctypes.c_double
    ncmp = ctypes.c_int32(0)
    v = L.wfb_ani_estimate_identity(_ptr(qs), _ptr(qc), _ptr(qg), len(qg), _ptr(ts), _ptr(tc), _ptr(tg), len(tg), int(qs.shape[1]), kmer_size,
                                    ani_percentile, ctypes.c_float(ani_adjustment), ctypes.byref(ncmp))
    return float(v), int(ncmp.value)


# ---- the two phases as one C-ABI call each (wfmash_b200/csrc/phases_host.cu) --------------------------------------------------
class _Seq(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char_p), ("seq", ctypes.c_char_p), ("len", ctypes.c_int64)]


class MapPhaseParams(ctypes.Structure):
    """wfb_map_phase_params_t with the CLI defaults (percentage_identity <= 0 = ANI auto-identity)."""
    _fields_ = [("kmer_size", ctypes.c_int32), ("sketch_size", ctypes.c_int32), ("minimum_hits", ctypes.c_int32), ("index_threads", ctypes.c_int32),
                ("window_length", ctypes.c_int64), ("percentage_identity", ctypes.c_float), ("ani_adjustment", ctypes.c_float),
                ("ani_percentile", ctypes.c_int32), ("skip_self", ctypes.c_int32), ("skip_prefix", ctypes.c_int32), ("lower_triangular", ctypes.c_int32),
                ("stage1_top_ani_filter", ctypes.c_int32), ("keep_low_pct_id", ctypes.c_int32), ("prefix_delim", ctypes.c_int32), ("reserved_", ctypes.c_int32),
                ("max_kmer_freq", ctypes.c_double), ("hg_numerator", ctypes.c_double), ("ani_diff", ctypes.c_float), ("ani_diff_conf", ctypes.c_float),
                ("filter", FilterParams)]

    def __init__(self, filter=None, **kw):  # noqa: A002
        d = dict(kmer_size=15, sketch_size=0, minimum_hits=-1, index_threads=1, window_length=1000, percentage_identity=0.0, ani_adjustment=-2.0,
                 ani_percentile=50, skip_self=1, skip_prefix=1, lower_triangular=0, stage1_top_ani_filter=1, keep_low_pct_id=1, prefix_delim=ord("#"),
                 reserved_=0, max_kmer_freq=0.0002, hg_numerator=1.0, ani_diff=0.0, ani_diff_conf=0.999)
        unknown = set(kw) - set(d)
        if unknown:
            raise TypeError(f"unknown mapping parameter(s): {sorted(unknown)}")
        d.update(kw)
        super().__init__(**d)
        self.filter = filter if filter is not None else FilterParams(window_length=d["window_length"])


class MapPhaseStats(ctypes.Structure):
    _fields_ = [("fragments", ctypes.c_int64), ("l2_mappings", ctypes.c_int64), ("mappings", ctypes.c_int64), ("sketch_size", ctypes.c_int32),
                ("minimum_hits", ctypes.c_int32), ("percentage_identity", ctypes.c_float), ("stale_absorbed", ctypes.c_int32), ("index_seconds", ctypes.c_double),
                ("map_kernel_ms", ctypes.c_double), ("filter_seconds", ctypes.c_double), ("total_seconds", ctypes.c_double), ("ani_seconds", ctypes.c_double),
                ("index_kernel_ms", ctypes.c_double), ("ani_kernel_ms", ctypes.c_double)]


class AlignPhaseParams(ctypes.Structure):
    _fields_ = [("target_padding", ctypes.c_uint64), ("query_padding", ctypes.c_uint64), ("wflign_max_len_minor", ctypes.c_uint64),
                ("batch_records", ctypes.c_int32), ("reserved_", ctypes.c_int32), ("output", _PafParams)]


class AlignPhaseStats(ctypes.Structure):
    _fields_ = [("records", ctypes.c_int64), ("written", ctypes.c_int64), ("skipped_lines", ctypes.c_int64), ("aligned_bp", ctypes.c_uint64),
                ("kernel_ms", ctypes.c_double), ("total_seconds", ctypes.c_double), ("persist_kernel_ms", ctypes.c_double), ("patch_kernel_ms", ctypes.c_double),
                ("batches", ctypes.c_int64)] + [(n, ctypes.c_uint64) for n in ("cells", "base_cells", "extend_matches", "base_extend_matches", "overlap_tests",
                                                                                "score_steps", "base_score_steps", "h2d_bytes", "d2h_bytes", "patch_cap_kept_main",
                                                                                "main_device_cap")]


def _seq_array(seqs):
    arr = (_Seq * max(len(seqs), 1))()
    keep = []
    for i, (name, seq) in enumerate(seqs):
        nb = name.encode() if isinstance(name, str) else name
        keep.append((nb, seq))
        arr[i] = _Seq(nb, seq, len(seq))
    return arr, keep


def map_phase(targets, queries, params: MapPhaseParams = None, device: int = 0, all_queries=None):
    """wfb_map_phase: the whole `wfmash -m` phase over in-memory sequences -> (mapping PAF bytes, MapPhaseStats).
    all_queries: every query of the run when `queries` is only this process's share (wfb_map_phase_subset: ids, groups, ANI estimate and
    fragment order come from all of them, so the shares' texts concatenate to the text of one call)."""
    P = params or MapPhaseParams()
    ta, tk = _seq_array(targets)
    txt, n, st = ctypes.c_void_p(), ctypes.c_int64(0), MapPhaseStats()
    L = lib()
    if all_queries is not None:
        mine = {name for name, _ in queries}
        qa, qk = _seq_array(all_queries)
        sel = (ctypes.c_uint8 * max(1, len(all_queries)))(*[1 if name in mine else 0 for name, _ in all_queries])
        rc = L.wfb_map_phase_subset(device, ctypes.byref(P), ta, len(targets), qa, len(all_queries), sel, ctypes.byref(txt), ctypes.byref(n), ctypes.byref(st))
    else:
        qa, qk = _seq_array(queries)
        rc = L.wfb_map_phase(device, ctypes.byref(P), ta, len(targets), qa, len(queries), ctypes.byref(txt), ctypes.byref(n), ctypes.byref(st))
    if rc != 0:
        raise _err(rc)
    out = ctypes.string_at(txt, n.value)
    L.wfb_free_text.argtypes = [ctypes.c_void_p]
    L.wfb_free_text(txt)
    return out, st


def align_phase(aligner: "Aligner", mapping_paf: bytes, targets, queries, window_length: int = 1000, target_padding: int = -1, query_padding: int = -1,
                batch_records: int = 0, min_identity=0.0, min_alignment_length=32, min_block_identity=0.1, disable_chain_patching=False, term_group=8,
                sam_format=False, emit_md_tag=False, no_seq_in_sam=False):
    """wfb_align_phase: mapping PAF text -> alignment PAF / SAM text (bytes, AlignPhaseStats); defaults = the CLI's."""
    pad = min(window_length, 5000)
    P = AlignPhaseParams(pad if target_padding < 0 else target_padding, pad if query_padding < 0 else query_padding, window_length * 128, batch_records, 0,
                         _PafParams(int(disable_chain_patching), term_group, min_identity, min_block_identity, min_alignment_length, int(sam_format),
                                    int(emit_md_tag), int(no_seq_in_sam), 0))
    ta, tk = _seq_array(targets)
    qa, qk = _seq_array(queries)
    txt, n, st = ctypes.c_void_p(), ctypes.c_int64(0), AlignPhaseStats()
    L = lib()
    rc = L.wfb_align_phase(ctypes.c_void_p(aligner._h), ctypes.byref(P), mapping_paf, ctypes.c_int64(len(mapping_paf)), ta, len(targets), qa, len(queries),
                           ctypes.byref(txt), ctypes.byref(n), ctypes.byref(st))
    if rc != 0:
        raise _err(rc)
    out = ctypes.string_at(txt, n.value)
    L.wfb_free_text.argtypes = [ctypes.c_void_p]
    L.wfb_free_text(txt)
    return out, st


# ---- SURVEY 8 f4 (second half): the reference's index file (-W / -I) ----------------------------------------------------
class _IndexView(ctypes.Structure):
    _fields_ = [("minmers", ctypes.c_void_p), ("n_minmers", ctypes.c_int64), ("uhash", ctypes.c_void_p), ("n_uniq", ctypes.c_int64),
                ("ustart", ctypes.c_void_p), ("ucount", ctypes.c_void_p), ("points", ctypes.c_void_p), ("n_points", ctypes.c_int64)]


class _IndexFileHeader(ctypes.Structure):
    _fields_ = [("batch_idx", ctypes.c_uint64), ("total_batches", ctypes.c_uint64), ("index_by_size", ctypes.c_int64), ("window_length", ctypes.c_int64),
                ("sketch_size", ctypes.c_int32), ("kmer_size", ctypes.c_int32), ("n_targets", ctypes.c_int32), ("n_ids", ctypes.c_int32),
                ("next_id", ctypes.c_int32), ("reserved_", ctypes.c_int32), ("target_names", ctypes.c_void_p), ("id_names", ctypes.c_void_p),
                ("id_values", ctypes.c_void_p)]


def _view_of(minmers, uhash, ustart, ucount, points):
    mi = np.ascontiguousarray(minmers, dtype=MINMER_DTYPE); uh = np.ascontiguousarray(uhash, dtype=np.uint64)
    us = np.ascontiguousarray(ustart, dtype=np.uint32); uc = np.ascontiguousarray(ucount, dtype=np.uint32); pt = np.ascontiguousarray(points, dtype=np.uint64)
    return _IndexView(mi.ctypes.data, len(mi), uh.ctypes.data, len(uh), us.ctypes.data, uc.ctypes.data, pt.ctypes.data, len(pt)), (mi, uh, us, uc, pt)


def index_file_write(path: str, exported, kmer_size: int, window_length: int, sketch_size: int, target_names, id_map, append: bool = False,
                     batch_idx: int = 0, total_batches: int = 1, index_by_size: int = 2**63 - 1):
    """Sketch::writeIndex (winSketch.hpp:616-659): one subset of a `-W` file. exported = Index.export(); id_map = {name: id}."""
    view, keep = _view_of(*exported)
    tn = ("\n".join(target_names)).encode()
    names = list(id_map)
    idn = ("\n".join(names)).encode()
    idv = np.array([id_map[n] for n in names], dtype=np.int32)
    tb, ib = ctypes.create_string_buffer(tn), ctypes.create_string_buffer(idn)
    h = _IndexFileHeader(batch_idx, total_batches, index_by_size, window_length, sketch_size, kmer_size, len(target_names), len(names),
                         (max(id_map.values()) + 1) if id_map else 0, 0, ctypes.addressof(tb), ctypes.addressof(ib), idv.ctypes.data)
    rc = lib().wfb_index_file_write(path.encode(), int(append), ctypes.byref(h), ctypes.byref(view))
    if rc != 0:
        raise _err(rc)


def index_file_read(path: str, offset: int = 0):
    """Sketch::readIndex (winSketch.hpp:840-979) for the subset at `offset` -> (header dict, (minmers, uhash, ustart, ucount, points), next offset)."""
    L = lib()
    h, v, off = _IndexFileHeader(), _IndexView(), ctypes.c_int64(offset)
    rc = L.wfb_index_file_read(path.encode(), ctypes.byref(off), ctypes.byref(h), ctypes.byref(v))
    if rc != 0:
        raise _err(rc)
    try:
        def arr(ptr, n, dt):
            return np.frombuffer(ctypes.string_at(ptr, n * np.dtype(dt).itemsize), dtype=dt).copy() if n else np.zeros(0, dtype=dt)
        names = ctypes.string_at(h.id_names).decode().split("\n")[:-1] if h.n_ids else []
        ids = arr(h.id_values, h.n_ids, np.int32)
        hdr = {"batch_idx": h.batch_idx, "total_batches": h.total_batches, "index_by_size": h.index_by_size, "window_length": h.window_length,
               "sketch_size": h.sketch_size, "kmer_size": h.kmer_size, "next_id": h.next_id,
               "target_names": ctypes.string_at(h.target_names).decode().split("\n")[:-1] if h.n_targets else [],
               "id_map": {n: int(i) for n, i in zip(names, ids)}}
        data = (arr(v.minmers, v.n_minmers, MINMER_DTYPE), arr(v.uhash, v.n_uniq, np.uint64), arr(v.ustart, v.n_uniq, np.uint32),
                arr(v.ucount, v.n_uniq, np.uint32), arr(v.points, v.n_points, np.uint64))
    finally:
        L.wfb_index_file_release(ctypes.byref(h), ctypes.byref(v))
    return hdr, data, off.value


def _index_import(cls, exported, kmer_size, window_size, sketch_size, device=0):
    """wfb_index_import: a device index from host arrays in the Index.export() layout (e.g. index_file_read's)."""
    L = lib()
    L.wfb_index_import.restype = ctypes.c_void_p
    L.wfb_index_free.argtypes = [ctypes.c_void_p]
    view, keep = _view_of(*exported)
    self = cls.__new__(cls)
    self._L = L
    self.k, self.w, self.s = kmer_size, window_size, sketch_size
    self.stats = IndexStats()
    self.stats.kept_minmers, self.stats.unique_hashes, self.stats.interval_points = len(keep[0]), len(keep[1]), len(keep[4])
    prm = _IndexParams(kmer_size, window_size, sketch_size, 1, 0.0002)
    self._h = L.wfb_index_import(device, ctypes.byref(prm), ctypes.byref(view))
    if not self._h:
        raise WfbError(L.wfb_last_error().decode())
    return self


Index.from_export = classmethod(_index_import)
