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


class AlignStats(ctypes.Structure):
    _fields_ = [("cells", ctypes.c_uint64), ("extend_matches", ctypes.c_uint64), ("overlap_tests", ctypes.c_uint64),
                ("score_steps", ctypes.c_uint64), ("break_tasks", ctypes.c_uint64), ("base_tasks", ctypes.c_uint64),
                ("base_cells", ctypes.c_uint64), ("base_extend_matches", ctypes.c_uint64), ("base_score_steps", ctypes.c_uint64),
                ("levels", ctypes.c_uint64), ("kernel_ms", ctypes.c_double), ("break_kernel_ms", ctypes.c_double)]

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


class MinmerStats(ctypes.Structure):
    _fields_ = [("stream_kernel_ms", ctypes.c_double), ("total_kernel_ms", ctypes.c_double), ("bases", ctypes.c_uint64),
                ("raw_records", ctypes.c_uint64), ("chunks", ctypes.c_uint64), ("stale_absorbed", ctypes.c_uint64),
                ("stitch_miss", ctypes.c_uint64)]

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
