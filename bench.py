#!/usr/bin/env python3
"""bench.py — the driver's measurement contract for wfmash_b200.

A "step" = one pass of the hot path over one batch of synthetic mapping records shaped like
BASELINE.json configs[1] (LPA.subset self all-vs-all, -k15 -w1k -P50k: 861 records / 13.0 Mbp of
query, doc/performance-tuning.md:311-319; the real FASTA lives in /root/reference and does not exist
on the GPU box, so the records are seeded synthetic ones of that shape).

  python bench.py --gpus N --steps K --warmup W          # our arm (CUDA, sm_100a)
  python bench.py --impl reference --gpus N ...          # the reference's CPU biWFA on the host cores

Prints ONE JSON line (rank 0). `value` = aligned bp/s with the sequences resident in HBM (CUDA-event
time), `e2e` = the same through the host-buffer C ABI with H2D/D2H inside the timed region.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOAD = {
    "workload": "C2-shaped synthetic mapping records -> biWFA (wfmash path 2): 861 records, query 5-25 kb "
                "(mean 15 kb, 13 Mbp/step), target = query mutated at 1/2/5/10 % (sub:ins:del 8:1:1) + 1 kb flanks "
                "each side (wfmash target padding), penalties 0,5,8,2,24,1",
    "records": 861, "len_lo": 5000, "len_hi": 25000, "divergences": [0.01, 0.02, 0.05, 0.10], "pad": 1000,
}
METRIC = "aligned bp/s (biWFA base-level alignment of mapping records)"
UNIT = "bp/s"
PEN = (5, 8, 2, 24, 1)


def make_records(rank, n=None):
    from wfmash_b200 import synth
    return synth.mapping_records(n or WORKLOAD["records"], seed=1234 + 7919 * rank, len_lo=WORKLOAD["len_lo"],
                                 len_hi=WORKLOAD["len_hi"], divergences=WORKLOAD["divergences"], pad=WORKLOAD["pad"])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------------
# CPU side: the reference's own biWFA (oracle/_ref, unmodified WFA2-lib compiled in place) or, when that
# is absent, the oracle port. Test/measurement infrastructure only.
# ---------------------------------------------------------------------------------------------------
def cpu_reference_lib():
    ref = os.path.join(ROOT, "oracle", "_ref", "libwfa2ref.so")
    if os.path.exists(ref):
        return ctypes.CDLL(ref), "reference"
    orc = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(orc):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], check=True)
    return ctypes.CDLL(orc), "port"


def cpu_align_records(lib, kind, recs, threads):
    """Align recs with `threads` host threads (one aligner per record, like wfmash). Returns (bp, seconds)."""
    class Pen(ctypes.Structure):
        _fields_ = [(n, ctypes.c_int) for n in "x o1 e1 o2 e2".split()]

    def one(rec):
        p, t, _ = rec
        buf = ctypes.create_string_buffer(len(p) + len(t) + 16)
        n, sc = ctypes.c_int(), ctypes.c_int()
        if kind == "reference":
            st = lib.ref_wfa_end2end(p, len(p), t, len(t), *PEN, 3, buf, len(buf), ctypes.byref(n), ctypes.byref(sc))
        else:
            P = Pen(*PEN)
            st = lib.orc_biwfa_align(p, len(p), t, len(t), ctypes.byref(P), buf, len(buf), ctypes.byref(n), ctypes.byref(sc), None)
        return len(t) if st == 0 else 0

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        bp = sum(ex.map(one, recs))
    return bp, time.perf_counter() - t0


def est_score(rec):
    """Rough optimal-score estimate of a record: events * mean penalty + the two flank deletions."""
    p, t, d = rec
    return len(t) * d * 6.4 + 2 * (24 + WORKLOAD["pad"]) + 10.0


def run_cpu_sample(recs, target_s, threads):
    lib, kind = cpu_reference_lib()
    # calibrate the cost model c * score^2 core-seconds on the cheapest record, then take the prefix of
    # the step's records that fills about target_s seconds on `threads` cores
    probe = min(recs, key=est_score)
    bp, dt = cpu_align_records(lib, kind, [probe], 1)
    unit = dt / est_score(probe) ** 2
    budget = target_s * threads
    sample, acc = [], 0.0
    for r in recs:
        c = unit * est_score(r) ** 2
        if len(sample) >= threads and acc + c > budget:
            break
        sample.append(r)
        acc += c
    bp, dt = cpu_align_records(lib, kind, sample, threads)
    return {"value": bp / dt, "unit": UNIT, "cores": threads, "kind": kind, "sample_ms": 1e3 * dt,
            "sample": f"first {len(sample)} of {len(recs)} records of the step ({bp} query bp), {dt:.1f} s wall on {threads} threads"}


def paf_aligned_bp(lines):
    """Aligned bp as the reference counts it: sum of qEnd - qStart over the output records (computeAlignments.hpp:481,528)."""
    bp = 0
    for ln in lines:
        if ln:
            f = ln.split(b"\t", 4)
            bp += int(f[3]) - int(f[2])
    return bp


def record_path_section(al, recs, cores, with_cpu, cpu_seconds):
    """Whole-record path (row a14 / boundary b4): do_biwfa_alignment's job — main biWFA, head / tail patch alignments,
    swizzles, trimming, PAF text — through wfb_biwfa_paf_batch with HOST buffers, next to the unmodified reference
    do_biwfa_alignment (oracle/_ref/libwflignref.so) on all host threads over a bounded prefix of the same records."""
    import wfmash_b200 as wb
    rr = [dict(query_name=f"q{i}", target_name=f"t{i}", query=t_, target=p_, mashmap_estimated_identity=1.0 - d_)
          for i, (p_, t_, d_) in enumerate(recs)]
    al.biwfa_paf_batch(rr)  # warm-up with the same batch: the timed call below is the steady state (grow-only workspaces sized)
    l0 = wb.launch_count()
    t0 = time.perf_counter()
    lines, st = al.biwfa_paf_batch(rr, min_identity=0.0, min_alignment_length=32, min_block_identity=0.1)  # CLI defaults, parse_args.hpp:566-584
    dt = time.perf_counter() - t0
    bp = paf_aligned_bp(lines)
    out = {"metric": "aligned_bp_per_s_record_path", "value": bp / dt, "unit": UNIT, "records": len(rr), "lines_written": sum(1 for x in lines if x),
           "patch_cap_failures": sum(1 for x in st if x == wb.REC_PATCH_CAP), "aligned_bp": bp, "ms": 1e3 * dt,
           "main_kernel_ms": al.last_stats.kernel_ms, "gpu_launches": int(wb.launch_count() - l0),
           "paf_bytes": sum(len(x) for x in lines), "timing": "host wall clock around wfb_biwfa_paf_batch (H2D, three kernel rounds, D2H, PAF text)"}
    ref = os.path.join(ROOT, "oracle", "_ref", "libwflignref.so")
    if with_cpu and os.path.exists(ref):
        R = ctypes.CDLL(ref)
        R.ref_do_biwfa_alignment.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int,
                                             ctypes.c_char_p, ctypes.c_char_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64,
                                             ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                             ctypes.c_uint64, ctypes.c_float, ctypes.c_uint64, ctypes.c_float, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_int, ctypes.c_char_p, ctypes.c_int]

        def one(r):
            cap = len(r["query"]) + len(r["target"]) + 4096
            buf = ctypes.create_string_buffer(cap)
            n = R.ref_do_biwfa_alignment(r["query_name"].encode(), r["query"], len(r["query"]), 0, len(r["query"]), 0, r["target_name"].encode(),
                                         r["target"], len(r["target"]), 0, len(r["target"]), *PEN, 0, 0.0, 32, 0.1, 0,
                                         r["mashmap_estimated_identity"], 0, 0, 0, buf, cap)
            return buf.raw[:max(n, 0)]
        # same cost model as run_cpu_sample: prefix of the step filling ~cpu_seconds on all threads
        probe = min(range(len(recs)), key=lambda i: est_score(recs[i]))
        t1 = time.perf_counter(); one(rr[probe]); unit = (time.perf_counter() - t1) / est_score(recs[probe]) ** 2
        k, acc = 0, 0.0
        while k < len(recs) and (k < cores or acc + unit * est_score(recs[k]) ** 2 <= cpu_seconds * cores):
            acc += unit * est_score(recs[k]) ** 2
            k += 1
        t1 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=cores) as ex:
            ref_lines = list(ex.map(one, rr[:k]))
        dtc = time.perf_counter() - t1
        out["cpu_reference"] = {"value": paf_aligned_bp(ref_lines) / dtc, "unit": UNIT, "cores": cores, "kind": "reference",
                                "sample": f"first {k} of {len(rr)} records, {dtc:.1f} s wall on {cores} threads (unmodified do_biwfa_alignment)"}
        out["lines_identical_to_reference_on_sample"] = bool(all(a == b for a, b in zip(ref_lines, lines[:k])))
    return out


def cpu_sketch_build(seqs, k, w, ssz, cores, bases):
    """Sketch::Sketch (index build: thread-pool addMinmers + frequency cut-off + minmerPosLookupIndex) of the UNMODIFIED reference
    on the host cores, through oracle/ref_sketch_driver.cpp. Checker / baseline only."""
    import tempfile
    ref = os.path.join(ROOT, "oracle", "_ref", "libsketchref.so")
    if not os.path.exists(ref):
        return {"unavailable": "oracle/_ref/libsketchref.so not built"}
    R = ctypes.CDLL(ref)
    R.ref_sketch_build.restype = ctypes.c_void_p
    n = len(seqs)
    names = (ctypes.c_char_p * n)(*[f"g{i}#1#c".encode() for i in range(n)])
    with tempfile.TemporaryDirectory() as d:
        t0 = time.perf_counter()
        h = R.ref_sketch_build(os.path.join(d, "t.fa").encode(), names, (ctypes.c_char_p * n)(*seqs), (ctypes.c_int64 * n)(*[len(x) for x in seqs]), n, k,
                               ctypes.c_int64(w), ssz, cores, ctypes.c_double(0.0002), b"#")
        dt = time.perf_counter() - t0
        nm, nh, npt = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        R.ref_sketch_sizes(ctypes.c_void_p(h), ctypes.byref(nm), ctypes.byref(nh), ctypes.byref(npt))
        R.ref_sketch_free(ctypes.c_void_p(h))
    return {"value": bases / dt / 1e6, "unit": "Mbp/s", "kind": "reference", "cores": cores, "kept_minmers": int(nm.value), "unique_hashes": int(nh.value),
            "sample": f"all {n} sequences ({bases} bp) incl. writing the FASTA the reference reads, {dt:.2f} s wall"}


def map_path_section(dev, cores, with_cpu):
    """Path 1 (MashMap 3.5 sketch / index / L1) on a C3-shaped synthetic pangenome slice: 8 haplotypes x 3 Mbp at
    3 % divergence, -k15 -w1k, s=29 (scerevisiae8 parameters). Reported beside the headline."""
    import wfmash_b200 as wb
    from wfmash_b200 import synth
    rng = np.random.default_rng(77)
    root = synth.random_seq(3_000_000, rng)
    seqs = [root.tobytes()] + [synth.mutate(root, 0.03, rng).tobytes() for _ in range(7)]
    ids = list(range(8))
    k, w, ssz = 15, 1000, 29
    bases = sum(len(x) for x in seqs)
    t0 = time.perf_counter()
    ix = wb.Index(seqs, ids, k, w, ssz, index_threads=cores)
    ix.close()
    ix = wb.Index(seqs, ids, k, w, ssz, index_threads=cores)   # second build: warm context / allocator
    t_index = time.perf_counter() - t0
    st = ix.stats
    blob = b"".join(seqs)
    offs = np.cumsum([0] + [len(x) for x in seqs])
    frags, fqs = [], []
    for qi, sq in enumerate(seqs):
        for j in range(len(sq) // w):
            frags.append((int(offs[qi]) + j * w, w, qi)); fqs.append((qi, qi))
    cut = np.array([max(1, int(i * 0.6)) for i in range(1001)], dtype=np.int32)
    grp = np.arange(8, dtype=np.int32)
    r = ix.l1(blob, frags, fqs, 3, cut, grp)       # warm-up
    r = ix.l1(blob, frags, fqs, 3, cut, grp)
    rho = st.total_windows / bases
    pbar = st.interval_points / max(1, st.unique_hashes)
    idx_bytes = bases * (1 + 32 * rho + 48 * rho)                    # ~7.2 B per indexed base
    l1_bytes = len(frags) * w * (1 + 32 * ssz / w) + len(frags) * ssz * (16 + 2 * 8 * pbar)
    peak, _ = load_peaks()
    out = {
        "workload": "C3-shaped synthetic: 8 haplotypes x 3 Mbp, 3 % divergence, -k15 -w1k s=29, all-vs-all fragments",
        "index": {"bases": bases, "windows": int(st.total_windows), "kept_minmers": int(st.kept_minmers),
                  "unique_hashes": int(st.unique_hashes), "interval_points": int(st.interval_points),
                  "count_threshold": int(st.count_threshold), "stale_absorbed": int(st.minmer.stale_absorbed),
                  "stream_kernel_ms": st.minmer.stream_kernel_ms, "minmer_total_kernel_ms": st.minmer.total_kernel_ms,
                  "index_kernel_ms": st.index_kernel_ms,
                  "mbp_per_s_stream_kernel": bases / st.minmer.stream_kernel_ms / 1e3,
                  "roofline_frac_stream_kernel": idx_bytes / (st.minmer.stream_kernel_ms / 1e3) / 1e9 / peak},
        "l1": {"fragments": len(frags), "loci": int(len(r["loci"])), "kernel_ms": r["kernel_ms"],
               "fragments_per_s": len(frags) / (r["kernel_ms"] / 1e3), "query_mbp_per_s": len(frags) * w / r["kernel_ms"] / 1e3,
               "roofline_frac": l1_bytes / (r["kernel_ms"] / 1e3) / 1e9 / peak},
    }
    # L1 + L2 fused on the device: the L1 loci stay in HBM, the L2 kernel streams minmerIndex (32 B / record)
    s1 = wb.stage1_min_hits(k, ssz)
    ms = wb.l2_min_shared(0.85, k, ssz)
    frags_a = np.array(frags, dtype=wb.FRAG_DTYPE)
    fqs_a = np.array(fqs, dtype=wb.FRAG_QUERY_DTYPE)
    m = ix.map_fragments(blob, frags_a, fqs_a, 3, cut, grp, stage1_min_hits=s1, l2_min_shared=ms)   # warm-up (sizes the workspace)
    t_map = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        m = ix.map_fragments(blob, frags_a, fqs_a, 3, cut, grp, stage1_min_hits=s1, l2_min_shared=ms)
        t_map = min(t_map, time.perf_counter() - t0)
    out["l2"] = {"fragments": len(frags), "l1_loci": int(m["n_l1_loci"]), "l2_loci": int(m["l2_loci"]), "mappings": int(len(m["mappings"])),
                 "index_records_visited": int(m["l2_steps"]), "l1_kernel_ms": m["l1_kernel_ms"], "l2_kernel_ms": m["l2_kernel_ms"],
                 "sort_kernel_ms": m["sort_kernel_ms"], "loci_per_s": m["l2_loci"] / max(m["l2_kernel_ms"], 1e-9) * 1e3,
                 "roofline_frac": 32.0 * m["l2_steps"] / (max(m["l2_kernel_ms"], 1e-9) / 1e3) / 1e9 / peak,
                 "e2e_query_mbp_per_s": len(frags) * w / t_map / 1e6,
                 "e2e_note": "wfb_map_fragments_batch with host buffers: H2D of the query bases, L1 + L2 + sort kernels, D2H of the mappings"}
    if with_cpu:
        # the reference's own computeL2MappedRegions (oracle/_ref/libl2ref.so: slidingMap.hpp + mappingCore.hpp compiled
        # unmodified) on a sample of the same L1 loci, all host threads
        ref2 = os.path.join(ROOT, "oracle", "_ref", "libl2ref.so")
        if os.path.exists(ref2):
            r1 = ix.map_fragments(blob, frags[:20000], fqs[:20000], 3, cut, grp, stage1_min_hits=s1, l2_min_shared=ms, with_l1=True)
            l2lib = ctypes.CDLL(ref2)
            l2lib.ref_l2_open.restype = ctypes.c_void_p
            kept = np.ascontiguousarray(ix.export()[0])
            h = ctypes.c_void_p(l2lib.ref_l2_open(ctypes.c_void_p(kept.ctypes.data), ctypes.c_int64(len(kept))))
            L1 = r1["l1"]
            todo = [(f, j) for f in range(len(L1["count"])) for j in range(int(L1["offset"][f]), int(L1["offset"][f]) + int(L1["count"][f]))
                    if L1["loci"]["intersectionSize"][j] >= s1[int(L1["q_count"][f])]]
            l2dt = np.dtype([("seqId", "<i4"), ("shared", "<i4"), ("mean", "<i8"), ("start", "<i8"), ("end", "<i8"), ("strand", "<i4"), ("pad", "<i4")])

            def one_l2(chunk):
                o = np.zeros(64, dtype=l2dt)
                tot = 0
                for f, j in chunk:
                    lc = L1["loci"][j]
                    q = np.ascontiguousarray(L1["q_minmers"][f])
                    tot += l2lib.ref_l2_locus(h, ctypes.c_void_p(q.ctypes.data), int(L1["q_count"][f]), w, int(lc["seqId"]),
                                              ctypes.c_int64(int(lc["rangeStartPos"])), ctypes.c_int64(int(lc["rangeEndPos"])),
                                              ctypes.c_void_p(o.ctypes.data), 64)
                return tot
            nthr = max(1, min(cores, 64))
            chunks = [todo[i::nthr] for i in range(nthr)]
            t0 = time.perf_counter()
            with ThreadPoolExecutor(max_workers=nthr) as ex:
                tot = sum(ex.map(one_l2, chunks))
            dtc = time.perf_counter() - t0
            l2lib.ref_l2_close(h)
            out["l2"]["cpu_computeL2MappedRegions"] = {"value": len(todo) / dtc, "unit": "loci/s", "kind": "reference", "cores": nthr,
                                                       "sample": f"{len(todo)} L1 loci of the first 20000 fragments, {tot} L2 loci, {dtc:.2f} s wall (ctypes call overhead included)"}
    ix.close()
    if with_cpu:
        # CPU side of addMinmers on all host threads (one sequence per thread, like Sketch::build's worker pool,
        # winSketch.hpp:188-234): the unmodified reference (oracle/_ref/libmapref.so) when it travelled with the
        # snapshot, else the oracle port (exact restatement of the same state machine).
        ref = os.path.join(ROOT, "oracle", "_ref", "libmapref.so")
        orc = os.path.join(ROOT, "oracle", "liboracle.so")
        kind = "reference" if os.path.exists(ref) else ("port" if os.path.exists(orc) else None)
        if kind:
            lib = ctypes.CDLL(ref if kind == "reference" else orc)
            fn = lib.ref_add_minmers if kind == "reference" else lib.orc_add_minmers
            fn.restype = ctypes.c_int64
            dt = np.dtype([("hash", "<u8"), ("wpos", "<i8"), ("wpos_end", "<i8"), ("seqId", "<i4"), ("strand", "<i2"), ("pad_", "<i2")])

            def one(i):
                o = np.zeros(len(seqs[i]) // 4 + 1000, dtype=dt)
                buf = ctypes.create_string_buffer(seqs[i], len(seqs[i]) + 16)  # the reference upper-cases in place
                return fn(buf, ctypes.c_int64(len(seqs[i])), k, w, ssz, i, ctypes.c_void_p(o.ctypes.data), ctypes.c_int64(len(o)))
            one(0)  # warm-up (page-in, the reference's meter set-up)
            t0 = time.perf_counter()
            with ThreadPoolExecutor(max_workers=min(cores, len(seqs))) as ex:
                list(ex.map(one, range(len(seqs))))
            dtc = time.perf_counter() - t0
            out["index"]["cpu_addMinmers"] = {"value": bases / dtc / 1e6, "unit": "Mbp/s", "kind": kind, "cores": min(cores, len(seqs)),
                                              "sample": f"all {len(seqs)} sequences ({bases} bp), {dtc:.2f} s wall"}
        try:  # the reference's whole index build (unmodified skch::Sketch compiled in place, oracle/_ref/libsketchref.so)
            out["index"]["cpu_sketch_build"] = cpu_sketch_build(seqs, k, w, ssz, cores, bases)
        except Exception as e:
            out["index"]["cpu_sketch_build"] = {"error": str(e)}
    return out


def cpu_reference_pipeline(seqs, mapping_paf, cores):
    """The reference's own two phases on the host cores, UNMODIFIED (oracle/_ref/libmapperref.so = skch::Map, libalignref.so =
    align::Aligner, compiled in place): the mapping phase on the same sequences and parameters (-p 90 -k15 -w1k -P50k, defaults
    otherwise), the alignment phase on OUR mapping PAF so that both sides align the same records. Baseline / checker only."""
    import tempfile
    mlib, alib = os.path.join(ROOT, "oracle", "_ref", "libmapperref.so"), os.path.join(ROOT, "oracle", "_ref", "libalignref.so")
    if not (os.path.exists(mlib) and os.path.exists(alib)):
        return {"unavailable": "oracle/_ref/libmapperref.so / libalignref.so not built"}

    class MP(ctypes.Structure):
        _fields_ = [(n, ctypes.c_int32) for n in "kmer_size sketch_size threads filter_mode skip_self skip_prefix lower_triangular merge_mappings split minimum_hits".split()] + \
                   [(n, ctypes.c_int64) for n in "window_length block_length chain_gap scaffold_gap scaffold_max_deviation scaffold_min_length".split()] + \
                   [("max_mapping_length", ctypes.c_uint64), ("num_mappings_for_segment", ctypes.c_uint32), ("num_mappings_for_scaffold", ctypes.c_uint32),
                    ("percentage_identity", ctypes.c_float), ("prefix_delim", ctypes.c_int32), ("overlap_threshold", ctypes.c_double),
                    ("scaffold_overlap_threshold", ctypes.c_double), ("max_kmer_freq", ctypes.c_double)]
    import wfmash_b200 as wb
    s = wb.sketch_size(0.90, 1000, 15)
    prm = MP(15, s, cores, 1, 1, 1, 0, 1, 1, -1, 1000, 0, 2000, 100000, 100000, 10000, 50000, 2**32 - 1, 1, 0.90, ord("#"), 0.95, 0.5, 0.0002)
    n = len(seqs)
    names = (ctypes.c_char_p * n)(*[a.encode() for a, _ in seqs]); sq = (ctypes.c_char_p * n)(*[b for _, b in seqs]); ln = (ctypes.c_int64 * n)(*[len(b) for _, b in seqs])
    M, A = ctypes.CDLL(mlib), ctypes.CDLL(alib)
    M.ref_map_phase.restype = ctypes.c_int64
    A.ref_align_phase.restype = ctypes.c_int64
    buf = ctypes.create_string_buffer(max(64 << 20, 8 * len(mapping_paf)))
    with tempfile.TemporaryDirectory() as d:
        t0 = time.perf_counter()
        k = M.ref_map_phase(d.encode(), ctypes.byref(prm), names, sq, ln, n, names, sq, ln, n, 1, buf, ctypes.c_int64(len(buf)))
        t_map = time.perf_counter() - t0
    ref_map = buf.raw[:max(k, 0)]
    strip = lambda t: sorted(b"\t".join(x.split(b"\t")[:14]) for x in t.split(b"\n") if x)   # ch:Z: is schedule dependent in the reference itself
    A.ref_align_set_threads(cores)
    with tempfile.TemporaryDirectory() as d:
        t0 = time.perf_counter()
        k2 = A.ref_align_phase(d.encode(), names, sq, ln, n, names, sq, ln, n, mapping_paf, ctypes.c_int64(len(mapping_paf)), ctypes.c_uint64(1000), ctypes.c_uint64(1000),
                               ctypes.c_uint64(128000), ctypes.c_float(0.0), ctypes.c_uint64(32), ctypes.c_float(0.1), 0, 0, 0, 0, buf, ctypes.c_int64(len(buf)))
        t_aln = time.perf_counter() - t0
    A.ref_align_set_threads(1)
    ref_paf = buf.raw[:max(k2, 0)]
    aligned = 0
    for row in mapping_paf.split(b"\n"):
        if row:
            f = row.split(b"\t")
            aligned += int(f[3]) - int(f[2])
    return {"kind": "reference", "cores": cores, "seconds": {"map_phase": t_map, "align_phase": t_aln, "total": t_map + t_aln},
            "aligned_bp_per_s": aligned / (t_map + t_aln) if t_map + t_aln > 0 else None, "mapping_rows": ref_map.count(b"\n"),
            "mapping_columns_1_14_identical": strip(ref_map) == strip(mapping_paf), "paf_lines": ref_paf.count(b"\n"), "_paf_sorted": sorted(ref_paf.split(b"\n")),
            "sample": "the whole workload of this section (reference mapper on the same sequences; reference aligner on our mapping PAF), FASTA writing included"}


# ---------------------------------------------------------------------------------------------------
# BASELINE.json's metric is quoted end to end (map + WFA). pipeline_section runs the two phases chained
# (wfmash_b200.pipeline.wfmash: index build -> L1/L2 kernels -> host chain merge + filters -> mapping PAF ->
# padded records -> biWFA kernels + patches -> alignment PAF) on a C4-shaped synthetic pair of haplotypes,
# host buffers in, PAF text out, wall clock around the whole call.
# ---------------------------------------------------------------------------------------------------
def pipeline_section(dev, contigs=16, contig_bp=500_000, ani=0.95, runs=2, with_cpu=True, cores=1):
    from wfmash_b200 import synth
    rng = np.random.default_rng(4242)
    d = 1.0 - ani ** 0.5  # each haplotype derived from the root at d = 1 - sqrt(ANI)
    seqs = []
    for c in range(contigs):
        root = synth.random_seq(contig_bp, rng)
        seqs.append((f"gA#1#chr{c + 1:02d}", synth.mutate(root, d, rng).tobytes()))
        seqs.append((f"gB#1#chr{c + 1:02d}", synth.mutate(root, d, rng).tobytes()))
    import wfmash_b200 as wb
    MP = wb.MapPhaseParams(percentage_identity=0.90)
    al = wb.Aligner(dev)
    best, st, n_launch = None, None, 0
    for i in range(runs + 1):  # first run = warm-up (workspaces, page-in)
        l0 = wb_launches()
        t0 = time.perf_counter()
        mp, ms = wb.map_phase(seqs, seqs, MP, dev)                 # one C-ABI call: ids, index, fragments, L1/L2, chain + filters, mapping PAF
        t1 = time.perf_counter()
        paf, a = wb.align_phase(al, mp, seqs, seqs)                # one C-ABI call: rows -> padded records -> biWFA + patches -> PAF text
        t2 = time.perf_counter()
        if i and (best is None or t2 - t0 < best[0]):
            best, st, n_launch = (t2 - t0, t1 - t0, t2 - t1), (ms, a, len(paf)), wb_launches() - l0
    al.close()
    ms, a, paf_bytes = st
    total_bp = sum(len(x) for _, x in seqs)
    cpu_ref = None
    if with_cpu:
        try:
            cpu_ref = cpu_reference_pipeline(seqs, mp, cores)
            if "_paf_sorted" in cpu_ref:
                cpu_ref["paf_lines_identical"] = cpu_ref.pop("_paf_sorted") == sorted(paf.split(b"\n"))
                cpu_ref["aligned_bp_per_s"] = int(a.aligned_bp) / cpu_ref["seconds"]["total"]   # same numerator as ours (padded record spans)
        except Exception as e:
            cpu_ref = {"error": str(e)}
    return {"cpu_reference": cpu_ref, "workload": f"C4-shaped synthetic: 2 haplotypes x {contigs} contigs x {contig_bp} bp at {ani:.0%} ANI, all-vs-all, -p 90 -k15 -w1k -P50k (defaults otherwise)",
            "sequence_bp": total_bp, "mapping_records": int(ms.mappings), "fragments": int(ms.fragments), "records_aligned": int(a.records),
            "paf_lines": int(a.written), "paf_bytes": paf_bytes, "aligned_bp": int(a.aligned_bp),
            "seconds": {"total": best[0], "map_phase": best[1], "align_phase": best[2], "index_build": ms.index_seconds, "map_kernels": ms.map_kernel_ms / 1e3,
                        "chain_filter_paf": ms.filter_seconds, "align_kernels": a.kernel_ms / 1e3},
            "aligned_bp_per_s": int(a.aligned_bp) / best[0], "mapped_bp_per_s": total_bp / best[1], "gpu_launches": int(n_launch),
            "note": "wfb_map_phase + wfb_align_phase (C ABI), host sequences in, PAF text out; wall clock around the two calls; best of %d runs after one warm-up" % runs}


def wb_launches():
    import wfmash_b200 as wb
    return wb.launch_count()


# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--records", type=int, default=0, help="override records per step (debug)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (tuning sweeps)")
    ap.add_argument("--no-map", action="store_true", help="skip the mapping-path (path 1) section")
    ap.add_argument("--no-record", action="store_true", help="skip the whole-record (PAF) section")
    ap.add_argument("--no-pipeline", action="store_true", help="skip the chained map + align section")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return 0
        recs = make_records(0, args.records or None)
        vals, ms, last = [], [], None
        for i in range(args.warmup + args.steps):
            last = run_cpu_sample(recs, max(2.0, args.cpu_seconds / max(1, args.steps)), cores)
            if i >= args.warmup:
                vals.append(last["value"])
                ms.append(last["sample_ms"])
        v = statistics.mean(vals)
        last["value"] = v
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": statistics.mean(ms), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "int32", "data": "synthetic", "config": dict(WORKLOAD), "cpu_baseline": last,
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import torch
    import wfmash_b200 as wb
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank if world > 1 else 0
    torch.cuda.set_device(dev)
    assert wb.device_count() > dev, "no CUDA device: wfmash_b200 has no CPU path"
    recs = make_records(rank, args.records or None)
    pairs = [(p, t) for p, t, _ in recs]
    n = len(pairs)
    bp_step = sum(len(t) for _, t in pairs)
    L = wb.lib()
    al = wb.Aligner(dev)
    # device-resident copy of the inputs for the `value` leg
    blob = b"".join(p + t for p, t in pairs)
    d_blob = L.wfb_device_malloc(dev, len(blob) + 64)
    assert d_blob and L.wfb_memcpy_h2d(dev, d_blob, blob, len(blob)) == 0
    poff = np.zeros(n, dtype=np.int64); toff = np.zeros(n, dtype=np.int64)
    plen = np.zeros(n, dtype=np.int32); tlen = np.zeros(n, dtype=np.int32)
    o = 0
    for i, (p, t) in enumerate(pairs):
        poff[i] = o; plen[i] = len(p); o += len(p)
        toff[i] = o; tlen[i] = len(t); o += len(t)
    cap = int(plen.sum() + tlen.sum()) + 16
    ops = ctypes.create_string_buffer(cap)
    res = (wb._Res * n)()
    arr = (wb._Pair * n)(*[wb._Pair(p, len(p), t, len(t)) for p, t in pairs])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{dev}")  # > 126 MB L2

    def step_device(stats):
        rc = L.wfb_align_batch_device(al._h, d_blob, poff.ctypes.data, plen.ctypes.data, toff.ctypes.data, tlen.ctypes.data,
                                      n, ops, cap, res, ctypes.byref(stats))
        assert rc == 0, L.wfb_last_error()

    def step_host(stats):
        rc = L.wfb_align_batch(al._h, arr, n, ops, cap, res, ctypes.byref(stats))
        assert rc == 0, L.wfb_last_error()

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        st = wb.AlignStats()
        for _ in range(warmup):
            fn(st)
        wall, dev_ms, brk_ms, last = [], [], [], None
        launches0 = wb.launch_count()
        barrier()
        for _ in range(steps):
            flush.fill_(1)  # flush L2 between timed iterations (outside the per-step timers)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn(st)
            torch.cuda.synchronize()
            wall.append(time.perf_counter() - t0)
            dev_ms.append(st.kernel_ms); brk_ms.append(st.break_kernel_ms)
            last = st.as_dict()
        barrier()
        return wall, dev_ms, brk_ms, last, wb.launch_count() - launches0

    sampler = ClockSampler(dev)
    sampler.start()
    wall_d, devms_d, brk_d, stats_d, launches_d = timed(step_device, args.steps, args.warmup)
    wall_h, devms_h, brk_h, stats_h, launches_h = timed(step_host, args.steps, max(1, min(args.warmup, 1)))
    clocks = sampler.stop()
    ok = sum(1 for r in res if r.status == 0)
    aligned_bp = sum(int(tlen[i]) for i in range(n) if res[i].status == 0)

    t_dev = sum(devms_d) / 1e3          # CUDA-event seconds for K steps
    t_host = sum(wall_h)                # wall seconds for K steps, H2D + D2H inside
    if world > 1:
        import torch.distributed as dist
        tt = torch.tensor([t_dev, t_host], dtype=torch.float64, device=f"cuda:{dev}")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_host = float(tt[0]), float(tt[1])
        tb = torch.tensor([aligned_bp], dtype=torch.float64, device=f"cuda:{dev}")
        dist.all_reduce(tb, op=dist.ReduceOp.SUM)
        total_bp = float(tb[0])
    else:
        total_bp = float(aligned_bp)
    value = total_bp * args.steps / t_dev
    e2e = total_bp * args.steps / t_host

    if rank == 0:
        peak, peak_src = load_peaks()
        # roofline of the dominant kernel (wfb_persist_kernel: the whole biWFA recursion tree, breakpoint + base-case
        # tasks, in one launch): algorithmic bytes of SURVEY §8d = 48 B/cell + 2 B/extended base + 8 B/overlap test
        # over the cells / extends / overlap tests of BOTH task kinds, per step, over its CUDA-event time
        alg_bytes = (48 * (stats_d["cells"] + stats_d["base_cells"]) + 2 * (stats_d["extend_matches"] + stats_d["base_extend_matches"])
                     + 8 * stats_d["overlap_tests"])
        brk_s = (sum(brk_d) / len(brk_d)) / 1e3
        n_brk_launch = max(1, int(stats_d["levels"]))
        achieved = alg_bytes / brk_s / 1e9 if brk_s > 0 else 0.0
        traffic, traffic_src = None, None
        try:  # DRAM bytes of ONE launch of this kernel from the committed ncu --set full capture of this workload
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            if tj.get("records") == n and world == 1:
                traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
        except Exception:
            pass
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "traffic_source": traffic_src,
                    "kernel": "wfb_persist_kernel", "peak_source": peak_src,
                    "algorithmic_bytes_per_step": alg_bytes, "kernel_ms_per_step": brk_s * 1e3, "launches_per_step": n_brk_launch,
                    "cells_per_step": stats_d["cells"] + stats_d["base_cells"],
                    "gcells_per_s": (stats_d["cells"] + stats_d["base_cells"]) / brk_s / 1e9 if brk_s > 0 else 0.0}
        cpu = run_cpu_sample(recs, args.cpu_seconds, cores) if (world == 1 and not args.no_cpu) else None
        h2d = int(plen.sum() + tlen.sum()) + 64 * n
        d2h = sum(r.ops_len for r in res) + 24 * n
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": dict(WORKLOAD, l2="flushed between timed iterations (256 MiB write); per-step workspace >> L2",
                           records_per_gpu=n, query_bp_per_gpu=bp_step, completed=ok,
                           timing="value: CUDA events on the library stream; e2e: host wall clock around the C-ABI call",
                           stats=stats_d),
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * t_host / args.steps},
            "gpu_launches": int(launches_d),
            "roofline": roofline,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if world == 1 and not args.no_record:
            try:
                line["record_path"] = record_path_section(al, recs, cores, not args.no_cpu, args.cpu_seconds)
            except Exception as e:
                line["record_path"] = {"error": str(e)}
        if world == 1 and not args.no_map:
            try:
                line["map_path"] = map_path_section(dev, cores, not args.no_cpu)
            except Exception as e:  # the headline must still be printed
                line["map_path"] = {"error": str(e)}
        if world == 1 and not args.no_pipeline:
            try:
                line["pipeline"] = pipeline_section(dev, with_cpu=not args.no_cpu, cores=cores)
            except Exception as e:
                line["pipeline"] = {"error": str(e)}
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
