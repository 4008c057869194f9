#!/usr/bin/env python3
"""bench.py — the driver's measurement contract for wfmash_b200.

Metric (BASELINE.json): aligned bp/s, END TO END (map + WFA). Workload = BASELINE.json configs[2], the config the target is quoted on:
`wfmash data/scerevisiae8.fa.gz -Y '#'` — 8 yeast genomes, 136 sequences, 96 255 507 bp, all-vs-all, CLI defaults (ANI
auto-identity, -k15 -w1k -P50k): 21 129 mapping records / 672 177 587 aligned bp. The input file ships under tests/data/ (byte
copy of /root/reference/data/scerevisiae8.fa.gz); it fits one GPU. A "step" = one pass of both hot paths over it:

    wfb_map_phase   (ANI estimate, index build, L1 + L2 kernels, chain merge + filters, mapping PAF)
    wfb_align_phase (mapping rows -> padded records -> biWFA + head / tail patches -> alignment PAF)

host sequences in, PAF text out, exactly what `wfmash target.fa` does. "Aligned bp" is the reference's own counter
(computeAlignments.hpp:481,528: qEnd - qStart of every processed record).

  python bench.py --gpus N --steps K --warmup W          # our arm (CUDA, sm_100a)
  python bench.py --impl reference --gpus N ...          # the reference's UNMODIFIED skch::Map + align::Aligner on the host cores

Prints ONE JSON line (rank 0).
  value  = aligned bp / device time of the step: the sum of the CUDA-event times of every kernel round of the step (index build, L1 / L2,
           ANI, biWFA, patches) — the throughput with the inputs resident in HBM, host work excluded;
  e2e    = aligned bp / host wall clock around the two C-ABI calls (host buffers in, H2D / D2H and all host work inside): the headline;
  N > 1  : ONE fixed job partitioned over the ranks (strong scaling): queries by length for the mapping phase, mapping rows by expected
           cost for the alignment phase, rows / PAF bytes exchanged as NCCL all-gathers of byte tensors; rank 0 checks that the text
           equals the single-GPU text.
"""
import argparse
import ctypes
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "aligned bp/s (end-to-end map+WFA)"
UNIT = "bp/s"
CONFIG_NAME = "C3"
WORKLOAD = {  # the SAME dict in both arms' lines (the driver compares them); what a run measured goes under "run"
    "workload": "scerevisiae8 all-vs-all -Y '#' (BASELINE.json configs[2]): data/scerevisiae8.fa.gz, 8 genomes, 136 sequences, 96 255 507 bp; "
                "CLI defaults (-p ani50-2 -k15 -w1k -P50k); mapping phase + alignment phase, host sequences in, PAF text out",
    "config": CONFIG_NAME, "input": "tests/data/scerevisiae8.fa.gz (byte copy of the reference's data file)",
    "l2": "flushed between timed iterations (256 MiB write); the step streams GBs of wavefronts",
    "timing": "value: sum of the CUDA-event times of the step's kernel rounds, max over ranks; e2e: host wall clock around the C-ABI calls "
              "(+ the NCCL exchanges at N > 1), max over ranks; reference arm: host wall clock of the reference's own two phases",
}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "500"],
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


def load_config(name):
    from tests import configrun, configs
    cfg = configs.by_name(name)
    targets, queries = configs.sequences(cfg)
    return cfg, targets, queries, configrun.golden()[name]


def sha_sorted(text: bytes) -> str:
    return hashlib.sha256(b"\n".join(sorted(ln for ln in text.split(b"\n") if ln))).hexdigest()


# ---------------------------------------------------------------------------------------------------
# CPU side: the reference's own two phases, UNMODIFIED (oracle/_ref*/libmapperref.so = skch::Map, libalignref.so = align::Aligner,
# compiled in place from /root/reference by oracle/Makefile; prebuilt files travel to the GPU box). Baseline / checker only.
# ---------------------------------------------------------------------------------------------------
def reference_libs():
    """The timing build the host CPU can run: the reference's Release flags (-Ofast -funroll-all-loops) with its AVX-512 kernels
    (x86-64-v4) when /proc/cpuinfo has them, its AVX2 kernels (x86-64-v3) otherwise; the -O3 parity build as the last resort."""
    flags = set()
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("flags"):
                flags = set(ln.split(":", 1)[1].split())
                break
    except OSError:
        pass
    order = []
    if {"avx512f", "avx512bw", "avx512vl", "avx512dq", "avx512cd"} <= flags:
        order.append(("_ref_fast_v4", "-Ofast -funroll-all-loops -march=x86-64-v4 (AVX-512 kernels)"))
    order += [("_ref_fast_v3", "-Ofast -funroll-all-loops -march=x86-64-v3 (AVX2 kernels)"), ("_ref", "-O3 -march=x86-64-v3 (AVX2 kernels; parity build)")]
    for d, desc in order:
        m, a = os.path.join(ROOT, "oracle", d, "libmapperref.so"), os.path.join(ROOT, "oracle", d, "libalignref.so")
        if os.path.exists(m) and os.path.exists(a):
            return ctypes.CDLL(m), ctypes.CDLL(a), desc
    return None, None, None


class _Quiet:
    """The reference logs progress to stderr from C++; keep the bench's stderr readable."""

    def __enter__(self):
        sys.stderr.flush()
        self.saved = os.dup(2)
        null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(null, 2)
        os.close(null)

    def __exit__(self, *a):
        os.dup2(self.saved, 2)
        os.close(self.saved)


def reference_step(cfg, targets, queries, gold, cores, align_seconds, seed):
    """One bounded sample of the workload on the host cores: the reference's WHOLE mapping phase (all threads) + its alignment phase
    over a seeded random sample of the mapping rows sized for about `align_seconds` of wall time; the sample's alignment time is
    extrapolated to all rows by aligned bp. -> dict."""
    from tests import pipeutil
    M, A, build = reference_libs()
    if M is None:
        return {"unavailable": "oracle/_ref*/libmapperref.so + libalignref.so not built (make -C oracle ref fast)"}
    prm = dict(cfg["params"])
    prm["percentage_identity"] = gold["percentage_identity"]   # the ANI estimate is not part of the timed path on either side
    P = pipeutil.params(prm)
    same = queries is targets
    t0 = time.perf_counter()
    with _Quiet():
        mp = pipeutil.reference_map_phase(M, targets, P, threads=cores, queries=None if same else queries)
    t_map = time.perf_counter() - t0
    rows = [ln for ln in mp.split(b"\n") if ln]
    import wfmash_b200 as wb
    w = P.resolved().window_length
    span = []
    for ln in rows:
        row, _, _ = wb.mapping_paf_parse(ln, min(w, 5000), min(w, 5000), w * 128)
        span.append(row.q_end - row.q_start)
    total_bp = int(sum(span))
    # sample size from the reference's own speed in the build container (tests/golden/config_reference.json.gz: seconds on 8 cores),
    # scaled by the core count; at least 64 rows
    ref_s = gold.get("reference_seconds", {}).get("align", 0.0) * 8.0 / max(1, cores)
    frac = 1.0 if ref_s <= 0 else min(1.0, max(64.0 / max(1, len(rows)), align_seconds / ref_s))
    rng = np.random.default_rng(seed)
    pick = np.sort(rng.choice(len(rows), size=max(1, int(round(frac * len(rows)))), replace=False))
    sample = b"".join(rows[i] + b"\n" for i in pick)
    sample_bp = int(sum(span[i] for i in pick))
    A.ref_align_set_threads(cores)
    t0 = time.perf_counter()
    with _Quiet():
        out = pipeutil.reference_align_phase(A, sample, targets if same else targets + queries, P)
    t_aln = time.perf_counter() - t0
    A.ref_align_set_threads(1)
    t_aln_full = t_aln * total_bp / max(1, sample_bp)
    return {"value": total_bp / (t_map + t_aln_full), "unit": UNIT, "cores": cores, "kind": "reference", "build": build,
            "sample_ms": 1e3 * (t_map + t_aln),
            "seconds": {"map_phase_all_rows": t_map, "align_phase_sample": t_aln, "align_phase_extrapolated": t_aln_full},
            "mapping_rows": len(rows),
            # the timing build's -Ofast changes the last digit of some float tags (id:f, kc:f, mapq) in the reference itself; the coordinates
            # (columns 1-9) must equal the fixture's, which was written by the IEEE-exact -O3 build
            "mapping_coordinates_equal_fixture": sorted(b"\t".join(ln.split(b"\t")[:9]) for ln in rows) ==
                                                 sorted("\t".join(d["head"].split("\t")[:9]).encode() for d in gold["mapping"]),
            "sample": f"whole mapping phase on {cores} threads ({t_map:.1f} s) + alignment phase of a seeded random sample of {len(pick)} of {len(rows)} "
                      f"mapping rows ({sample_bp} of {total_bp} aligned bp, {t_aln:.1f} s on {cores} threads, {out.count(b'\n')} PAF lines), alignment time "
                      f"extrapolated by aligned bp; unmodified skch::Map + align::Aligner, FASTA writing of the driver included"}


# ---------------------------------------------------------------------------------------------------
# Our arm
# ---------------------------------------------------------------------------------------------------
from wfmash_b200.shard import allgather_bytes, job_sharded, partition_queries, partition_rows  # noqa: E402,F401  (the library's sharding of one job)


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line of the contract, written to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # Libraries print to stdout behind our back (NCCL's "NCCL version ..." banner is a printf): keep the original stdout for the JSON line
    # and point fd 1 at stderr for everything else.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=CONFIG_NAME, help="tests/configs.py name (debug: C2, C3sub)")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="wall-time budget of the cpu_baseline alignment sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (tuning)")
    ap.add_argument("--batch-records", type=int, default=0)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    cfg, targets, queries, gold = load_config(args.config)
    workload = dict(WORKLOAD, config=args.config)
    data_desc = ("reference's own example data (data/scerevisiae8.fa.gz)" if args.config.startswith("C3") else
                 "synthetic (wfmash_b200/synth.py: xoshiro256** seed 42, SURVEY 8d's model)" if str(cfg["target"]).startswith("synth") else "reference's own example data")
    if args.config != CONFIG_NAME:  # a parity-test config run as a bench (tests/configs.py): say what it is, not what the default is
        workload["workload"] = (f"tests/configs.py {args.config}: {len(targets)} target / {len(queries)} query sequences, {sum(len(x) for _, x in targets)} target bp, "
                                f"parameters {cfg['params']}; mapping phase + alignment phase, host sequences in, PAF text out")
        workload["input"] = f"tests/configs.py target={cfg['target']} query={cfg['query']}"

    if args.impl == "reference":
        if rank != 0:
            return 0
        total = args.warmup + args.steps
        per_step = min(30.0, max(4.0, 150.0 / max(1, total)))   # the whole run stays within a few minutes
        vals, ms, last = [], [], None
        for i in range(total):
            last = reference_step(cfg, targets, queries, gold, cores, per_step, seed=1000 + i)
            if "unavailable" in last:
                emit({"impl": "reference", "unavailable": last["unavailable"]})
                return 0
            if i >= args.warmup:
                vals.append(last["value"]); ms.append(last["sample_ms"])
        v = statistics.mean(vals)
        last["value"] = v
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": statistics.mean(ms), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32",
                "data": data_desc, "config": workload, "cpu_baseline": last,
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        emit(line)
        return 0

    import torch
    import wfmash_b200 as wb
    from tests import configrun
    if world > 1:
        import torch.distributed as dist
        # (NCCL's version banner is a printf to stdout: see emit() — the process's fd 1 already points at stderr)
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank if world > 1 else 0
    tdev = torch.device("cuda", dev)
    torch.cuda.set_device(dev)
    assert wb.device_count() > dev, "no CUDA device: wfmash_b200 has no CPU path"
    MP, w = configrun.phase_params(wb, cfg)
    same = queries is targets
    al = wb.Aligner(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=tdev)  # > 126 MB L2
    q_owner = partition_queries(queries, w, world)
    my_queries = [(n, s) for n, s in queries if q_owner[n] == rank] if world > 1 else queries

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        """-> dict of this rank's timings + the whole job's text (wfmash_b200.shard.job_sharded: the C++ phases, rows / PAF exchanged as
        all-gathers of byte tensors at N > 1)."""
        r = job_sharded(wb, al, targets, queries, MP, w, dev, tdev, rank, world, batch_records=args.batch_records, my_queries=my_queries if world > 1 else None)
        mst, ast = r["mst"], r["ast"]
        # device time of the step: every kernel round's CUDA-event time
        r["device_s"] = (mst.index_kernel_ms + mst.map_kernel_ms + mst.ani_kernel_ms + ast.kernel_ms + ast.patch_kernel_ms) / 1e3
        return r

    for _ in range(args.warmup):
        last = step()
    sampler = ClockSampler(dev)
    sampler.start()
    launches0 = wb.launch_count()
    steps = []
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        flush.fill_(1)  # flush L2 between timed iterations (the step's working set is >> L2 anyway)
        torch.cuda.synchronize()
        steps.append(step())
    ev1.record()
    barrier()
    t_wall = time.perf_counter() - t_begin
    launches = wb.launch_count() - launches0
    clocks = sampler.stop()
    last = steps[-1]
    t_host = sum(s["t_total"] for s in steps)          # wall seconds of K steps (host buffers in, text out)
    t_dev = sum(s["device_s"] for s in steps)          # CUDA-event seconds of K steps' kernel rounds
    aligned_bp = float(last["ast"].aligned_bp)
    if world > 1:
        import torch.distributed as dist
        tt = torch.tensor([t_host, t_dev], dtype=torch.float64, device=tdev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_host, t_dev = float(tt[0]), float(tt[1])
        tb = torch.tensor([aligned_bp, float(last["my_records"])], dtype=torch.float64, device=tdev)
        per_rank = [torch.zeros(2, dtype=torch.float64, device=tdev) for _ in range(world)]
        dist.all_gather(per_rank, tb)
        aligned_bp = float(sum(x[0] for x in per_rank))
        records_per_gpu = [int(x[1]) for x in per_rank]
        tl = torch.tensor([launches], dtype=torch.float64, device=tdev)
        dist.all_reduce(tl, op=dist.ReduceOp.SUM)
        launches = int(tl[0])
    else:
        records_per_gpu = [last["my_records"]]
    value = aligned_bp * args.steps / t_dev
    e2e = aligned_bp * args.steps / t_host

    if rank == 0:
        peak, peak_src = load_peaks()
        ast, mst = last["ast"], last["mst"]
        # roofline of the dominant kernel (wfb_persist_kernel: the whole biWFA recursion tree of a batch — breakpoint + base-case tasks —
        # in one launch): algorithmic bytes of SURVEY 8(d) = 48 B / cell + 2 B / extended base + 8 B / overlap test over the counts of
        # this rank's launches, over their CUDA-event time
        alg_bytes = 48 * (ast.cells + ast.base_cells) + 2 * (ast.extend_matches + ast.base_extend_matches) + 8 * ast.overlap_tests
        k_s = ast.persist_kernel_ms / 1e3
        achieved = alg_bytes / k_s / 1e9 if k_s > 0 else 0.0
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            if tj.get("config") == args.config and world == 1:
                traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
        except Exception:
            pass
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                    "kernel": "wfb_persist_kernel", "peak_source": peak_src, "algorithmic_bytes_per_step": alg_bytes, "kernel_ms_per_step": k_s * 1e3,
                    "launches_per_step": int(ast.batches), "share_of_device_time": k_s / last["device_s"] if last["device_s"] > 0 else None,
                    "cells_per_step": int(ast.cells + ast.base_cells), "extend_matches_per_step": int(ast.extend_matches + ast.base_extend_matches),
                    "overlap_tests_per_step": int(ast.overlap_tests), "score_steps_per_step": int(ast.score_steps + ast.base_score_steps),
                    "gcells_per_s": (ast.cells + ast.base_cells) / k_s / 1e9 if k_s > 0 else 0.0}
        total_seq = sum(len(s) for _, s in targets) + (0 if same else sum(len(s) for _, s in queries))
        parity = {"mapping_lines": last["mapping_paf"].count(b"\n"), "paf_lines": last["paf"].count(b"\n"),
                  "mapping_lines_identical": sha_sorted(last["mapping_paf"]) == gold["mapping_sha_sorted"],
                  "paf_lines_identical": sha_sorted(last["paf"]) == gold["alignment_sha_sorted"],
                  "aligned_bp_equals_reference_counter": int(aligned_bp) == gold["aligned_bp"], "stale_absorbed": int(mst.stale_absorbed),
                  "patch_cap_kept_main": int(ast.patch_cap_kept_main), "main_device_cap": int(ast.main_device_cap),
                  "against": "tests/golden/config_reference.json.gz: the unmodified reference's two phases on the same file (sorted lines, sha256)"}
        mean = lambda k: statistics.mean(s.get(k, 0.0) for s in steps)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32",
            "data": data_desc,
            "config": workload,
            "run": dict(sequence_bp=total_seq, mapping_records=parity["mapping_lines"], aligned_bp=int(aligned_bp), records_per_gpu=records_per_gpu,
                        percentage_identity=float(mst.percentage_identity), sketch_size=int(mst.sketch_size)),
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(total_seq + ast.h2d_bytes), "d2h_bytes_per_step": int(ast.d2h_bytes + len(last["mapping_paf"])),
                    "ms_per_step": 1e3 * t_host / args.steps, "wall_ms_of_timed_region_per_step": 1e3 * t_wall / args.steps},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "parity": parity,
            "phases_ms": {"map_phase": 1e3 * mean("t_map"), "align_phase": 1e3 * mean("t_align"), "exchange_rows": 1e3 * mean("t_exchange_rows"),
                          "gather_paf": 1e3 * mean("t_gather_paf"), "gather_bytes_per_step": int(last["gather_bytes"]),
                          "index_build_s": mst.index_seconds, "ani_s": mst.ani_seconds, "map_kernels_ms": mst.map_kernel_ms, "chain_filter_paf_s": mst.filter_seconds,
                          "align_main_kernels_ms": ast.kernel_ms, "align_persist_kernel_ms": ast.persist_kernel_ms, "align_patch_kernels_ms": ast.patch_kernel_ms,
                          "align_batches": int(ast.batches)},
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = reference_step(cfg, targets, queries, gold, cores, args.cpu_seconds, seed=1)
        emit(line)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
