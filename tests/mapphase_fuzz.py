"""Differential fuzz of the mapping phase's host half (wfb_map_phase under the host emulation build, WFB_LIB: ids, PanSN groups, fragments, run-level
constants, chain merge + filters, mapping PAF; the fragment mappings are injected from the oracle through the emulation build's test hook) against the
reference's UNMODIFIED skch::Map (oracle/_ref/libmapperref.so): random sequence sets incl. targets of exactly / barely one window, other name
delimiters, twelve CLI options. This generator found the minmer tie-order issue. The one-to-one filter mode is left out: its
kc:f: column has a documented deviation. TEST INFRASTRUCTURE. python tests/mapphase_fuzz.py SEED SECONDS [MAX_CASES]"""
import ctypes
import json
import os
import random
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wfmash_b200 as wb
from wfmash_b200 import pipeline
from tests import pipeutil, util
oracle, M = util.load_oracle(), util.load_ref("libmapperref.so")
if M is None:
    print(json.dumps({"cases": 0, "rows": 0, "mismatches": 0, "skipped": "oracle/_ref/libmapperref.so not built"}))
    sys.exit(0)
MAX_CASES = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 30
rnd = random.Random(int(sys.argv[1]))
T_END = time.time() + float(sys.argv[2])
n = bad = rows = 0
null = os.open(os.devnull, os.O_WRONLY); saved = os.dup(2)
while time.time() < T_END and bad < 3 and n < MAX_CASES:
    seqs = pipeutil.case(seed=rnd.randrange(1, 10**6), length=rnd.choice([8_000, 14_000, 24_000, 31_000]))
    if rnd.random() < 0.3:
        seqs = seqs + [("d#1#chrZ", seqs[1][1][: rnd.choice([999, 1000, 1001, 2500])])]
    if rnd.random() < 0.2:
        seqs = [(n_.replace("#", rnd.choice(["#", "_"])), s_) for n_, s_ in seqs]
    prm = {}
    f = {}
    no_tag = False
    def maybe(d, k, vals, p=0.3):
        if rnd.random() < p: d[k] = rnd.choice(vals)
    maybe(prm, "lower_triangular", [True], 0.15)
    if rnd.random() < 0.15: prm["skip_self"] = False; prm["skip_prefix"] = False
    maybe(prm, "kmer_size", [15, 17, 19], 0.2)
    maybe(prm, "window_length", [500, 1000], 0.3)
    maybe(prm, "percentage_identity", [0.70, 0.80, 0.90, 0.95], 0.6)
    maybe(f, "num_mappings_for_segment", [1, 2, 3])
    maybe(f, "overlap_threshold", [0.5, 0.95, 1.0])
    maybe(f, "block_length", [0, 3000, 5000])
    maybe(f, "chain_gap", [500, 2000, 20000])
    maybe(f, "max_mapping_length", [5000, 10000, 50000])
    maybe(f, "scaffold_min_length", [0, 500, 2000, 5000])
    maybe(f, "split", [0, 1], 0.2)
    maybe(f, "merge_mappings", [0, 1], 0.15)
    if f: prm["filter"] = f
    try:
        P = pipeutil.params(prm)
        R = P.resolved()
    except Exception as e:
        continue
    ids = pipeline.SequenceIds(seqs, seqs, R.prefix_delim if R.skip_prefix else "")
    fake = pipeutil.OracleIndex(oracle, [s for _, s in seqs], [ids.id_of[x] for x, _ in seqs], ids.group, R.kmer_size, R.window_length, R.sketch_size,
                                R.max_kmer_freq, R.index_threads)
    min_hits = max(R.minimum_hits, wb.estimate_minimum_hits_relaxed(R.sketch_size, R.kmer_size, R.percentage_identity))
    W = R.window_length
    r = fake.map_fragments(None, [0] * sum(len(s) // W + (1 if len(s) % W and len(s) >= W else 0) for _, s in seqs), None,
                           min_hits, wb.sketch_cutoffs(R.sketch_size, R.kmer_size), None, skip_self=R.skip_self, skip_prefix=R.skip_prefix,
                           lower_triangular=R.lower_triangular, stage1_min_hits=wb.stage1_min_hits(R.kmer_size, R.sketch_size),
                           l2_min_shared=wb.l2_min_shared_relaxed(R.percentage_identity, R.kmer_size, R.sketch_size))
    maps, off = np.ascontiguousarray(r["mappings"]), np.ascontiguousarray(r["offset"], dtype=np.int64)
    wb.lib().wfb_emu_inject_l2(ctypes.c_void_p(maps.ctypes.data), ctypes.c_void_p(off.ctypes.data), ctypes.c_int64(len(off) - 1))
    MP = wb.MapPhaseParams(filter=R.filter, kmer_size=R.kmer_size, window_length=R.window_length, percentage_identity=R.percentage_identity,
                           skip_self=int(R.skip_self), skip_prefix=int(R.skip_prefix), lower_triangular=int(R.lower_triangular))
    ours, st = wb.map_phase(seqs, seqs, MP)
    os.dup2(null, 2)
    ref = pipeutil.reference_map_phase(M, seqs, P)
    os.dup2(saved, 2)
    cut = (lambda t: sorted(b"\t".join(x.split(b"\t")[:14]) for x in t.split(b"\n") if x)) if no_tag else (lambda t: sorted(x for x in t.split(b"\n") if x))
    n += 1; rows += len(cut(ref))
    if cut(ours) != cut(ref) or st.sketch_size != R.sketch_size or st.minimum_hits != min_hits:
        bad += 1
        a, b = set(cut(ours)), set(cut(ref))
        print("MISMATCH", prm, [(x, len(y)) for x, y in seqs], len(a), len(b), "\n only ours:", sorted(a - b)[:2], "\n only ref:", sorted(b - a)[:2], flush=True)
print(json.dumps({"cases": n, "rows": rows, "mismatches": bad}))
