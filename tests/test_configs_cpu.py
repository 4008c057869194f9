"""BASELINE.json's configs on the reference's own input files (tests/data/, byte copies of /root/reference/data/) — the CPU half:

  * the shipped files are intact (sizes equal their .fai);
  * the HOST half of the mapping phase (ids, PanSN groups, ANI-adopted identity -> sketch size, fragments, the one-thread executor
    order that decides the ch:Z: tags, chain merge + filters, PAF text) — Python `pipeline.map` and the C++ `wfb_map_phase` under the
    emulation build — reproduces, byte for byte, what the reference's UNMODIFIED skch::Map wrote for these inputs
    (tests/golden/config_reference.json.gz, made by tests/golden/make_config_golden.py), with the device calls answered by the oracle's
    mapping restatement (which the GPU kernels are compared with in tests/test_gpu_parity.py, and on these same inputs in
    tests/test_gpu_configs.py).
No /root/reference needed at run time."""
import gzip
import hashlib
import json
import os
import shutil
import subprocess
import sys

import pytest

from tests import configs, datasets, pipeutil, util


def golden():
    with gzip.open(os.path.join(util.GOLD, "config_reference.json.gz"), "rt") as f:
        return json.load(f)


def sha_sorted(text: bytes) -> str:
    return hashlib.sha256(b"\n".join(sorted(ln for ln in text.split(b"\n") if ln))).hexdigest()


def test_shipped_inputs_are_intact():
    for key, fn in datasets.FILES.items():
        seqs = datasets.load(key)
        with open(datasets.path(key) + ".fai") as f:
            fai = [ln.split("\t") for ln in f if ln.strip()]
        assert [n for n, _ in seqs] == [r[0] for r in fai], key
        assert [len(s) for _, s in seqs] == [int(r[1]) for r in fai], key
    y = datasets.load("yeast")
    assert len(y) == 136 and sum(len(s) for _, s in y) == 96_255_507   # doc/performance-tuning.md:339
    assert sum(len(s) for _, s in datasets.load("lpa")) == 2_317_910    # doc/performance-tuning.md:295


def test_one_thread_executor_order():
    from wfmash_b200 import pipeline
    # few fragments: the last emplaced fragment runs first
    assert pipeline.one_thread_fragment_order([3, 0, 2]) == [[2, 1, 0], [], [1, 0]]
    # query 1 (popped first) has 253 free slots: fragments 252..0, then its overflow 253.. from the free list — but only after query 0
    # (still in the worker's queue) has run nested inside query 1's join
    o = pipeline.one_thread_fragment_order([2, 300])
    assert o[0] == [1, 0] and o[1] == list(range(253, -1, -1)) + list(range(254, 300))
    o = pipeline.one_thread_fragment_order([400] * 300)   # more queries than queue slots: everything is still run exactly once
    assert all(sorted(x) == list(range(400)) for x in o)


def _host_half(name, oracle, doc):
    from wfmash_b200 import pipeline
    cfg = configs.by_name(name)
    t, q = configs.sequences(cfg)
    prm = dict(cfg["params"])
    if prm["percentage_identity"] is None:
        prm["percentage_identity"] = doc[name]["percentage_identity"]
    P = pipeutil.params(prm)
    R = P.resolved()
    ids = pipeline.SequenceIds(t, q, R.prefix_delim if R.skip_prefix else "")
    fake = pipeutil.OracleIndex(oracle, [s for _, s in t], [ids.id_of[n] for n, _ in t], ids.group, R.kmer_size, R.window_length, R.sketch_size,
                                R.max_kmer_freq, R.index_threads, queries=None if q is t else [(s, ids.id_of[n]) for n, s in q])
    return pipeline.map(t, q, P, index=fake).paf


@pytest.mark.parametrize("name", ["C1", "C1w250", "C2", "C3sub"])
def test_mapping_host_half_on_the_real_inputs_equals_the_reference_text(name, oracle):
    doc = golden()
    ours = _host_half(name, oracle, doc)
    assert ours.count(b"\n") == doc[name]["mapping_rows"]
    assert sha_sorted(ours) == doc[name]["mapping_sha_sorted"], name   # whole lines, ch:Z: tags included


C_PHASE = r"""
import ctypes, json, sys
sys.path.insert(0, %(root)r)
import numpy as np
import wfmash_b200 as wb
from wfmash_b200 import pipeline
from tests import configs, pipeutil, util
from tests.test_configs_cpu import golden, sha_sorted
oracle, doc, out = util.load_oracle(), golden(), {}
for name in ("C2", "C3sub"):
    cfg = configs.by_name(name)
    t, q = configs.sequences(cfg)
    prm = dict(cfg["params"]); prm["percentage_identity"] = doc[name]["percentage_identity"]
    R = pipeutil.params(prm).resolved()
    ids = pipeline.SequenceIds(t, q, R.prefix_delim)
    fake = pipeutil.OracleIndex(oracle, [s for _, s in t], [ids.id_of[x] for x, _ in t], ids.group, R.kmer_size, R.window_length, R.sketch_size, R.max_kmer_freq, 1)
    w = R.window_length
    min_hits = max(R.minimum_hits, wb.estimate_minimum_hits_relaxed(R.sketch_size, R.kmer_size, R.percentage_identity))
    r = fake.map_fragments(None, [0] * sum(len(s) // w + (1 if len(s) %% w else 0) for _, s in t if len(s) >= w), None, min_hits,
                           wb.sketch_cutoffs(R.sketch_size, R.kmer_size), None, stage1_min_hits=wb.stage1_min_hits(R.kmer_size, R.sketch_size),
                           l2_min_shared=wb.l2_min_shared_relaxed(R.percentage_identity, R.kmer_size, R.sketch_size))
    maps, off = np.ascontiguousarray(r["mappings"]), np.ascontiguousarray(r["offset"], dtype=np.int64)
    wb.lib().wfb_emu_inject_l2(ctypes.c_void_p(maps.ctypes.data), ctypes.c_void_p(off.ctypes.data), ctypes.c_int64(len(off) - 1))
    MP = wb.MapPhaseParams(filter=R.filter, kmer_size=R.kmer_size, window_length=w, percentage_identity=R.percentage_identity, sketch_size=R.sketch_size)
    ours, st = wb.map_phase(t, q, MP)
    out[name] = {"rows": ours.count(b"\n"), "sha": sha_sorted(ours)}
# C1w250: fragments exist but none maps anywhere (the reads do not come from this reference): the phase must return empty text, not an error
t, q = configs.sequences(configs.by_name("C1w250"))
nf = sum(len(s) // 250 + (1 if len(s) %% 250 else 0) for _, s in q if len(s) >= 250)
maps, off = np.zeros(1, dtype=wb.L2_MAPPING_DTYPE), np.zeros(nf + 1, dtype=np.int64)
wb.lib().wfb_emu_inject_l2(ctypes.c_void_p(maps.ctypes.data), ctypes.c_void_p(off.ctypes.data), ctypes.c_int64(nf))
ours, st = wb.map_phase(t, q, wb.MapPhaseParams(window_length=250, filter=wb.FilterParams(window_length=250)))
out["C1w250"] = {"rows": ours.count(b"\n"), "sha": sha_sorted(ours), "fragments": int(st.fragments), "identity": float(st.percentage_identity)}
print(json.dumps(out))
"""


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ for the host emulation build")
def test_c_map_phase_host_half_on_the_real_inputs_equals_the_reference_text():
    so = subprocess.run([os.path.join(util.ROOT, "tests", "emu", "build_emu.sh")], check=True, capture_output=True, text=True).stdout.strip().splitlines()[-1]
    env = dict(os.environ, WFB_LIB=os.path.join(util.ROOT, so))
    r = subprocess.run([sys.executable, "-c", C_PHASE % {"root": util.ROOT}], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    doc = golden()
    for name in ("C2", "C3sub", "C1w250"):
        assert res[name]["rows"] == doc[name]["mapping_rows"] and res[name]["sha"] == doc[name]["mapping_sha_sorted"], name
    assert res["C1w250"]["fragments"] == 16 and abs(res["C1w250"]["identity"] - doc["C1w250"]["percentage_identity"]) < 1e-7
