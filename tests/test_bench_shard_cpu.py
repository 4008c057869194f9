"""bench.py's multi-GPU partitioning (one fixed job over the ranks: queries by length for the mapping phase, mapping rows by expected
cost for the alignment phase), checked on the CPU on config C3sub: the ranks' shares, computed one after the other through the C++
phases under the emulation build (fragment mappings injected from the oracle; alignment kernels emulated), re-assemble to exactly
the text the unmodified reference wrote (tests/golden/config_reference.json.gz). The NCCL exchange itself is plain byte all-gathers."""
import json
import os
import shutil
import subprocess
import sys

import pytest

from tests import util

SCRIPT = r"""
import ctypes, json, sys
sys.path.insert(0, %(root)r)
import numpy as np
import bench
import wfmash_b200 as wb
from wfmash_b200 import pipeline
from tests import configs, configrun, pipeutil, util
from tests.test_configs_cpu import sha_sorted
oracle, doc = util.load_oracle(), configrun.golden()
name, world = "C3sub", 3
cfg = configs.by_name(name)
t, q = configs.sequences(cfg)
prm = dict(cfg["params"]); prm["percentage_identity"] = doc[name]["percentage_identity"]
R = pipeutil.params(prm).resolved()
w = R.window_length
ids = pipeline.SequenceIds(t, q, R.prefix_delim)
fake = pipeutil.OracleIndex(oracle, [s for _, s in t], [ids.id_of[x] for x, _ in t], ids.group, R.kmer_size, w, R.sketch_size, R.max_kmer_freq, 1)
min_hits = max(R.minimum_hits, wb.estimate_minimum_hits_relaxed(R.sketch_size, R.kmer_size, R.percentage_identity))
nfr = [len(s) // w + (1 if len(s) %% w else 0) if len(s) >= w else 0 for _, s in q]
r = fake.map_fragments(None, [0] * sum(nfr), None, min_hits, wb.sketch_cutoffs(R.sketch_size, R.kmer_size), None,
                       stage1_min_hits=wb.stage1_min_hits(R.kmer_size, R.sketch_size), l2_min_shared=wb.l2_min_shared_relaxed(R.percentage_identity, R.kmer_size, R.sketch_size))
maps, off = r["mappings"], r["offset"]
first = np.cumsum([0] + nfr)
owner = bench.partition_queries(q, w, world)
MP = wb.MapPhaseParams(filter=R.filter, kmer_size=R.kmer_size, window_length=w, percentage_identity=R.percentage_identity, sketch_size=R.sketch_size)
parts = []
for rank in range(world):
    mine = [(n, s) for n, s in q if owner[n] == rank]
    # the fragment mappings of this rank's queries, renumbered to the rank's own fragment list
    sel, noff = [], [0]
    for qi, (n, s) in enumerate(q):
        if owner[n] != rank:
            continue
        for f in range(first[qi], first[qi + 1]):
            m = maps[off[f]: off[f + 1]].copy()
            m["frag"] = len(noff) - 1
            sel.append(m); noff.append(noff[-1] + len(m))
    sub = np.ascontiguousarray(np.concatenate(sel)) if sel else np.zeros(1, dtype=wb.L2_MAPPING_DTYPE)
    noff = np.array(noff, dtype=np.int64)
    wb.lib().wfb_emu_inject_l2(ctypes.c_void_p(sub.ctypes.data), ctypes.c_void_p(noff.ctypes.data), ctypes.c_int64(len(noff) - 1))
    txt, st = wb.map_phase(t, mine, MP, all_queries=q)
    parts.append(txt)
per_query = {}
for part in parts:
    for ln in part.split(b"\n"):
        if ln:
            per_query.setdefault(ln.split(b"\t", 1)[0], []).append(ln)
rows = [ln for n, _ in q for ln in per_query.get(n.encode(), [])]
out = {"mapping_rows": len(rows), "mapping_sha": sha_sorted(b"".join(x + b"\n" for x in rows)), "ranks_with_queries": len({owner[n] for n, _ in q})}
row_owner = bench.partition_rows(rows, world)
al = wb.Aligner(0)
paf = []
for rank in range(world):
    share = b"".join(rows[i] + b"\n" for i in range(len(rows)) if row_owner[i] == rank)
    text, ast = wb.align_phase(al, share, t, t, window_length=w)
    paf.append(text)
out.update(paf_lines=sum(x.count(b"\n") for x in paf), paf_sha=sha_sorted(b"".join(paf)), rows_per_rank=[row_owner.count(r) for r in range(world)])
print(json.dumps(out))
"""


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ for the host emulation build")
def test_partitioned_job_reassembles_to_the_reference_text():
    from tests import configrun
    so = subprocess.run([os.path.join(util.ROOT, "tests", "emu", "build_emu.sh")], check=True, capture_output=True, text=True).stdout.strip().splitlines()[-1]
    env = dict(os.environ, WFB_LIB=os.path.join(util.ROOT, so))
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": util.ROOT}], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    g = configrun.golden()["C3sub"]
    assert res["ranks_with_queries"] == 3 and min(res["rows_per_rank"]) > 0
    assert res["mapping_rows"] == g["mapping_rows"] and res["mapping_sha"] == g["mapping_sha_sorted"]
    assert res["paf_lines"] == g["alignment_lines"] and res["paf_sha"] == g["alignment_sha_sorted"]
