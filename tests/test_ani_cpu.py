"""SURVEY 8 f3 — ANI auto-identity (wfb_ani_group_sketches + wfb_ani_estimate_identity) WITHOUT a GPU: the hashing kernel's
body runs under the single-thread host emulation of tests/emu (TEST INFRASTRUCTURE, -DWFB_EMU, in a subprocess; never the
product library). Checked against (1) the oracle's literal StreamingMinHash restatement, group sketch by group sketch,
(2) the identity the reference's UNMODIFIED Stat::estimate_identity_for_groups returns (oracle/_ref/libstatsref.so: map_stats.hpp
compiled in place over FASTA / GSL stand-ins) — live and through the committed fixture tests/golden/ani_reference.json.gz.
The same checks run on the real kernel in tests/test_gpu_parity.py::test_ani_auto_identity_matches_reference."""
import gzip
import json
import os
import shutil
import subprocess
import sys

import pytest

from tests import util

# (name, query slice, target slice, percentile, adjustment) over tests.aniutil.case()
SETS = [("all_vs_all", (0, 6), (0, 6), 50, -2.0), ("split_sets", (0, 3), (2, 6), 25, 0.0), ("p75_plus1", (1, 6), (0, 5), 75, 1.0)]

SCRIPT = r"""
import hashlib, json, sys
sys.path.insert(0, %(root)r)
import numpy as np
import wfmash_b200 as wb
from wfmash_b200 import pipeline
from tests import aniutil, util
from tests.test_ani_cpu import SETS
seqs = aniutil.case()
oracle = util.load_oracle()
ids = pipeline.SequenceIds(seqs, seqs, "#")
gids = sorted(set(ids.group)); dense = {g: i for i, g in enumerate(gids)}
grp = [dense[ids.group[ids.id_of[n]]] for n, _ in seqs]
out = {"sketch_equal": [], "passes": [], "ident": {}, "ref": {}}
for s in (4096, 64, 16):   # 4096: the CLI value (threshold path); 64 / 16: the satellite array overflows the candidate capacity
    sk, cnt, st = wb.ani_group_sketches([x for _, x in seqs], grp, len(gids), 21, s)
    osk, ocnt = aniutil.oracle_group_sketches(oracle, [x for _, x in seqs], grp, len(gids), 21, s)
    out["sketch_equal"].append(bool((cnt == ocnt).all() and (sk == osk).all()))
    out["passes"].append(int(st.passes))
    if s == 4096:
        out["sha"] = hashlib.sha256(sk.tobytes() + cnt.tobytes()).hexdigest()
        out["valid_kmers"] = int(st.valid_kmers); out["candidates"] = int(st.candidates)
R = util.load_ref("libstatsref.so")
for name, (q0, q1), (t0, t1), pct, adj in SETS:
    P = pipeline.Params(percentage_identity=None, ani_percentile=pct, ani_adjustment=adj)
    out["ident"][name] = float(pipeline.estimate_identity(seqs[t0:t1], seqs[q0:q1], P)[0]).hex()
    if R is not None:
        out["ref"][name] = float(aniutil.reference_identity(R, seqs[t0:t1], seqs[q0:q1], "#", pct, adj)).hex()
P2 = pipeline.auto_identity(seqs, seqs, pipeline.Params(percentage_identity=None))
out["auto"] = [float(P2.percentage_identity).hex(), P2.sketch_size]
print(json.dumps(out))
"""


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ for the host emulation build")
def test_ani_kernel_body_and_identity_under_emulation():
    so = subprocess.run([os.path.join(util.ROOT, "tests", "emu", "build_emu.sh")], check=True, capture_output=True, text=True).stdout.strip().splitlines()[-1]
    env = dict(os.environ, WFB_LIB=os.path.join(util.ROOT, so))
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": util.ROOT}], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    with gzip.open(os.path.join(util.GOLD, "ani_reference.json.gz"), "rt") as f:
        gold = json.load(f)
    assert res["sketch_equal"] == [True, True, True]
    assert res["passes"][0] == 1 and max(res["passes"][1:]) >= 2      # the capacity-overflow re-run is exercised
    assert res["candidates"] < 0.6 * res["valid_kmers"]                # ... and so is the threshold filter (the one group above 4 x 8 x 4096 k-mers)
    assert res["sha"] == gold["sketch_sha"]
    assert res["ident"] == gold["identity"]                            # bit patterns of the doubles the unmodified reference returned
    if res["ref"]:
        assert res["ref"] == res["ident"]                              # live
    ident = float.fromhex(res["auto"][0])
    assert 0.85 < ident < 0.99 and res["auto"][1] == int(0.02 * (1 + (1 - ident) / 0.1) * (1000 - 15))


@pytest.mark.ref
@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ for the host emulation build")
def test_ani_identity_differential_fuzz_against_the_unmodified_reference():
    """tests/ani_fuzz.py: random sequence sets / slices / percentiles / adjustments; the estimated identity must be the double the unmodified
    estimate_identity_for_groups returns, bit for bit (9 600 cases ran clean at the end of round 2; 60 here)."""
    if util.load_ref("libstatsref.so") is None:
        pytest.skip("oracle/_ref not built (reference sources absent)")
    so = subprocess.run([os.path.join(util.ROOT, "tests", "emu", "build_emu.sh")], check=True, capture_output=True, text=True).stdout.strip().splitlines()[-1]
    env = dict(os.environ, WFB_LIB=os.path.join(util.ROOT, so))
    r = subprocess.run([sys.executable, os.path.join(util.ROOT, "tests", "ani_fuzz.py"), "6", "200", "60"], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert res["mismatches"] == 0 and res["cases"] == 60, res
