"""Reference-side windowed minmers (mm_stream_kernel + stitch + post passes behind wfb_minmers_build) WITHOUT a
GPU: the kernel bodies run under the single-thread host emulation of tests/emu (TEST INFRASTRUCTURE, -DWFB_EMU, in a
subprocess; never the product library) and are compared with the oracle's exact addMinmers restatement (pinned to the
compiled reference in tests/test_map_oracle_cpu.py) on the cases the GPU parity test uses: random, tandem repeats, N runs,
lower case, short sequences, several (k, w, s), plus the reference's LPA example file. Guards the kernel's control flow (rolling
k-mer registers, chunk warm-up, stitching) between GPU runs. Three builds of the same result: the candidate-filtered stream (the default:
mm_cand_kernel + mm_stream_cand_kernel, with the exact re-run of the chunks that flag themselves: N runs, low-complexity sequence), the
same with a candidate capacity so small that every tile overflows (all chunks take the re-run path), and the unfiltered stream."""
import json
import os
import shutil
import subprocess
import sys

import pytest

from tests import util

SCRIPT = r"""
import json, sys
sys.path.insert(0, %(root)r)
import numpy as np
import wfmash_b200 as wb
from tests import datasets, maputil, util
oracle = util.load_oracle()
bad, n = [], 0
redo = 0
cases = maputil.minmer_cases()
if %(lpa)r:
    cases.append(("LPA", b"NN".join(x for _, x in datasets.load("lpa")), 15, 1000, 24))
for name, sq, k, w, s in cases:
    got, st = wb.minmers_build([sq, sq[: len(sq) // 2]], [7, 9], k, w, s)
    exp = np.concatenate([maputil.orc_add_minmers(oracle, x, k, w, s, sid) for x, sid in ((sq, 7), (sq[: len(sq) // 2], 9)) if len(x) >= w])
    ok = len(got) == len(exp) and all((got[f] == exp[f]).all() for f in ("hash", "wpos", "wpos_end", "seqId", "strand")) and st.stitch_miss == 0
    n += len(exp)
    redo += st.redo_chunks
    if not ok or st.filtered != %(filtered)d:
        bad.append(name)
print(json.dumps({"bad": bad, "records": n, "redo": int(redo)}))
"""


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ for the host emulation build")
@pytest.mark.parametrize("mode", ["filtered", "filtered-tiles-overflow", "unfiltered"])
def test_minmer_stream_kernel_body_under_emulation_matches_oracle(mode):
    so = subprocess.run([os.path.join(util.ROOT, "tests", "emu", "build_emu.sh")], check=True, capture_output=True, text=True).stdout.strip().splitlines()[-1]
    env = dict(os.environ, WFB_LIB=os.path.join(util.ROOT, so))
    env.pop("WFB_MM_FILTER", None)
    env.pop("WFB_MM_CAND_CAP", None)
    if mode == "unfiltered":
        env["WFB_MM_FILTER"] = "0"
    if mode == "filtered-tiles-overflow":
        env["WFB_MM_CAND_CAP"] = "150"
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": util.ROOT, "lpa": mode == "filtered", "filtered": int(mode != "unfiltered")}], env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert res["bad"] == [] and res["records"] > 20000
    # the filtered build re-runs only the chunks that must (N runs, ACGT repeats, two-letter sequence); with overflowing tiles, all of them
    assert (res["redo"] == 0) == (mode == "unfiltered")
    if mode == "filtered":
        assert res["records"] > 120000 and res["redo"] < 200


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ for the host emulation build")
def test_minmer_build_differential_fuzz_under_emulation():
    """600 random cases (tests/minmer_fuzz.py: parameters, boundary lengths, degenerate sequences) of the default (candidate-filtered) build against
    the oracle; 63 000 cases / 61 M records of the same generator ran clean when the filtered build was written."""
    so = subprocess.run([os.path.join(util.ROOT, "tests", "emu", "build_emu.sh")], check=True, capture_output=True, text=True).stdout.strip().splitlines()[-1]
    env = dict(os.environ, WFB_LIB=os.path.join(util.ROOT, so))
    for k in ("WFB_MM_FILTER", "WFB_MM_CAND_CAP", "WFB_MM_LCUR", "WFB_MM_FCHUNK"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, os.path.join(util.ROOT, "tests", "minmer_fuzz.py"), "11", "120", "600"], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert res["mismatches"] == 0 and res["cases"] == 600 and res["filtered_builds"] > 400 and res["redo_chunks"] > 100, res


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ for the host emulation build")
def test_fragment_sketch_differential_fuzz_under_emulation():
    """tests/sketch_fuzz.py: the query-fragment sketch kernel body against the oracle on random parameters and degenerate sequences
    (535 000 fragments ran clean at the end of round 2; 3 000 here)."""
    so = subprocess.run([os.path.join(util.ROOT, "tests", "emu", "build_emu.sh")], check=True, capture_output=True, text=True).stdout.strip().splitlines()[-1]
    env = dict(os.environ, WFB_LIB=os.path.join(util.ROOT, so))
    r = subprocess.run([sys.executable, os.path.join(util.ROOT, "tests", "sketch_fuzz.py"), "4", "120", "3000"], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert res["mismatches"] == 0 and res["fragments"] >= 3000, res
