"""Reference-side windowed minmers (mm_stream_kernel + stitch + post passes behind wfb_minmers_build) WITHOUT a
GPU: the kernel bodies run under the single-thread host emulation of tests/emu (TEST INFRASTRUCTURE, -DWFB_EMU, in a
subprocess; never the product library) and are compared with the oracle's exact addMinmers restatement (pinned to the
compiled reference in tests/test_map_oracle_cpu.py) on the cases the GPU parity test uses: random, tandem repeats, N runs,
lower case, short sequences, several (k, w, s). Guards the kernel's control flow (rolling k-mer registers, chunk warm-up,
stitching) between GPU runs."""
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
from tests import maputil, util
oracle = util.load_oracle()
bad, n = [], 0
for name, sq, k, w, s in maputil.minmer_cases():
    got, st = wb.minmers_build([sq, sq[: len(sq) // 2]], [7, 9], k, w, s)
    exp = np.concatenate([maputil.orc_add_minmers(oracle, x, k, w, s, sid) for x, sid in ((sq, 7), (sq[: len(sq) // 2], 9)) if len(x) >= w])
    ok = len(got) == len(exp) and all((got[f] == exp[f]).all() for f in ("hash", "wpos", "wpos_end", "seqId", "strand")) and st.stitch_miss == 0
    n += len(exp)
    if not ok:
        bad.append(name)
print(json.dumps({"bad": bad, "records": n}))
"""


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ for the host emulation build")
def test_minmer_stream_kernel_body_under_emulation_matches_oracle():
    so = subprocess.run([os.path.join(util.ROOT, "tests", "emu", "build_emu.sh")], check=True, capture_output=True, text=True).stdout.strip().splitlines()[-1]
    env = dict(os.environ, WFB_LIB=os.path.join(util.ROOT, so))
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": util.ROOT}], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert res["bad"] == [] and res["records"] > 20000
