"""CPU check of the a14 host logic (wfmash_b200/csrc/epilogue.cu: erosion, junction merge, swizzles, trimming, PAF
metrics). The alignments themselves need the GPU; here the kernel bodies run under the single-thread host emulation
of tests/emu (TEST INFRASTRUCTURE, -DWFB_EMU) in a subprocess, so that the record-level driver can be compared with
the committed reference PAF fixture without a device. The product library is never replaced by this build."""
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
import wfmash_b200 as wb
from tests import util
recs = util.paf_records()
gold = util.paf_golden()
al = wb.Aligner(0, penalties=tuple(gold["penalties"]))
bad = []
for kw, want in zip(gold["filter_sets"], gold["lines"]):
    lines, status = al.biwfa_paf_batch(recs, term_group=gold["term_group"], **kw)
    bad += [(i, str(kw)) for i, (g, w) in enumerate(zip(lines, want)) if g.decode() != w]
# the reference's two patch rounds (head, then tail on the head-patched CIGAR) instead of the fused round: same lines
import os
os.environ["WFB_PATCH_FUSE"] = "0"
unfused, _ = al.biwfa_paf_batch(recs, term_group=gold["term_group"], **gold["filter_sets"][0])
del os.environ["WFB_PATCH_FUSE"]
fused, fst = al.biwfa_paf_batch(recs, term_group=gold["term_group"], **gold["filter_sets"][0])
fuse_bad = [i for i, (a, b) in enumerate(zip(fused, unfused)) if a != b]
# term_group -1 = what a -march=native build of the reference does on this host (cpuid): equals the explicit value
import re
flags = set(re.findall(r"\w+", open("/proc/cpuinfo").read().split("flags", 1)[1].split("\n", 1)[0])) if os.path.exists("/proc/cpuinfo") else set()
host_tg = 16 if {"avx512cd", "avx512vl"} <= flags else 8 if "avx2" in flags else 1
auto, _ = al.biwfa_paf_batch(recs, term_group=-1, **gold["filter_sets"][0])
expl, _ = al.biwfa_paf_batch(recs, term_group=host_tg, **gold["filter_sets"][0])
tg_bad = [i for i, (a, b) in enumerate(zip(auto, expl)) if a != b]
caps = [int(al.last_stats.patch_cap_kept_main), int(al.last_stats.main_device_cap)]
# SURVEY 8 f4: the SAM branch (write_alignment_sam, MD tag) against the committed reference fixture
import hashlib
sgold = util.sam_golden()
sbad = []
for kw, want in zip(sgold["sets"], sgold["lines"]):
    lines, status = al.biwfa_paf_batch(recs, term_group=sgold["term_group"], sam_format=True, **kw)
    sbad += [(i, str(kw), g[:80].decode()) for i, (g, w) in enumerate(zip(lines, want)) if hashlib.sha256(g).hexdigest() != w["sha"]]
# the kernel bodies themselves (4-diagonal groups, batched extend, overlap scan, base case) on the reference's own
# known-answer vectors and on medium random pairs against the oracle
pairs = util.golden_pairs()
gold_alg = util.golden_alg("wfa_utest.biwfa.affine2p.alg.gz")
ag = wb.Aligner(0, penalties=util.GOLDEN_PEN)
gbad = [i for i, (r, (gs, gc)) in enumerate(zip(ag.align_end2end_batch(pairs), gold_alg))
        if r.status != 0 or util.rle(r.ops) != gc or r.score != int(gs)]
orc = util.load_oracle()
rp = [pt for pt in util.random_pairs(60, seed=41, lengths=(400, 1500, 4000), rates=(0.01, 0.05, 0.15)) if pt[0] and pt[1]]
rbad = [i for i, ((p_, t_), r) in enumerate(zip(rp, al.align_end2end_batch(rp))) if (r.status, r.ops) != util.orc_biwfa(orc, p_, t_, util.WFMASH_PEN)[:2]]
print(json.dumps({"bad": bad, "sam_bad": sbad, "fuse_bad": fuse_bad, "tg_bad": tg_bad, "host_tg": host_tg, "caps": caps, "n": len(recs), "golden_bad": gbad, "golden_n": len(pairs), "random_bad": rbad, "random_n": len(rp)}))
"""


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ for the host emulation build")
def test_record_epilogue_and_kernel_bodies_under_emulation():
    so = subprocess.run([os.path.join(util.ROOT, "tests", "emu", "build_emu.sh")], check=True, capture_output=True, text=True).stdout.strip().splitlines()[-1]
    env = dict(os.environ, WFB_LIB=os.path.join(util.ROOT, so))
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": util.ROOT}], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert res["n"] == 63 and res["bad"] == []
    assert res["sam_bad"] == []
    assert res["fuse_bad"] == [] and res["tg_bad"] == [] and res["host_tg"] in (1, 8, 16)
    assert res["caps"] == [0, 0]
    assert res["golden_n"] == 305 and res["golden_bad"] == []
    assert res["random_n"] > 40 and res["random_bad"] == []
