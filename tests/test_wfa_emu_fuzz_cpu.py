"""The biWFA kernel bodies under the single-thread host emulation (tests/emu, TEST INFRASTRUCTURE) against the oracle on adversarial pairs
(tests/wfa_fuzz.py), and the whole record path (patching, swizzles, trimming, PAF text) against the reference's UNMODIFIED do_biwfa_alignment on
records with adversarial ends (tests/paf_fuzz.py). 25 000 pairs / 29 000 records of the same generators ran clean at the end of round 2; the
suite runs 240 / 320."""
import json
import os
import shutil
import subprocess
import sys

import pytest

from tests import util


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ for the host emulation build")
def test_biwfa_differential_fuzz_under_emulation():
    so = subprocess.run([os.path.join(util.ROOT, "tests", "emu", "build_emu.sh")], check=True, capture_output=True, text=True).stdout.strip().splitlines()[-1]
    env = dict(os.environ, WFB_LIB=os.path.join(util.ROOT, so))
    r = subprocess.run([sys.executable, os.path.join(util.ROOT, "tests", "wfa_fuzz.py"), "5", "200", "240"], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert res["mismatches"] == 0 and res["pairs"] >= 240, res


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ for the host emulation build")
def test_record_path_differential_fuzz_against_the_unmodified_reference():
    if util.load_wflign_ref() is None:
        pytest.skip("oracle/_ref/libwflignref.so not built (needs /root/reference)")
    so = subprocess.run([os.path.join(util.ROOT, "tests", "emu", "build_emu.sh")], check=True, capture_output=True, text=True).stdout.strip().splitlines()[-1]
    env = dict(os.environ, WFB_LIB=os.path.join(util.ROOT, so))
    r = subprocess.run([sys.executable, os.path.join(util.ROOT, "tests", "paf_fuzz.py"), "9", "200", "320"], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert res["mismatches"] == 0 and res["records"] >= 320, res
