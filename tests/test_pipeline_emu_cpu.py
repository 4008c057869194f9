"""wfmash_b200.pipeline.align (mapping PAF -> padded, strand-corrected records -> batched do_biwfa_alignment -> alignment
PAF) WITHOUT a GPU: the kernel bodies run under the single-thread host emulation of tests/emu (TEST INFRASTRUCTURE,
-DWFB_EMU, in a subprocess; never the product library). The mapping PAF comes from the reference-side composition of
tests/pipeutil.py (the mapping kernels are not part of the emulation), and the alignment PAF is compared with the
reference's UNMODIFIED do_biwfa_alignment on the same records. The whole pipeline, mapping kernels included, is compared
the same way on the real device in tests/test_gpu_parity.py::test_pipeline_map_align_matches_reference_pieces."""
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
from tests import pipeutil, util
from wfmash_b200 import pipeline
seqs = pipeutil.case(length=14_000)
P = pipeline.Params()
fref, wref = util.load_ref("libfilterref.so"), util.load_wflign_ref()
out = {"have_ref": fref is not None and wref is not None}
if out["have_ref"]:
    mp, lines = pipeutil.expected(seqs, P, util.load_oracle(), fref, wref)
    paf, st = pipeline.align(mp, seqs, seqs, P)
    import wfmash_b200 as wb
    al = wb.Aligner(0)
    c_paf, c_st = wb.align_phase(al, mp + b"not a mapping row\n", seqs, seqs)   # the one-call C ABI phase (phases_host.cu) on the same text
    # robustness of the phase entry point: empty text, rows naming unknown sequences, the SAM branch
    e_paf, e_st = wb.align_phase(al, b"", seqs, seqs)
    rows = mp.split(b"\n")
    u_paf, u_st = wb.align_phase(al, rows[0].replace(b"a#1#chr1", b"nobody#1#x") + b"\n" + rows[1] + b"\n", seqs, seqs)
    s_paf, s_st = wb.align_phase(al, rows[1] + b"\n", seqs, seqs, sam_format=True, emit_md_tag=True)
    out.update(empty=[len(e_paf), int(e_st.records)], unknown=[int(u_st.records), int(u_st.skipped_lines), u_paf.decode() == lines[1].decode()],
               sam=s_paf.decode().split("\t")[:4] + [s_paf.decode().rstrip("\n").split("\t")[-1][:5]])
    A = util.load_ref("libalignref.so")   # the reference's whole alignment phase (align::Aligner::compute) on the same mapping PAF
    if A is not None:
        out["aligner_paf"] = pipeutil.reference_align_phase(A, mp, seqs, P).decode()
        out["aligner_sam"] = pipeutil.reference_align_phase(A, rows[1] + b"\n", seqs, P, sam_format=True, emit_md_tag=True).decode()
        out["ours_sam"] = s_paf.decode()
    out.update(c_paf=c_paf.decode(), c_records=int(c_st.records), c_skipped=int(c_st.skipped_lines), c_aligned_bp=int(c_st.aligned_bp))
    out.update(ref_map=mp.decode(), ref_paf=b"".join(lines).decode(), ours_paf=paf.decode(), records=st["records"], written=st["written"],
               aligned_bp=st["aligned_bp"])
print(json.dumps(out))
"""


@pytest.mark.ref
@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ for the host emulation build")
def test_align_phase_under_emulation_matches_reference_do_biwfa_alignment():
    so = subprocess.run([os.path.join(util.ROOT, "tests", "emu", "build_emu.sh")], check=True, capture_output=True, text=True).stdout.strip().splitlines()[-1]
    env = dict(os.environ, WFB_LIB=os.path.join(util.ROOT, so))
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": util.ROOT}], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    if not res["have_ref"]:
        pytest.skip("oracle/_ref not built (reference sources absent)")
    assert res["records"] >= 8 and res["written"] >= 8
    assert res["ours_paf"] == res["ref_paf"]
    assert res["c_paf"] == res["ref_paf"] and res["c_records"] == res["records"] and res["c_skipped"] == 1 and res["c_aligned_bp"] == res["aligned_bp"]
    assert {ln.split("\t")[4] for ln in res["ours_paf"].splitlines()} == {"+", "-"}
    if "aligner_paf" in res:   # byte-identical to what the reference's own Aligner writes to its output file
        assert res["aligner_paf"] == res["c_paf"] == res["ours_paf"]
        sam_records = [ln for ln in res["aligner_sam"].splitlines() if not ln.startswith("@")]   # the reference's file starts with @SQ / @PG header lines
        assert sam_records == res["ours_sam"].splitlines()
    assert res["empty"] == [0, 0] and res["unknown"] == [1, 1, True]
    assert res["sam"][0] == res["ref_paf"].splitlines()[1].split("\t")[0] and res["sam"][1] in ("0", "16") and res["sam"][4] == "MD:Z:"
    spans = [int(f[3]) - int(f[2]) for f in (ln.split("\t") for ln in res["ref_map"].splitlines())]
    assert res["aligned_bp"] >= sum(spans)  # + the query padding of the chain ends


@pytest.mark.ref
@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ for the host emulation build")
def test_align_phase_differential_fuzz_against_the_unmodified_aligner():
    """tests/alignphase_fuzz.py: hand-made mapping rows (sequence ends, both strands, shifted / resized target intervals, random chain tags; PAF
    and SAM + MD) through wfb_align_phase under emulation against the unmodified align::Aligner::compute. 1 430 cases / 5 100 rows ran clean at the
    end of round 2; 12 cases here."""
    if util.load_ref("libalignref.so") is None:
        pytest.skip("oracle/_ref not built (reference sources absent)")
    so = subprocess.run([os.path.join(util.ROOT, "tests", "emu", "build_emu.sh")], check=True, capture_output=True, text=True).stdout.strip().splitlines()[-1]
    env = dict(os.environ, WFB_LIB=os.path.join(util.ROOT, so))
    r = subprocess.run([sys.executable, os.path.join(util.ROOT, "tests", "alignphase_fuzz.py"), "3", "200", "12"], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert res["mismatches"] == 0 and res["cases"] == 12 and res["rows"] >= 20, res
