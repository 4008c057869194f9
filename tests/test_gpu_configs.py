"""BASELINE.json's configs C1-C3 on the reference's own input files, end to end on the GPU: `wfb_map_phase` -> `wfb_align_phase`
(host sequences in, PAF text out) against what the reference's UNMODIFIED skch::Map + align::Aligner wrote for the same files
(tests/golden/config_reference.json.gz, generated in the build container by tests/golden/make_config_golden.py; the GPU box has
no /root/reference). Every mapping line (ch:Z: tags of the `-t 1` run included) and every alignment line (cg:Z: CIGAR, gi / bi / md
tags) must be byte-identical; the adopted ANI identity must be the same float; the index build must not have met the one
condition under which its chunked minmer stream could differ from the reference's (stale_absorbed == 0)."""
import pytest

from tests import configrun

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def wb():
    import wfmash_b200 as w
    if w.device_count() < 1:
        pytest.fail("no CUDA device: the -m gpu tests must run on the B200 box")
    return w


def _check(r, name):
    s = configrun.summary(r)
    assert r["identity_identical"], s
    assert r["map_stats"].stale_absorbed == 0, s
    assert r["mapping_identical"], (s, r["mapping_only_ours"][:3], r["mapping_only_ref"][:3])
    assert r["alignment_identical"], (s, r["alignment_only_ours"][:3], r["alignment_only_ref"][:3])
    g = configrun.golden()[name]
    assert r["mapping_rows"] == g["mapping_rows"] and r["alignment_lines"] == g["alignment_lines"] and r["aligned_bp"] == g["aligned_bp"], s


@pytest.mark.parametrize("name", ["C1", "C1w250"])
def test_c1_reads_vs_reference_is_empty_on_both_sides(wb, name):
    # every read is shorter than the segment length (C1) / none of them comes from this reference (C1w250): the reference writes
    # nothing, and so must we, through the same plumbing (separate query file, ANI fallback 0.7)
    r = configrun.run(wb, name)
    _check(r, name)
    assert r["mapping_paf"] == b"" and r["alignment_paf"] == b""


@pytest.mark.parametrize("name", ["C2", "C2p80n5", "C3sub"])
def test_config_paf_is_byte_identical_to_the_reference(wb, name):
    r = configrun.run(wb, name)
    _check(r, name)
    assert r["alignment_lines"] > 50


def test_c3_all_eight_yeast_genomes_paf_is_byte_identical_to_the_reference(wb):
    # the bench workload: 136 sequences, 96 Mbp, 21 129 mapping records, 659 Mbp aligned
    r = configrun.run(wb, "C3")
    _check(r, "C3")


def test_capacity_fallbacks_and_scheduling_variants_do_not_change_the_text(wb, monkeypatch):
    """The paths that only run when a capacity is exceeded, forced through their test hooks on C3sub (two yeast genomes), and the
    switches that only reorder work must leave every output byte where it was:
      * the library's L1 loci buffer starting at 64 entries (grown from the kernel's own count, index_host.cu);
      * per-fragment L1 scratch of 8192 gathered points / 4 candidate regions (the flagged fragments are re-run alone with larger scratch);
      * head and tail patches in the reference's two rounds instead of one fused round (epilogue.cu patch_records);
      * no packed sequences in shared memory (wfa_kernels.h wfb_pack_seq)."""
    base = configrun.run(wb, "C3sub")
    _check(base, "C3sub")
    assert base["align_stats"].patch_cap_kept_main == 0 and base["align_stats"].main_device_cap == 0
    for env in ({"WFB_L1_LOCI_CAP0": "64"}, {"WFB_L1_FRAG_CAPS": "8192,4"}, {"WFB_PATCH_FUSE": "0"}, {"WFB_SEQ_SMEM_KB": "0"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        r = configrun.run(wb, "C3sub", align=not any(k.startswith("WFB_L1") for k in env))
        for k in env:
            monkeypatch.delenv(k)
        assert r["mapping_paf"] == base["mapping_paf"], env
        if "alignment_paf" in r:
            assert r["alignment_paf"] == base["alignment_paf"], env
