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
