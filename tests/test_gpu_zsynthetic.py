"""Round-2 additions that run LAST among the GPU tests (file name order):
  * SURVEY 8(d)'s synthetic configs C4 / C5, scaled (tests/configs.py C4s, C5s: xoshiro256** seed 42, PanSN names, sequences generated on
    the spot), end to end through wfb_map_phase + wfb_align_phase against the text the reference's UNMODIFIED skch::Map + align::Aligner
    wrote for the same sequences (tests/golden/config_reference.json.gz, made by tests/golden/make_config_golden.py --full --only C4s / C5s);
  * the three builds of the reference-side minmers (candidate-filtered, filtered with every tile overflowing, unfiltered) on real sequence;
  * targets of exactly one window: the order of minmers that tie on (wpos, wpos_end) must be the reference's (its unstable std::sort's)."""
import json

import pytest

from tests import configrun
from tests.test_gpu_configs import _check

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def wb():
    import wfmash_b200 as w
    if w.device_count() < 1:
        pytest.fail("no CUDA device: the -m gpu tests must run on the B200 box")
    return w


@pytest.mark.parametrize("name", ["C4s", "C5s"])
def test_synthetic_c4_c5_scaled_paf_is_byte_identical_to_the_reference(wb, name):
    # SURVEY 8(d)'s synthetic configs (xoshiro256** seed 42, PanSN names, `-p 90 -P50k` / `-p 80`) at a size the CPU reference finishes in
    # minutes: C4s = 2 genomes x 8 contigs x 2.5 Mbp at 90 % ANI (4 970 mapping records, 42 Mbp aligned at ~10 % divergence), C5s = 5
    # haplotypes x 1 Mbp at 80 % ANI (3 123 records, 17 Mbp aligned at ~20 % divergence). The sequences are generated here, from the seed.
    r = configrun.run(wb, name)
    print(json.dumps(configrun.summary(r)))
    _check(r, name)
    assert r["alignment_lines"] > 3000
    assert r["align_stats"].patch_cap_kept_main == 0 and r["align_stats"].main_device_cap == 0


def test_minmer_build_modes_give_the_same_index_and_mappings(wb, monkeypatch):
    """The candidate-filtered minmer build (mm_cand_kernel + mm_stream_cand_kernel, the default), the same with every tile overflowing
    (all chunks re-run by the exact kernel) and the unfiltered build must produce the same minmers on real sequence (two yeast genomes)
    and hence the same mapping text."""
    from tests import configs
    seqs, _ = configs.sequences(configs.by_name("C3sub"))
    outs = {}
    for mode, env in (("filtered", {}), ("overflow", {"WFB_MM_CAND_CAP": "150"}), ("unfiltered", {"WFB_MM_FILTER": "0"})):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        mm, st = wb.minmers_build([x for _, x in seqs], list(range(len(seqs))), 15, 1000, 24)
        r = configrun.run(wb, "C3sub", align=False)
        for k in env:
            monkeypatch.delenv(k)
        assert st.filtered == (mode != "unfiltered") and st.stale_absorbed == 0 and st.stitch_miss == 0, (mode, st.as_dict())
        assert (st.redo_chunks == st.chunks) == (mode == "overflow"), (mode, st.as_dict())
        assert r["mapping_identical"], mode
        outs[mode] = (mm.tobytes(), r["mapping_paf"])
    assert outs["filtered"] == outs["unfiltered"] == outs["overflow"]


def test_one_window_targets_map_like_the_reference(wb):
    """tests/golden/tiny_target_reference.json.gz (made by tests/golden/make_tiny_target_golden.py from the unmodified skch::Map): targets of
    1000 - 2000 bases, whose minmers all tie on (wpos, wpos_end). The index must hold them in the order the reference's std::sort leaves
    (wfb_minmers_build puts such sequences in order on the host: stats.tie_sequences), or the L2 stage counts other shared minmers
    (c#1#chr1 -> d#1#chrZ: 5 in the reference, 4 with the ties in hash order)."""
    import gzip
    import os
    from tests import util
    with gzip.open(os.path.join(util.GOLD, "tiny_target_reference.json.gz"), "rt") as f:
        doc = json.load(f)
    seqs = [(n, s.encode()) for n, s in doc["sequences"]]
    mm, st = wb.minmers_build([x for _, x in seqs], list(range(len(seqs))), 15, 1000, 59)
    assert st.tie_sequences >= 3
    for name, c in doc["cases"].items():
        prm = dict(c["params"])
        f = dict(prm.pop("filter", {}))
        MP = wb.MapPhaseParams(filter=wb.FilterParams(window_length=1000, **f), **prm)
        mp, mst = wb.map_phase(seqs, seqs, MP)
        cut = lambda rows: sorted("\t".join(r.split("\t")[:14]) for r in rows)   # ch:Z: is schedule dependent in the reference for long queries
        assert cut(ln.decode() for ln in mp.split(b"\n") if ln) == cut(c["rows"]), name
