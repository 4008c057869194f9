"""The generator of SURVEY 8(d)'s synthetic configs (wfmash_b200/synth.py: xoshiro256** seed 42, PanSN names, 8 : 1 : 1 events):
published known-answer vectors of the two generators, the draws' distributions, and the digest that ties the generated C5s sequences to the
reference fixture (tests/golden/config_reference.json.gz holds what the unmodified reference wrote for exactly these bytes)."""
import numpy as np

from tests import configrun, configs, datasets
from wfmash_b200 import synth


def test_xoshiro256ss_and_splitmix64_known_answers():
    # splitmix64(0) and xoshiro256** from state (1, 2, 3, 4): the vectors of the authors' reference C code
    assert synth._splitmix64(0)[1] == 0xE220A8397B1DCDAF
    s = [1, 2, 3, 4]
    assert [synth._xo_next(s) for _ in range(4)] == [11520, 0, 1509978240, 1215971899390074240]
    # lane 0 of the vectorised generator is the scalar generator seeded through splitmix64; lane 1 is one jump() ahead
    x, st = 42, []
    for _ in range(4):
        x, z = synth._splitmix64(x)
        st.append(z)
    lane0 = list(st)
    want0 = [synth._xo_next(lane0) for _ in range(3)]
    lane1 = list(st)
    synth._xo_jump(lane1)
    want1 = [synth._xo_next(lane1) for _ in range(3)]
    v = synth.Xoshiro256ss(42, lanes=4).raw(12).reshape(3, 4)
    assert [int(a) for a in v[:, 0]] == want0 and [int(a) for a in v[:, 1]] == want1


def test_draws_have_the_stated_distributions():
    r = synth.Xoshiro256ss(7)
    u = r.random(400_000)
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.003
    g = r.geometric(1.0 / 3.0, 400_000)
    assert g.min() == 1 and abs(g.mean() - 3.0) < 0.03 and abs((g == 1).mean() - 1 / 3) < 0.005
    i = r.integers(1, 4, 300_000)
    assert set(np.unique(i)) == {1, 2, 3}


def test_c5s_sequences_are_the_ones_the_fixture_was_made_from():
    seqs = datasets.load("synth_c5s")  # asserts the digest
    assert [n for n, _ in seqs] == [f"h{i:03d}#1#chr1" for i in range(1, 6)]
    a, b = (np.frombuffer(x, dtype=np.uint8) for _, x in seqs[:2])
    assert abs(len(a) - 1_000_000) < 5000 and set(np.unique(a)) == set(b"ACGT")
    g = configrun.golden()
    for name, rows in (("C4s", 4970), ("C5s", 3123)):
        assert g[name]["mapping_rows"] == rows and g[name]["alignment_lines"] > 3000 and configs.by_name(name)["full_only"]
