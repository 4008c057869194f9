"""BASELINE.json's configs C1-C3 as (input files, CLI-equivalent parameters). TEST / BENCH INFRASTRUCTURE ONLY.

The reference's own CTest runs exactly these files (CMakeLists.txt:436-464); its timings are in doc/performance-tuning.md:176,209,311-319.
percentage_identity None = the CLI default `-p ani50-2` (main.cpp:75-134)."""
import numpy as np

from tests import datasets

CONFIGS = [
    # C1: `wfmash data/reference.fa.gz data/reads.255bps.fa.gz`: every read is shorter than the segment length, so no fragment is mapped
    dict(name="C1", target="reference", query="reads255", params=dict(), full_only=False),
    # companion run with -w 250
    dict(name="C1w250", target="reference", query="reads255", params=dict(window_length=250), full_only=False),
    # C2: `wfmash data/LPA.subset.fa.gz -k15 -w1k -P50k`
    dict(name="C2", target="lpa", query=None, params=dict(kmer_size=15, window_length=1000, filter=dict(max_mapping_length=50000)), full_only=False),
    # the reference's own CTest / tuning-log run on the same file: `wfmash data/LPA.subset.fa.gz -p 80 -n 5` (CMakeLists.txt:438-441,
    # doc/performance-tuning.md:311-319: divergent records, compute_affine2p dominated)
    dict(name="C2p80n5", target="lpa", query=None, params=dict(percentage_identity=0.80, filter=dict(num_mappings_for_segment=5)), full_only=False),
    # C3 subset: two yeast genomes x three chromosomes, -Y '#'
    dict(name="C3sub", target="yeast", query=None, subset=dict(genomes=2, chroms=("chrI", "chrVI", "chrIII")), params=dict(), full_only=False),
    # C3: `wfmash data/scerevisiae8.fa.gz -Y '#'` (all 8 genomes)
    dict(name="C3", target="yeast", query=None, params=dict(), full_only=True),
    # C4 / C5 (synthetic, xoshiro256** seed 42, PanSN names) at 1/50 and 1/1000 of their stated sizes; `-p 90 -P50k` / `-p 80`
    # explicit as 8(d) asks, so the ANI estimate is not on the path. Pairwise divergence 10 % / 20 %: the regime of C4 / C5's records.
    dict(name="C4s", target="synth_c4s", query=None, params=dict(percentage_identity=0.90, filter=dict(max_mapping_length=50000)), full_only=True),
    dict(name="C5s", target="synth_c5s", query=None, params=dict(percentage_identity=0.80), full_only=True),
]
for _c in CONFIGS:
    _c["params"].setdefault("percentage_identity", None)

_cache = {}


def _load(key):
    if key not in _cache:
        _cache[key] = datasets.load(key)
    return _cache[key]


def sequences(cfg):
    """-> (targets, queries); the same list object twice for a self all-vs-all run."""
    t = _load(cfg["target"])
    if cfg.get("subset"):
        t = datasets.yeast_subset(t, **cfg["subset"])
    if cfg["query"] is None:
        return t, t
    return t, _load(cfg["query"])


def by_name(name):
    return next(c for c in CONFIGS if c["name"] == name)


def adopt_identity(estimate: float) -> float:
    """main.cpp:104: `map_parameters.percentageIdentity = estimated_identity` narrows the double to the float member."""
    return float(np.float32(estimate))
