"""Host stage between L2 and the aligner (first part): wfb_chain_mappings_batch and
wfb_l2_to_query_mappings against the reference's UNMODIFIED mappingFilter.hpp / mappingOutput.hpp compiled in place
(oracle/_ref/libfilterref.so) and against the committed fixture generated from it (tests/golden/chain_reference.json.gz).
These entry points are host C++ inside the product library: no GPU is needed to run them."""
import ctypes
import gzip
import hashlib
import json
import os

import numpy as np
import pytest

from tests import chainutil, util

CASES = [(1, dict(w=1000), dict(chain_gap=2000, max_mapping_length=50000)), (2, dict(w=1000), dict(chain_gap=20000, max_mapping_length=50000)),
         (3, dict(w=500, qlen=120_000), dict(chain_gap=1000, max_mapping_length=10000)), (4, dict(w=1000), dict(chain_gap=2000, max_mapping_length=2**62)),
         (5, dict(w=1000), dict(chain_gap=2000, max_mapping_length=50000, split=False))]


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def ours(seed, gen, prm):
    import wfmash_b200 as wb
    m, off = chainutil.batch(seed, **gen)
    return wb.chain_mappings_batch(m, off, gen.get("w", 1000), host_threads=3, **prm), m, off


def reference(ref, m, off, w, chain_gap, max_mapping_length, split=True):
    import wfmash_b200 as wb
    ref.ref_merge_with_chains.restype = ctypes.c_int64
    ms, gs, cs, mo = [], [], [], [0]
    for q in range(len(off) - 1):
        a = np.array(m[off[q]: off[q + 1]], copy=True)
        g = np.zeros(len(a) + 4, dtype=wb.MAPPING_DTYPE); c = np.zeros(len(a) + 4, dtype=wb.CHAIN_INFO_DTYPE)
        n = ref.ref_merge_with_chains(ctypes.c_void_p(a.ctypes.data), ctypes.c_int64(len(a)), int(split), ctypes.c_int64(chain_gap), ctypes.c_int64(w),
                                      ctypes.c_uint64(max_mapping_length), q, ctypes.c_int64(300000), ctypes.c_void_p(g.ctypes.data),
                                      ctypes.c_void_p(c.ctypes.data), ctypes.c_int64(len(g)))
        ms.append(a); gs.append(g[:n]); cs.append(c[:n]); mo.append(mo[-1] + n)
    return np.concatenate(ms), np.concatenate(gs), np.concatenate(cs), np.array(mo, dtype=np.int64)


def test_chain_merge_reproduces_reference_fixture():
    with gzip.open(os.path.join(util.GOLD, "chain_reference.json.gz"), "rt") as f:
        gold = json.load(f)
    assert len(gold["cases"]) == len(CASES)
    for (seed, gen, prm), g in zip(CASES, gold["cases"]):
        (m2, merged, info, mo), m, off = ours(seed, gen, prm)
        assert len(m) == g["n_in"] and len(merged) == g["n_merged"], seed
        assert mo.tolist() == g["merged_offset"]
        assert digest(m2) == g["sha_reordered"], seed
        assert digest(merged) == g["sha_merged"], seed
        assert digest(info) == g["sha_chain_info"], seed
        assert [[int(x) for x in r] for r in merged[:5].tolist()] == g["head"]


@pytest.mark.ref
def test_chain_merge_matches_compiled_reference_live():
    ref = util.load_ref("libfilterref.so")
    if ref is None:
        pytest.skip("oracle/_ref not built (reference sources absent)")
    tot = 0
    for seed, gen, prm in [(11, dict(w=1000), dict(chain_gap=2000, max_mapping_length=50000)), (12, dict(w=200, qlen=60_000, reflen=90_000), dict(chain_gap=500, max_mapping_length=5000)),
                           (13, dict(w=1000, nref=1), dict(chain_gap=100000, max_mapping_length=50000))]:
        (m2, merged, info, mo), m, off = ours(seed, gen, prm)
        r_m, r_merged, r_info, r_mo = reference(ref, m, off, gen.get("w", 1000), **prm)
        assert (mo == r_mo).all()
        assert m2.tobytes() == r_m.tobytes() and merged.tobytes() == r_merged.tobytes() and info.tobytes() == r_info.tobytes()
        tot += len(merged)
        assert info["chainLen"].max() > 5 and (merged["n_merged"] > 1).any()
    assert tot > 300


@pytest.mark.ref
def test_l2_to_query_mappings_matches_compiled_reference_live():
    import wfmash_b200 as wb
    ref = util.load_ref("libfilterref.so")
    if ref is None:
        pytest.skip("oracle/_ref not built (reference sources absent)")
    rng = np.random.default_rng(3)
    w, qlen = 1000, 57_300
    ref_len = np.array([80_000, 30_500, 1_200], dtype=np.int64)
    nfr = qlen // w + 1
    frag_index = np.arange(nfr, dtype=np.int32)  # the tail fragment carries noOverlapFragmentCount
    n = 4000
    l2 = np.zeros(n, dtype=wb.L2_MAPPING_DTYPE)
    l2["frag"] = rng.integers(0, nfr, size=n)
    l2["refSeqId"] = rng.integers(0, 3, size=n)
    l2["refStartPos"] = rng.integers(0, 81_000, size=n)  # also beyond the end of the shorter sequences
    l2["conservedSketches"] = rng.integers(1, 30, size=n)
    l2["strand"] = rng.choice([-1, 1], size=n)
    l2["nucIdentity"] = rng.random(n, dtype=np.float32)
    l2["kmerComplexity"] = rng.random(n, dtype=np.float32)
    got = wb.l2_to_query_mappings(l2, frag_index, w, qlen, ref_len)
    exp = np.zeros(n, dtype=wb.MAPPING_DTYPE)
    one = np.zeros(1, dtype=wb.MAPPING_DTYPE)
    for i in range(n):
        ref.ref_make_mapping(int(l2["refSeqId"][i]), ctypes.c_int64(int(l2["refStartPos"][i])), ctypes.c_int64(w), int(l2["conservedSketches"][i]),
                             ctypes.c_float(float(l2["nucIdentity"][i])), ctypes.c_float(float(l2["kmerComplexity"][i])), int(l2["strand"][i]),
                             ctypes.c_void_p(one.ctypes.data))
        one["queryStartPos"] += np.uint32(int(frag_index[l2["frag"][i]]) * w)
        exp[i] = one[0]
    ref.ref_boundary_sanity(ctypes.c_void_p(exp.ctypes.data), ctypes.c_int64(n), ctypes.c_int64(qlen), ctypes.c_void_p(ref_len.ctypes.data))
    assert got.tobytes() == exp.tobytes()
