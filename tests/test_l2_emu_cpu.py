"""CPU checks of the L2 stage:
  * the oracle's computeL2MappedRegions / SlideMapper restatement against the committed reference fixture and, when
    oracle/_ref is present, against the reference's unmodified slidingMap.hpp + mappingCore.hpp compiled in place;
  * the L2 CUDA kernel's BODY (wfmash_b200/csrc/l2_kernels.h) under the single-thread host emulation of tests/emu
    (TEST INFRASTRUCTURE, -DWFB_EMU; never the product library) against the oracle."""
import ctypes
import gzip
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

from tests import maputil, util

vp = maputil.vp
L2_CASES = [(41, (1, 1, 0, 3), (15, 1000, 29)), (9, (0, 0, 0, 2), (15, 1000, 59)), (7, (0, 1, 1, 5), (19, 500, 17)), (11, (1, 1, 0, 2), (15, 256, 11))]


def test_l2_oracle_reproduces_reference_fixture(oracle):
    with gzip.open(os.path.join(util.GOLD, "l2_reference.json.gz"), "rt") as f:
        gold = json.load(f)
    assert len(gold["cases"]) == len(L2_CASES)
    for (seed, mode, (k, w, s)), g in zip(L2_CASES, gold["cases"]):
        assert (seed, list(mode), [k, w, s]) == (g["seed"], g["mode"], g["kws"])
        seqs, ids, groups = maputil.l2_case(seed=seed)
        index = maputil.oracle_index(oracle, seqs, ids, k, w, s, 0.0002, 3)
        rows = maputil.l2_all_loci(oracle, "orc", index, seqs, ids, groups, k, w, s, oracle, mode)
        want = np.array(g["rows"], dtype=np.int64).reshape(-1, 8)
        assert len(want) > 100 and rows.shape == want.shape and (rows == want).all()
        assert (want[:, 7] == -1).sum() > 20  # both strands are exercised


@pytest.mark.ref
def test_l2_oracle_matches_compiled_reference_live(oracle):
    ref = util.load_ref("libl2ref.so")
    if ref is None:
        pytest.skip("oracle/_ref not built (reference sources absent)")
    for seed, mode, (k, w, s) in [(3, (1, 1, 0, 3), (15, 1000, 29)), (4, (0, 0, 0, 2), (15, 1000, 39)), (6, (0, 1, 0, 2), (21, 300, 13))]:
        seqs, ids, groups = maputil.l2_case(seed=seed)
        index = maputil.oracle_index(oracle, seqs, ids, k, w, s, 0.0002, 2)
        a = maputil.l2_all_loci(ref, "ref", index, seqs, ids, groups, k, w, s, oracle, mode)
        b = maputil.l2_all_loci(oracle, "orc", index, seqs, ids, groups, k, w, s, oracle, mode)
        assert len(a) > 100 and a.shape == b.shape and (a == b).all()


def same_mappings(got, want):
    if len(got) != len(want):
        return False
    for f in ("frag", "refSeqId", "refStartPos", "optimalStart", "optimalEnd", "conservedSketches", "strand"):
        if not (got[f] == want[f]).all():
            return False
    return bool((got["nucIdentity"].view(np.uint32) == want["nucIdentity"].view(np.uint32)).all()
                and (got["kmerComplexity"].view(np.uint32) == want["kmerComplexity"].view(np.uint32)).all())


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ for the host emulation build")
def test_l2_kernel_body_under_emulation_matches_oracle(oracle):
    so = subprocess.run([os.path.join(util.ROOT, "tests", "emu", "build_emu.sh")], check=True, capture_output=True, text=True).stdout.strip().splitlines()[-1]
    emu = ctypes.CDLL(os.path.join(util.ROOT, so))
    total = 0
    for seed, mode, (k, w, s) in L2_CASES:
        seqs, ids, groups = maputil.l2_case(seed=seed)
        index = maputil.oracle_index(oracle, seqs, ids, k, w, s, 0.0002, 3)
        kept = np.ascontiguousarray(index[0])
        for stage1 in (True, False):
            s1 = np.zeros(s + 1, dtype=np.int32)
            assert emu.wfb_stage1_min_hits(ctypes.c_double(1.0), ctypes.c_float(0.0), k, s, vp(s1)) == 0
            assert (s1 == maputil.stage1_table(oracle, 1.0, 0.0, k, s)).all()
            ms = np.zeros(s + 1, dtype=np.int32)
            assert emu.wfb_l2_min_shared(ctypes.c_float(0.85), k, s, vp(ms)) == 0
            frs, q_all, q_count, loci, want = maputil.oracle_map_fragments(oracle, index, seqs, ids, groups, k, w, s, mode, stage1=stage1,
                                                                           min_shared=ms if not stage1 else None)
            l1 = np.zeros(len(loci), dtype=maputil.L1PUBDT)
            l1["seqId"], l1["rangeStartPos"], l1["rangeEndPos"], l1["intersectionSize"] = loci[:, 1], loci[:, 2], loci[:, 3], loci[:, 4]
            lfrag = np.ascontiguousarray(loci[:, 0].astype(np.int32))
            out = np.zeros(len(want) + 64, dtype=maputil.L2MAPDT)
            n_out, steps = ctypes.c_int64(0), ctypes.c_uint64(0)
            status = np.zeros(len(frs), dtype=np.int32)
            rc = emu.wfb_emu_l2_loci(vp(kept), ctypes.c_int64(len(kept)), vp(l1), vp(lfrag), ctypes.c_int64(len(l1)), vp(q_all), vp(q_count), len(frs),
                                     k, w, s, vp(s1) if stage1 else None, None if stage1 else vp(ms), vp(out), ctypes.c_int64(len(out)),
                                     ctypes.byref(n_out), vp(status), ctypes.byref(steps))
            assert rc == 0 and (status == 0).all()
            assert same_mappings(out[: n_out.value], want), (seed, stage1)
            assert steps.value > len(l1)
            total += n_out.value
    assert total > 2000


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ for the host emulation build")
def test_l2_kernel_body_under_emulation_on_random_parameter_sets(oracle):
    """Random (k, w, s), filter modes, frequency cut-offs, index partitions, identity thresholds and sequence sets: the L2 kernel body against the
    oracle (L2_CASES above pins four hand-picked sets). 5 200 cases / 6 M mappings of this generator ran clean at the end of round 2; 12 here."""
    import random
    so = subprocess.run([os.path.join(util.ROOT, "tests", "emu", "build_emu.sh")], check=True, capture_output=True, text=True).stdout.strip().splitlines()[-1]
    emu = ctypes.CDLL(os.path.join(util.ROOT, so))
    rnd = random.Random(77)
    total = 0
    for _ in range(12):
        seed = rnd.randrange(1, 1 << 20)
        k, w, s = rnd.choice([(15, 1000, 29), (15, 1000, 59), (19, 500, 17), (15, 256, 11), (15, 1000, 24), (17, 2000, 39), (15, 1000, 5), (11, 300, 40)])
        mode = (rnd.choice([0, 1]), rnd.choice([0, 1]), rnd.choice([0, 1]), rnd.choice([1, 2, 3, 5]))
        seqs, ids, groups = maputil.l2_case(seed=seed)
        index = maputil.oracle_index(oracle, seqs, ids, k, w, s, rnd.choice([0.0002, 0.001, 0.01]), rnd.choice([1, 2, 3]))
        kept = np.ascontiguousarray(index[0])
        stage1 = rnd.random() < 0.5
        s1 = np.zeros(s + 1, dtype=np.int32)
        assert emu.wfb_stage1_min_hits(ctypes.c_double(1.0), ctypes.c_float(0.0), k, s, vp(s1)) == 0
        ms = np.zeros(s + 1, dtype=np.int32)
        assert emu.wfb_l2_min_shared(ctypes.c_float(rnd.choice([0.7, 0.85, 0.95])), k, s, vp(ms)) == 0
        frs, q_all, q_count, loci, want = maputil.oracle_map_fragments(oracle, index, seqs, ids, groups, k, w, s, mode, stage1=stage1,
                                                                       min_shared=ms if not stage1 else None)
        if len(loci) == 0:
            continue
        l1 = np.zeros(len(loci), dtype=maputil.L1PUBDT)
        l1["seqId"], l1["rangeStartPos"], l1["rangeEndPos"], l1["intersectionSize"] = loci[:, 1], loci[:, 2], loci[:, 3], loci[:, 4]
        lfrag = np.ascontiguousarray(loci[:, 0].astype(np.int32))
        out = np.zeros(len(want) + 64, dtype=maputil.L2MAPDT)
        n_out, steps = ctypes.c_int64(0), ctypes.c_uint64(0)
        status = np.zeros(len(frs), dtype=np.int32)
        rc = emu.wfb_emu_l2_loci(vp(kept), ctypes.c_int64(len(kept)), vp(l1), vp(lfrag), ctypes.c_int64(len(l1)), vp(q_all), vp(q_count), len(frs),
                                 k, w, s, vp(s1) if stage1 else None, None if stage1 else vp(ms), vp(out), ctypes.c_int64(len(out)),
                                 ctypes.byref(n_out), vp(status), ctypes.byref(steps))
        assert rc == 0 and (status == 0).all(), (seed, k, w, s, mode)
        assert same_mappings(out[: n_out.value], want), (seed, k, w, s, mode, stage1)
        total += n_out.value
    assert total > 3000


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++ for the host emulation build")
def test_l1_parallel_sweep_equals_serial_walk_under_emulation():
    """The data-parallel L1 sweep (ix_l1_regions_par: prefix sums / segment ids over the sorted interval points) against
    the literal two-pass walk of computeL1CandidateRegions (ix_l1_regions, itself checked against the oracle and the
    compiled reference on the GPU) on random interval sets: several sequences and PanSN groups, equal positions across a
    sequence boundary (the reference's position-only grouping), dense and sparse overlaps."""
    so = subprocess.run([os.path.join(util.ROOT, "tests", "emu", "build_emu.sh")], check=True, capture_output=True, text=True).stdout.strip().splitlines()[-1]
    emu = ctypes.CDLL(os.path.join(util.ROOT, so))
    rng = np.random.default_rng(5)
    cut = np.array([max(1, int(i * 0.5)) for i in range(1001)], dtype=np.int32)
    nl_total = 0
    for it in range(300):
        nseq = int(rng.integers(1, 7))
        groups = np.sort(rng.integers(0, 3, size=nseq)).astype(np.int32)  # ascending seqId -> non-decreasing group, like PanSN order
        w = int(rng.choice([200, 1000]))
        nint = int(rng.integers(1, 900))
        seq = rng.integers(0, nseq, size=nint)
        span = int(rng.choice([300, 3000, 50000]))
        start = rng.integers(0, span, size=nint)
        if it % 3 == 0:
            start = (start // 7) * 7  # many equal positions, also across sequences
        length = rng.integers(1, w + 1, size=nint)
        keys = np.concatenate([(seq.astype(np.uint64) << np.uint64(41)) | (start.astype(np.uint64) << np.uint64(1)) | np.uint64(1),
                               (seq.astype(np.uint64) << np.uint64(41)) | ((start + length).astype(np.uint64) << np.uint64(1))])
        keys = np.sort(keys)
        mh = int(rng.integers(1, 6))
        sp = int(rng.integers(0, 2))
        a = np.zeros(256, dtype=maputil.L1PUBDT); b = np.zeros(256, dtype=maputil.L1PUBDT)
        na, nb = ctypes.c_int32(0), ctypes.c_int32(0)
        rc = emu.wfb_emu_l1_sweeps(vp(keys), len(keys), 29, w, 29, mh, sp, vp(cut), len(cut), vp(groups), vp(a), ctypes.byref(na), vp(b), ctypes.byref(nb))
        if rc == -5:
            continue  # more than 256 loci: both paths report the capacity error
        assert rc == 0
        assert na.value == nb.value, (it, na.value, nb.value)
        assert (a[: na.value] == b[: nb.value]).all(), it
        nl_total += na.value
    assert nl_total > 1500
