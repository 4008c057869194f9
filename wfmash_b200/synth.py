"""Seeded synthetic genomes / mapping records of the shapes BASELINE.json names (there is no network
and /root/reference does not exist on the GPU box, so bench.py and the full-size tests use these).

Model: i.i.d. uniform ACGT root; a derived sequence applies per-base independent
events at total rate d split substitution : insertion : deletion = 8 : 1 : 1, indel length
geometric with mean 3. numpy's PCG64 seeded generator replaces xoshiro256** (any fixed, seeded
stream serves; the exact generator is not part of the metric)."""
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def random_seq(n, rng):
    return _ACGT[rng.integers(0, 4, size=n)]


def mutate(seq, d, rng, sub_frac=0.8, ins_frac=0.1, indel_mean=3.0):
    """Return a mutated copy of seq (uint8 array) with total per-base event rate d."""
    n = int(seq.shape[0])
    if n == 0 or d <= 0:
        return seq.copy()
    r = rng.random(n)
    sub = r < d * sub_frac
    ins = (r >= d * sub_frac) & (r < d * (sub_frac + ins_frac))
    dele = (r >= d * (sub_frac + ins_frac)) & (r < d)
    out = seq.copy()
    # substitutions: always to a different base
    ns = int(sub.sum())
    if ns:
        code = np.searchsorted(_ACGT, out[sub])  # A,C,G,T sorted ascending in ASCII
        out[sub] = _ACGT[(code + rng.integers(1, 4, size=ns)) % 4]
    keep = np.ones(n, dtype=bool)
    dpos = np.flatnonzero(dele)
    if dpos.size:
        dl = rng.geometric(1.0 / indel_mean, size=dpos.size)
        diff = np.zeros(n + 1, dtype=np.int64)
        np.add.at(diff, dpos, 1)
        np.add.at(diff, np.minimum(dpos + dl, n), -1)
        keep = np.cumsum(diff[:n]) == 0
    ins_len = np.zeros(n, dtype=np.int64)
    ipos = np.flatnonzero(ins)
    if ipos.size:
        ins_len[ipos] = rng.geometric(1.0 / indel_mean, size=ipos.size)
    contrib = keep.astype(np.int64) + ins_len
    total = int(contrib.sum())
    res = np.empty(total, dtype=np.uint8)
    start = np.cumsum(contrib) - contrib
    kept_idx = np.flatnonzero(keep)
    res_is_kept = np.zeros(total, dtype=bool)
    res_is_kept[start[kept_idx]] = True
    res[start[kept_idx]] = out[kept_idx]
    n_ins = total - kept_idx.size
    if n_ins:
        res[~res_is_kept] = _ACGT[rng.integers(0, 4, size=n_ins)]
    return res


def mapping_records(n, seed, len_lo, len_hi, divergences, pad=0):
    """n synthetic mapping records as the aligner sees them (computeAlignments.hpp:195-303):
    (target slice = pattern, query slice = text), query length ~ U[len_lo, len_hi], the target is the
    query's source mutated at a divergence drawn from `divergences`, plus `pad` unrelated flanking bases
    on both target ends (wfmash pads the target by min(w,5000), parse_args.hpp:608)."""
    rng = np.random.default_rng(seed)
    recs = []
    for _ in range(n):
        qlen = int(rng.integers(len_lo, len_hi + 1))
        d = float(divergences[int(rng.integers(0, len(divergences)))])
        q = random_seq(qlen, rng)
        t = mutate(q, d, rng)
        if pad:
            t = np.concatenate([random_seq(pad, rng), t, random_seq(pad, rng)])
        recs.append((t.tobytes(), q.tobytes(), d))
    return recs


def genome(n_contigs, contig_len, seed):
    rng = np.random.default_rng(seed)
    return [random_seq(contig_len, rng) for _ in range(n_contigs)]
