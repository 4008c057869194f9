"""Inputs of the chain / filter stage tests (checker side): synthetic per-query fragment mappings shaped like what the
L2 stage hands to Map::filterSubsetMappings: collinear runs of window-sized mappings on both strands with jitter, gaps,
rearrangements, duplicates and noise."""
import numpy as np

import wfmash_b200 as wb


def query_mappings(rng, w=1000, qlen=300_000, nref=4, reflen=400_000):
    rows = []
    nfr = qlen // w
    for _ in range(int(rng.integers(1, 9))):  # collinear runs
        ref = int(rng.integers(0, nref))
        rev = bool(rng.integers(0, 2))
        f0 = int(rng.integers(0, nfr - 2))
        f1 = int(min(nfr, f0 + rng.integers(2, 140)))
        r0 = int(rng.integers(0, reflen - (f1 - f0 + 2) * w))
        drift = 0
        for f in range(f0, f1):
            if rng.random() < 0.08:
                continue  # a fragment that did not map
            if rng.random() < 0.03:
                drift += int(rng.integers(-3000, 6000))  # indel / jump
            pos = r0 + ((f1 - 1 - f) if rev else (f - f0)) * w + drift + int(rng.integers(-40, 41))
            pos = int(np.clip(pos, 0, reflen - w - 1))
            rows.append((ref, pos, f * w, w, 1, int(rng.integers(5, 30)), int(rng.integers(8500, 10001)), 1 if rev else 0, int(rng.integers(60, 101))))
            if rng.random() < 0.02:
                rows.append(rows[-1])  # exact duplicate (ties in both sorts)
    for _ in range(int(rng.integers(0, 30))):  # noise
        rows.append((int(rng.integers(0, nref)), int(rng.integers(0, reflen - w - 1)), int(rng.integers(0, nfr)) * w, w, 1, int(rng.integers(3, 12)),
                     int(rng.integers(7000, 9500)), int(rng.integers(0, 2)), int(rng.integers(30, 101))))
    m = np.zeros(len(rows), dtype=wb.MAPPING_DTYPE)
    for i, r in enumerate(rows):
        m[i] = r
    return m[rng.permutation(len(m))]


def batch(seed, nq=12, **kw):
    rng = np.random.default_rng(seed)
    qs = [query_mappings(rng, **kw) for _ in range(nq)]
    off = np.cumsum([0] + [len(q) for q in qs]).astype(np.int64)
    return np.concatenate(qs), off
