"""A/B of the mapping-path kernels on the bench's C3-shaped set: prints L1 / L2 kernel times (best of 3)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wfmash_b200 as wb
from wfmash_b200 import synth
rng = np.random.default_rng(77)
root = synth.random_seq(3_000_000, rng)
seqs = [root.tobytes()] + [synth.mutate(root, 0.03, rng).tobytes() for _ in range(7)]
k, w, s = 15, 1000, 29
ix = wb.Index(seqs, list(range(8)), k, w, s, index_threads=8)
blob = b"".join(seqs)
offs = np.cumsum([0] + [len(x) for x in seqs])
frags = np.array([(int(offs[q]) + j * w, w, q) for q in range(8) for j in range(len(seqs[q]) // w)], dtype=wb.FRAG_DTYPE)
fqs = np.array([(q, q) for q in range(8) for j in range(len(seqs[q]) // w)], dtype=wb.FRAG_QUERY_DTYPE)
cut = np.array([max(1, int(i * 0.6)) for i in range(1001)], dtype=np.int32)
best = {}
for _ in range(4):
    t0 = time.perf_counter()
    m = ix.map_fragments(blob, frags, fqs, 3, cut, np.arange(8, dtype=np.int32), stage1_min_hits=wb.stage1_min_hits(k, s), l2_min_shared=wb.l2_min_shared(0.85, k, s))
    dt = (time.perf_counter() - t0) * 1e3
    for kk, v in (("l1_kernel_ms", m["l1_kernel_ms"]), ("l2_kernel_ms", m["l2_kernel_ms"]), ("sort_kernel_ms", m["sort_kernel_ms"]), ("call_ms", dt)):
        best[kk] = min(best.get(kk, 1e9), v)
print(os.environ.get("WFB_L1_SERIAL", "par"), {kk: round(v, 3) for kk, v in best.items()}, "loci", int(m["n_l1_loci"]), "mappings", len(m["mappings"]))
