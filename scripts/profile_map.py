"""One pass of the mapping path (index build, L1 + L2) on the bench's C3-shaped synthetic set, for ncu captures:
  ncu --set full --clock-control none --import-source on -k regex:'l2_kernel|ix_l1_kernel|mm_stream_kernel' -o gpurun_out/map python scripts/profile_map.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wfmash_b200 as wb
from wfmash_b200 import synth

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000
rng = np.random.default_rng(77)
root = synth.random_seq(scale, rng)
seqs = [root.tobytes()] + [synth.mutate(root, 0.03, rng).tobytes() for _ in range(7)]
k, w, s = 15, 1000, 29
ix = wb.Index(seqs, list(range(8)), k, w, s, index_threads=8)
blob = b"".join(seqs)
offs = np.cumsum([0] + [len(x) for x in seqs])
frags = np.array([(int(offs[q]) + j * w, w, q) for q in range(8) for j in range(len(seqs[q]) // w)], dtype=wb.FRAG_DTYPE)
fqs = np.array([(q, q) for q in range(8) for j in range(len(seqs[q]) // w)], dtype=wb.FRAG_QUERY_DTYPE)
cut = np.array([max(1, int(i * 0.6)) for i in range(1001)], dtype=np.int32)
m = ix.map_fragments(blob, frags, fqs, 3, cut, np.arange(8, dtype=np.int32), stage1_min_hits=wb.stage1_min_hits(k, s), l2_min_shared=wb.l2_min_shared(0.85, k, s))
print({kk: (float(v) if not hasattr(v, "__len__") else len(v)) for kk, v in m.items() if kk != "l1"})
