#!/usr/bin/env python3
"""Both phases sharded over the GPUs of one box (one process per GPU) and compared with the single-GPU run on rank 0:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/pipeline_sharded.py
Mapping: queries partitioned by length, index replicated; alignment: mapping rows partitioned by estimated cost; NCCL only
gathers the text on rank 0 (wfmash_b200/shard.py)."""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import torch.distributed as dist
from wfmash_b200 import pipeline, shard, synth

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rng = np.random.default_rng(4242)
seqs = []
for c in range(6):
    root = synth.random_seq(300_000, rng)
    seqs.append((f"gA#1#chr{c + 1:02d}", synth.mutate(root, 0.025, rng).tobytes()))
    seqs.append((f"gB#1#chr{c + 1:02d}", synth.mutate(root, 0.025, rng).tobytes()))
P = pipeline.Params(percentage_identity=0.90)
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
mp = shard.map_sharded(seqs, seqs, P, device=local)
box = [mp]
dist.broadcast_object_list(box, src=0)          # every rank needs the rows to take its share of the alignment work
paf = shard.align_paf_sharded(box[0], seqs, seqs, P, device=local)
dist.barrier(); torch.cuda.synchronize()
t1 = time.perf_counter()
if rank == 0:
    m1 = pipeline.map(seqs, seqs, P, local)
    p1, st = pipeline.align(m1.paf, seqs, seqs, P, local)
    t2 = time.perf_counter()
    print(json.dumps({"world": world, "mapping_paf_identical": sorted(mp.split(b"\n")) == sorted(m1.paf.split(b"\n")), "alignment_paf_identical": paf == p1 if mp == m1.paf else sorted(paf.split(b"\n")) == sorted(p1.split(b"\n")),
                      "records": st["records"], "aligned_bp": st["aligned_bp"], "sharded_s": t1 - t0, "single_gpu_s": t2 - t1}))
dist.barrier()
dist.destroy_process_group()
