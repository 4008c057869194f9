#!/bin/bash
# N = 2: ONE scerevisiae8 job partitioned over two GPUs (strong scaling), text checked against the single-GPU fixture
set -u
OUT=gpurun_out
TAG=${1:-r02g}
mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 > $OUT/${TAG}_bench_n2.json 2> $OUT/${TAG}_bench_n2.err; echo "bench n2 rc=$?"
tail -1 $OUT/${TAG}_bench_n2.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.1f e2e %.1f Mbp/s n=%d'%(d['value']/1e6,d['e2e']['value']/1e6,d['n_gpus']), d['run']['records_per_gpu'], {k:round(v,1) for k,v in d['phases_ms'].items()}, d['parity'])
"
tail -5 $OUT/${TAG}_bench_n2.err

