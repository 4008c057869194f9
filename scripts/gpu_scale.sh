#!/bin/bash
# bench.py at N GPUs of one box (the driver's launch line), steps 2 warmup 1
set -u
N=${1:-4}; TAG=${2:-r02v}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 2 --warmup 1 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err; echo "bench n$N rc=$?"
tail -1 $OUT/${TAG}_bench_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.1f e2e %.1f Mbp/s n=%d'%(d['value']/1e6,d['e2e']['value']/1e6,d['n_gpus']), d['run']['records_per_gpu'], {k:round(v,1) for k,v in d['phases_ms'].items()}, d['parity']['paf_lines_identical'], d['parity']['mapping_lines_identical'])
"
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" $OUT/${TAG}_bench_n$N.err | tail -5
