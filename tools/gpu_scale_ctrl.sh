#!/bin/bash
# one multi-GPU line of configs[1] (weak scaling): usage gpu_scale_ctrl.sh <tag> <ngpus>
TAG=${1:-r2s}; NG=${2:-8}
OUT=gpurun_out; mkdir -p $OUT
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --no-cpu-baseline --steps 60 --warmup 5 \
  > $OUT/${TAG}_bench_ctrl4096_g$NG.json 2> $OUT/${TAG}_bench_ctrl4096_g$NG.err; echo "rc=$?"
python -c "
import json
d=[json.loads(l) for l in open('$OUT/${TAG}_bench_ctrl4096_g$NG.json') if l.startswith('{')][-1]
print('n', d['n_gpus'], 'QP/s', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'e2e ms', round(d['e2e']['ms_per_step'],4))"
