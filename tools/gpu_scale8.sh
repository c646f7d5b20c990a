#!/bin/bash
# 8-GPU weak-scaling lines (one process per GPU, no collective on the data path): configs[1], [3] (8 x 8,192 = 65,536 vehicles), [2]
TAG=${1:-r2l}; NG=${2:-8}
OUT=gpurun_out; mkdir -p $OUT
run() { WL=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --workload $WL --no-cpu-baseline "$@" \
    > $OUT/${TAG}_bench_${WL}_g$NG.json 2> $OUT/${TAG}_bench_${WL}_g$NG.err; echo "$WL g$NG rc=$?"; grep -h '^{' $OUT/${TAG}_bench_${WL}_g$NG.json | cut -c1-220; }
run ctrl4096 --steps 30 --warmup 5
run mc8192 --steps 24 --warmup 3
run plan16384 --steps 3 --warmup 3
