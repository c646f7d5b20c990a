#!/bin/bash
# multi-GPU lines as the driver launches them (one rank per GPU, torchrun, NCCL): usage gpu_scale8.sh <tag> <n_gpus>
TAG=${1:-r3s}; N=${2:-8}
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 \
   > $OUT/${TAG}_bench_g$N.json 2> $OUT/${TAG}_bench_g$N.err; echo "rc=$?"; cut -c1-300 $OUT/${TAG}_bench_g$N.json; tail -3 $OUT/${TAG}_bench_g$N.err
