#!/bin/bash
# ncu --set full of the H16T kernel (variant 8) on ctrl4096, source-level sampling
TAG=${1:-r3c}; VAR=${2:-8}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lpv_solve -s 3 -c 1 -f -o $OUT/${TAG}_prof_v${VAR} \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-saturated --no-configs --variant $VAR > $OUT/${TAG}_ncu_v${VAR}.log 2>&1; echo "ncu rc=$?"
ls -la $OUT/${TAG}_prof_v${VAR}.ncu-rep
