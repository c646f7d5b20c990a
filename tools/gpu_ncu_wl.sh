#!/bin/bash
# ncu --set full of the solve kernel on a (reduced) workload: usage gpu_ncu_wl.sh <tag> <workload> <batch> [variant]
TAG=${1:-r3f}; WL=${2:-plan16384}; B=${3:-888}; VAR=${4:-0}
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:lpv_solve -s 3 -c 1 -f -o $OUT/${TAG}_prof_${WL}_b${B} \
    python bench.py --workload $WL --batch $B --steps 3 --warmup 3 --no-cpu-baseline --no-saturated --no-configs --variant $VAR > $OUT/${TAG}_ncu_${WL}_b${B}.log 2>&1; echo "ncu rc=$?"
ls -la $OUT/${TAG}_prof_${WL}_b${B}.ncu-rep
