#!/bin/bash
# usage: gpu_prof.sh <tag> <workload> <batch> [variant]  — one ncu --set full capture of the solve kernel + bench line at that batch
TAG=$1; WL=$2; B=$3; V=${4:-0}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python bench.py --workload $WL --batch $B --variant $V --steps 3 --warmup 3 --no-cpu-baseline --no-saturated > $OUT/${TAG}_bench_${WL}_b$B.json 2> $OUT/${TAG}_prof.err; echo "bench rc=$?"; cut -c1-300 $OUT/${TAG}_bench_${WL}_b$B.json
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:lpv_solve -s 3 -c 1 -f -o $OUT/${TAG}_prof_${WL}_b$B python bench.py --workload $WL --batch $B --variant $V --steps 3 --warmup 3 --no-cpu-baseline --no-saturated > $OUT/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 $OUT/${TAG}_ncu.log | cut -c1-200
ls -la $OUT | grep ${TAG}
