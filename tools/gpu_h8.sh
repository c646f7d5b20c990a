#!/bin/bash
# usage: gpu_h8.sh <tag> <variant> [plan]
TAG=${1:-r1d}; V=${2:-5}
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python tools/debug_h8.py $V 2>&1 | tail -6
timeout 1200 python -m pytest tests -m gpu -x -q -k "$V or schedule" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/${TAG}_pytest.log
timeout 300 python bench.py --workload ctrl4096 --variant $V --no-cpu-baseline > $OUT/${TAG}_bench_ctrl4096_v$V.json 2> $OUT/${TAG}_bench_v$V.err; echo "v$V rc=$?"; python -c "
import json,sys; d=json.load(open('$OUT/${TAG}_bench_ctrl4096_v$V.json')); print('ctrl4096 ms', d['ms_per_step'], 'QP/s', d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'solved', d['solved_fraction'], d['iters'])"; tail -3 $OUT/${TAG}_bench_v$V.err
if [ "$3" == "plan" ]; then
timeout 600 python bench.py --workload plan16384 --variant $V --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_plan16384_v$V.json 2> $OUT/${TAG}_bench_plan_v$V.err; echo "plan rc=$?"; python -c "
import json,sys; d=json.load(open('$OUT/${TAG}_bench_plan16384_v$V.json')); print('plan16384 ms', d['ms_per_step'], 'QP/s', d['value'], 'frac', d['roofline']['frac'], 'solved', d['solved_fraction'], d['iters'])"; tail -3 $OUT/${TAG}_bench_plan_v$V.err
fi
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lpv_solve -s 3 -c 1 -f -o $OUT/${TAG}_prof_v$V python bench.py --workload ctrl4096 --variant $V --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_v$V.log 2>&1; tail -1 $OUT/${TAG}_ncu_v$V.log | cut -c1-100
