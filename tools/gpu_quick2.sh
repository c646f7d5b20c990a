#!/bin/bash
TAG=${1:-q}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "planner or long_horizon" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/${TAG}_pytest.log
for WL in plan16384 ctrl1024N100 ctrl512N160; do
  timeout 300 python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_${WL}.json 2> $OUT/${TAG}_bench_${WL}.err
  python -c "
import json; d=json.load(open('$OUT/${TAG}_bench_${WL}.json')); print('$WL v', d['config']['kernel_variant'], 'ms', round(d['ms_per_step'],1), 'QP/s', round(d['value']), 'frac', round(d['roofline']['frac'],4), 'solved', d['solved_fraction'], d['iters'])"; tail -1 $OUT/${TAG}_bench_${WL}.err
done
