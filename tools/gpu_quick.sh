#!/bin/bash
# quick A/B of the H8 family: parity subset + planner / N=100 benches for the variants given
TAG=${1:-q}; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "planner or long_horizon or (converged and (5 or 7))" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/${TAG}_pytest.log
for V in "$@"; do
  for WL in plan16384 ctrl1024N100; do
    timeout 300 python bench.py --workload $WL --variant $V --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_${WL}_v$V.json 2> $OUT/${TAG}_bench_${WL}_v$V.err
    python -c "
import json; d=json.load(open('$OUT/${TAG}_bench_${WL}_v$V.json')); print('$WL v$V ms', round(d['ms_per_step'],1), 'QP/s', round(d['value']), 'frac', round(d['roofline']['frac'],4), 'solved', d['solved_fraction'])"; tail -1 $OUT/${TAG}_bench_${WL}_v$V.err
  done
done
