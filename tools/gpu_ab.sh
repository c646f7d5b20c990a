#!/bin/bash
# quick check of a controller-kernel change: all GPU tests, then ctrl4096 (3 runs) and the fleet
TAG=${1:-ab}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest.log
for i in 1 2 3; do
  timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_ctrl4096_$i.json 2> $OUT/${TAG}_bench.err
  python -c "
import json; d=json.load(open('$OUT/${TAG}_bench_ctrl4096_$i.json')); print('ctrl4096 ms', round(d['ms_per_step'],4), 'QP/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'sat', round(d['saturated']['value']), 'iters', d['iters']['mean'])"
done
timeout 300 python bench.py --workload mc8192 --steps 24 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_mc8192.json 2>> $OUT/${TAG}_bench.err; cut -c1-160 $OUT/${TAG}_bench_mc8192.json
