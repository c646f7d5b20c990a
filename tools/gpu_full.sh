#!/bin/bash
# Full validation on one GPU: all parity tests, smoke, the three workloads with the automatic kernel choice.
TAG=${1:-r1g}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log
for WL in ctrl4096 plan16384 ctrl1024N100; do
  ST="--steps 30 --warmup 5"; [ $WL != ctrl4096 ] && ST="--steps 3 --warmup 3"
  timeout 900 python bench.py --workload $WL $ST > $OUT/${TAG}_bench_$WL.json 2> $OUT/${TAG}_bench_$WL.err; echo "$WL rc=$?"
  python -c "
import json; d=json.load(open('$OUT/${TAG}_bench_$WL.json')); print('$WL', 'variant', d['config']['kernel_variant'], 'ms', round(d['ms_per_step'],3), 'QP/s', round(d['value']), 'frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']), 'solved', d['solved_fraction'], 'cpu', d.get('cpu_baseline',{}).get('value'))"
  tail -2 $OUT/${TAG}_bench_$WL.err
done
