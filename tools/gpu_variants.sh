#!/bin/bash
# Compare kernel variants on one GPU: parity tests, then bench per variant.
TAG=${1:-r1b}
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/${TAG}_pytest.log
for v in 2 3; do
  timeout 300 python bench.py --workload ctrl4096 --variant $v --no-cpu-baseline > $OUT/${TAG}_bench_ctrl4096_v$v.json 2> $OUT/${TAG}_bench_v$v.err; echo "v$v rc=$?"; cat $OUT/${TAG}_bench_ctrl4096_v$v.json
done
timeout 600 python bench.py --workload plan16384 --variant 3 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_plan16384_v3.json 2> $OUT/${TAG}_bench_plan_v3.err; echo "plan rc=$?"; cat $OUT/${TAG}_bench_plan16384_v3.json
