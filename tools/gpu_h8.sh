#!/bin/bash
TAG=${1:-r1c}
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q -k "5 or schedule" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/${TAG}_pytest.log
timeout 300 python bench.py --workload ctrl4096 --variant 5 --no-cpu-baseline > $OUT/${TAG}_bench_ctrl4096_v5.json 2> $OUT/${TAG}_bench_v5.err; echo "v5 rc=$?"; cat $OUT/${TAG}_bench_ctrl4096_v5.json; tail -3 $OUT/${TAG}_bench_v5.err
timeout 600 python bench.py --workload plan16384 --variant 5 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_plan16384_v5.json 2> $OUT/${TAG}_bench_plan_v5.err; echo "plan rc=$?"; cat $OUT/${TAG}_bench_plan16384_v5.json; tail -3 $OUT/${TAG}_bench_plan_v5.err
