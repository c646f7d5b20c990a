#!/bin/bash
# H8S (streamed factor) validation: parity tests of variants 7/8 + long horizon, then planner / N=100 benches per variant
TAG=${1:-r1l}
OUT=gpurun_out; mkdir -p $OUT
timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -k "fixed_iteration and (7 or 8) and not 200" > $OUT/${TAG}_pytest_a.log 2>&1; echo "pytest-a rc=$?"; tail -5 $OUT/${TAG}_pytest_a.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "7 or 8 or long_horizon" > $OUT/${TAG}_pytest_b.log 2>&1; echo "pytest-b rc=$?"; tail -8 $OUT/${TAG}_pytest_b.log
for V in 5 7 8; do
  timeout 300 python bench.py --workload plan16384 --variant $V --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_plan16384_v$V.json 2> $OUT/${TAG}_bench_plan_v$V.err; echo "plan v$V rc=$?"
  python -c "
import json; d=json.load(open('$OUT/${TAG}_bench_plan16384_v$V.json')); print('plan16384 v$V ms', d['ms_per_step'], 'QP/s', d['value'], 'frac', d['roofline']['frac'], 'solved', d['solved_fraction'], d['iters'], d['config']['smem_bytes_per_qp'])"; tail -2 $OUT/${TAG}_bench_plan_v$V.err
  timeout 300 python bench.py --workload ctrl1024N100 --variant $V --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_ctrl1024N100_v$V.json 2> $OUT/${TAG}_bench_n100_v$V.err; echo "n100 v$V rc=$?"
  python -c "
import json; d=json.load(open('$OUT/${TAG}_bench_ctrl1024N100_v$V.json')); print('ctrl1024N100 v$V ms', d['ms_per_step'], 'QP/s', d['value'], 'frac', d['roofline']['frac'], 'solved', d['solved_fraction'], d['iters'], d['config']['smem_bytes_per_qp'])"; tail -2 $OUT/${TAG}_bench_n100_v$V.err
done
