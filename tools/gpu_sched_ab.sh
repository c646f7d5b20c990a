#!/bin/bash
# A/B of the stand-alone scheduling kernel's three output paths (LPVMPC_SCHED_MODE 2 = TMA tensor stores, 1 = 256-bit stores, 0 = tile-staged)
TAG=${1:-r3k}
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -q -m gpu -k "schedule" 2>&1 | tail -3
for M in 2 1 0; do
  LPVMPC_SCHED_MODE=$M timeout 300 python bench.py --workload sched65536 --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_sched_mode$M.json 2>>$OUT/${TAG}.err
  python -c "
import json; d=json.loads(open('$OUT/${TAG}_sched_mode$M.json').read().strip().split(chr(10))[-1]); print('mode $M', d['ms_per_step'], d['roofline']['frac'], d['kernel_latency_ms'])"
done
tail -3 $OUT/${TAG}.err
