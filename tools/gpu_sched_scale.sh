#!/bin/bash
# how the scheduling kernel's time scales with the batch (latency- or throughput-bound?)
OUT=gpurun_out
for M in 2 1; do for B in 8192 16384 32768 65536 131072 262144; do
  LPVMPC_SCHED_MODE=$M timeout 300 python bench.py --workload sched65536 --batch $B --steps 20 --warmup 5 --no-cpu-baseline > $OUT/tmp_sched.json 2>>$OUT/tmp.err
  python -c "
import json; d=json.loads(open('$OUT/tmp_sched.json').read().strip().split(chr(10))[-1]); print('mode $M B $B', round(d['kernel_latency_ms']['p50']*1e3,1), 'us', round(d['roofline']['achieved']), 'GB/s')"
done; done
