#!/bin/bash
TAG=${1:-r1o}
OUT=gpurun_out; mkdir -p $OUT
run() { # name env... 
  local name=$1; shift
  env "$@" timeout 300 python bench.py --workload plan16384 --variant ${V:-7} --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_plan_$name.json 2> $OUT/${TAG}_plan_$name.err
  python -c "
import json; d=json.load(open('$OUT/${TAG}_plan_$name.json')); print('$name', 'ms', round(d['ms_per_step'],1), 'QP/s', round(d['value']), 'smem/qp', d['config']['smem_bytes_per_qp'])"; tail -1 $OUT/${TAG}_plan_$name.err
}
V=7 run v7_q1 A=1
V=7 run v7_q1_c4 LPVMPC_H8_CTAS=4
V=7 run v7_q1_c5 LPVMPC_H8_CTAS=5
V=7 run v7_q2 LPVMPC_H8_QPW=2
V=7 run v7_q4 LPVMPC_H8_QPW=4
V=5 run v5_q1_c2 LPVMPC_H8_CTAS=2
V=5 run v5_q1_c1 LPVMPC_H8_CTAS=1
# one round only: 888 QPs -> latency of the longest chain
timeout 300 python bench.py --workload plan16384 --variant 7 --batch 888 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_plan_b888.json 2>/dev/null; python -c "
import json; d=json.load(open('$OUT/${TAG}_plan_b888.json')); print('b888 ms', d['ms_per_step'], d['iters'])"
timeout 300 python bench.py --workload plan16384 --variant 7 --batch 148 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_plan_b148.json 2>/dev/null; python -c "
import json; d=json.load(open('$OUT/${TAG}_plan_b148.json')); print('b148 ms', d['ms_per_step'], d['iters'])"
