#!/bin/bash
OUT=gpurun_out; TAG=${1:-r2d}
run() { name=$1; shift; env "$@" timeout 300 python bench.py --variant 5 --no-cpu-baseline > $OUT/${TAG}_$name.json 2> $OUT/${TAG}_$name.err; python -c "
import json; d=json.load(open('$OUT/${TAG}_$name.json')); print('$name ms', round(d['ms_per_step'],3), 'QP/s', round(d['value']), 'sat', round(d.get('saturated',{}).get('value',0)), 'smem/qp', d['config']['smem_bytes_per_qp'])"; tail -1 $OUT/${TAG}_$name.err; }
run v5_default A=1
run v5_q1 LPVMPC_H8_QPW=1
run v5_q1_c8 LPVMPC_H8_QPW=1 LPVMPC_H8_CTAS=8
run v5_q2 LPVMPC_H8_QPW=2
