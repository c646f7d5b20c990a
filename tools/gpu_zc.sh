#!/bin/bash
# A/B of the host API's copy modes at ctrl4096: results by D2H copy vs written by the kernel into the pinned arena; inputs likewise
TAG=${1:-zc}
OUT=gpurun_out; mkdir -p $OUT
for MODE in "0 0" "1 0" "1 1" "1 0" "0 0"; do
  set -- $MODE
  LPVMPC_ZERO_COPY_OUT=$1 LPVMPC_ZERO_COPY_IN=$2 timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-saturated > $OUT/${TAG}_o$1i$2.json 2> $OUT/${TAG}_o$1i$2.err
  python -c "
import json; d=json.load(open('$OUT/${TAG}_o$1i$2.json')); e=d['e2e']; print('out=$1 in=$2 kernel ms', round(d['ms_per_step'],4), 'e2e ms', round(e['ms_per_step'],4), 'p50', round(e['latency_ms']['p50'],4), 'p99', round(e['latency_ms']['p99'],4), 'e2e QP/s', round(e['value']))"
done
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py -x -q 2>&1 | tail -3
