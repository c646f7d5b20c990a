#!/bin/bash
# A/B of environment switches on the SAME box: usage gpu_ab_env.sh "VAR=a" "VAR=b" [bench args...]
A=$1; B=$2; shift 2
for rep in 1 2; do for which in "$A" "$B"; do
  env $which python bench.py --no-cpu-baseline --no-saturated --no-configs --steps 40 "$@" | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print('$which', round(d['ms_per_step'],4), 'kernel p50', round(d['kernel_latency_ms']['p50'],4), 'e2e', round(d['e2e']['ms_per_step'],4))"
done; done
