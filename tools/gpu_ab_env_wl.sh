#!/bin/bash
# A/B of environment switches on the SAME box for one workload: usage gpu_ab_env_wl.sh <workload> "VAR=a" "VAR=b" ...
WL=$1; shift
for rep in 1 2; do for which in "$@"; do
  env $which timeout 300 python bench.py --workload $WL --no-cpu-baseline --no-saturated --no-configs --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print('$WL $which', round(d['ms_per_step'],3), 'ms  solved', d.get('solved_fraction'), 'iters', d.get('iters',{}).get('mean'))"
done; done
