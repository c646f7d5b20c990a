#!/bin/bash
# A/B of several builds of the library on the SAME box: usage gpu_ab_multi.sh "<workloads>" <alt1.so> [<alt2.so> ...]
# (ctrl4096 is timed with the default bench line, the other workloads with --workload)
WLS=$1; shift; MAIN=autonomous-racing-lpv-mpp-mpc_b200/liblpvmpc.so
cp $MAIN /tmp/main.so
for rep in 1 2; do for wl in $WLS; do for lib in /tmp/main.so "$@"; do
  cp $lib $MAIN
  if [ $wl = ctrl4096 ]; then ARGS="--steps 40"; else ARGS="--workload $wl --steps 3 --warmup 3"; fi
  timeout 300 python bench.py --no-cpu-baseline --no-saturated --no-configs $ARGS 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print('$wl', '$(basename $lib)', round(d['ms_per_step'],4), 'ms')"
done; done; done
cp /tmp/main.so $MAIN
