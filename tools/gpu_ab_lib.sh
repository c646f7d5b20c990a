#!/bin/bash
# A/B of two builds of the library on the SAME box (boxes differ by a few per cent): usage gpu_ab_lib.sh <alt.so>
ALT=$1; MAIN=autonomous-racing-lpv-mpp-mpc_b200/liblpvmpc.so
cp $MAIN /tmp/main.so
for rep in 1 2; do for which in main alt; do
  if [ $which = alt ]; then cp $ALT $MAIN; else cp /tmp/main.so $MAIN; fi
  python bench.py --no-cpu-baseline --no-saturated --no-configs --steps 40 | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print('$which', round(d['ms_per_step'],4), 'kernel p50', round(d['kernel_latency_ms']['p50'],4), 'e2e', round(d['e2e']['ms_per_step'],4))"
done; done
cp /tmp/main.so $MAIN
