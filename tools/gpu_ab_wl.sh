#!/bin/bash
# A/B of two builds of the library on the SAME box for the one-QP-per-warp workloads: usage gpu_ab_wl.sh <base.so> [workloads...]
BASE=$1; shift; MAIN=autonomous-racing-lpv-mpp-mpc_b200/liblpvmpc.so
WLS=${@:-"ctrl1024N100 plan16384"}
cp $MAIN /tmp/new.so
for wl in $WLS; do for which in base new; do
  if [ $which = base ]; then cp $BASE $MAIN; else cp /tmp/new.so $MAIN; fi
  LPVMPC_SKIP_STALE_CHECK=1 timeout 300 python bench.py --workload $wl --no-cpu-baseline --no-saturated --no-configs --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print('$wl $which', round(d['ms_per_step'],3), 'ms  solved', d.get('solved_fraction'), 'iters', d.get('iters',{}).get('mean'))"
done; done
cp /tmp/new.so $MAIN
