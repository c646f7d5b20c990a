#!/bin/bash
# ncu --set full of the stand-alone scheduling kernel + a clean bench line
TAG=${1:-r3h}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lpv_schedule -s 3 -c 1 -f -o $OUT/${TAG}_prof_sched65536 \
    python bench.py --workload sched65536 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_sched.log 2>&1; echo "ncu rc=$?"
timeout 600 python bench.py --workload sched65536 --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_sched65536.json 2> $OUT/${TAG}_sched.err; cut -c1-200 $OUT/${TAG}_bench_sched65536.json
python -c "
import json; d=json.loads(open('$OUT/${TAG}_bench_sched65536.json').read().strip().split(chr(10))[-1]); print(d['ms_per_step'], d['roofline']['frac'], d['kernel_latency_ms'])"
