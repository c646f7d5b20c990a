#!/bin/bash
# stand-alone scheduling kernel: parity tests that use it, then the HBM-roofline bench line (tiled kernel and the naive one), ncu launch list
TAG=${1:-sch}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py -x -q -k "schedule or sched or copy_modes or estimate" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest.log
timeout 300 python bench.py --workload sched65536 --steps 20 --warmup 3 > $OUT/${TAG}_bench_sched65536.json 2> $OUT/${TAG}_bench.err; echo "rc=$?"; tail -2 $OUT/${TAG}_bench.err
LPVMPC_SCHED_NAIVE=1 timeout 300 python bench.py --workload sched65536 --steps 20 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_sched65536_naive.json 2>> $OUT/${TAG}_bench.err
python -c "
import json
for f in ('', '_naive'):
    d=json.load(open('$OUT/${TAG}_bench_sched65536%s.json' % f)); r=d['roofline']; print(d['config']['kernel'], 'ms', round(d['ms_per_step'],4), 'GB/s', round(r['achieved'],1), 'frac', round(r['frac'],3), 'e2e QP/s', round(d['e2e']['value']), d.get('cpu_baseline',{}).get('value'))"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lpv_schedule_kernel -s 2 -c 1 -f -o $OUT/${TAG}_prof_sched65536 \
    python bench.py --workload sched65536 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1; echo "ncu rc=$?"
