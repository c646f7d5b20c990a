#!/bin/bash
TAG=${1:-r2g}; OUT=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_fleet.py -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py > $OUT/${TAG}_bench_ctrl4096.json 2> $OUT/${TAG}_err.txt; python -c "
import json; d=json.load(open('$OUT/${TAG}_bench_ctrl4096.json')); print('ctrl4096 ms', round(d['ms_per_step'],3), 'QP/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],4), 'sat', round(d['saturated']['value']), 'launches', d['gpu_launches'])"
timeout 300 python bench.py --workload mc8192 --steps 12 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_mc8192.json 2>> $OUT/${TAG}_err.txt; python -c "
import json; d=json.load(open('$OUT/${TAG}_bench_mc8192.json')); print('mc8192 QP/s', round(d['value']), 'ms/tick', d['kernel_latency_ms'])"
timeout 300 python bench.py --workload ctrl1024N100 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_n100.json 2>> $OUT/${TAG}_err.txt; python -c "
import json; d=json.load(open('$OUT/${TAG}_bench_n100.json')); print('n100 ms', round(d['ms_per_step'],2), 'QP/s', round(d['value']))"
tail -2 $OUT/${TAG}_err.txt
