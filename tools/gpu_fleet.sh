#!/bin/bash
# Fleet (configs[3]) validation on one GPU: parity test, smoke, mc8192 bench (+ reference arm), ncu launch list.
TAG=${1:-r1j}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_fleet.py -x -q > $OUT/${TAG}_pytest_fleet.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/${TAG}_pytest_fleet.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/${TAG}_smoke.log
timeout 600 python bench.py --workload mc8192 --steps 24 --warmup 3 > $OUT/${TAG}_bench_mc8192.json 2> $OUT/${TAG}_bench_mc8192.err; echo "bench rc=$?"; cut -c1-1500 $OUT/${TAG}_bench_mc8192.json; tail -3 $OUT/${TAG}_bench_mc8192.err
timeout 300 python bench.py --impl reference --workload mc8192 --steps 5 --warmup 2 > $OUT/${TAG}_benchref_mc8192.json 2>> $OUT/${TAG}_bench_mc8192.err; cut -c1-300 $OUT/${TAG}_benchref_mc8192.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/${TAG}_launches_mc8192.csv \
    python bench.py --workload mc8192 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_launch.log 2>&1; echo "ncu rc=$?"
