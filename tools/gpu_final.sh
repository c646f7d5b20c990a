#!/bin/bash
# Final round of a session: GPU tests, smoke, default bench (+ configs), reference arm, ncu launch list of the bench command,
# full-size parity dump.  usage (under gpurun): bash tools/gpu_final.sh <tag>
TAG=${1:-r3w}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $OUT/${TAG}_clocks.csv 2>/dev/null &
SMI=$!
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/${TAG}_smoke.log
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-300 $OUT/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/${TAG}_benchref.json 2>> $OUT/${TAG}_bench.err; cut -c1-300 $OUT/${TAG}_benchref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_ctrl4096.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-saturated --no-configs > $OUT/${TAG}_ncu_launch.log 2>&1; echo "launch list rc=$?"
kill $SMI
timeout 600 python tools/dump_fullsize.py $TAG > $OUT/${TAG}_dump.log 2>&1; echo "dump rc=$?"; tail -2 $OUT/${TAG}_dump.log
