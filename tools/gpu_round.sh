#!/bin/bash
# One GPU call: parity tests, smoke, bench of every workload (+ reference arm), ncu launch list, one ncu --set full capture.
# usage (under gpurun): bash tools/gpu_round.sh <tag>
TAG=${1:-r1}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $OUT/${TAG}_clocks.csv 2>/dev/null &
SMI=$!
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log; tail -3 $OUT/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log
timeout 600 python bench.py > $OUT/${TAG}_bench_ctrl4096.json 2> $OUT/${TAG}_bench_ctrl4096.err; echo "bench rc=$?"; cut -c1-600 $OUT/${TAG}_bench_ctrl4096.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/${TAG}_benchref_ctrl4096.json 2>> $OUT/${TAG}_bench_ctrl4096.err; cut -c1-300 $OUT/${TAG}_benchref_ctrl4096.json
for WL in plan16384 ctrl1024N100; do
  timeout 900 python bench.py --workload $WL --steps 3 --warmup 3 > $OUT/${TAG}_bench_$WL.json 2> $OUT/${TAG}_bench_$WL.err; echo "$WL rc=$?"; cut -c1-200 $OUT/${TAG}_bench_$WL.json
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_ctrl4096.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-saturated > $OUT/${TAG}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lpv_solve -s 3 -c 1 -f -o $OUT/${TAG}_prof_ctrl4096 \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-saturated > $OUT/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
kill $SMI
ls -la $OUT | tail -15
# closed-loop fleet (configs[3]) bench + its reference arm
timeout 600 python bench.py --workload mc8192 --steps 24 --warmup 3 > $OUT/${TAG}_bench_mc8192.json 2> $OUT/${TAG}_bench_mc8192.err; echo "mc8192 rc=$?"; cut -c1-200 $OUT/${TAG}_bench_mc8192.json
timeout 300 python bench.py --impl reference --workload mc8192 --steps 5 --warmup 2 > $OUT/${TAG}_benchref_mc8192.json 2>> $OUT/${TAG}_bench_mc8192.err
ls -la $OUT | grep ${TAG} | wc -l
for WL in ctrl512N160 planloop4096; do
  timeout 600 python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_$WL.json 2> $OUT/${TAG}_bench_$WL.err; echo "$WL rc=$?"; cut -c1-160 $OUT/${TAG}_bench_$WL.json
done
