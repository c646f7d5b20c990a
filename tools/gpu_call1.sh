#!/bin/bash
# GPU call 1 of round 2: parity tests, full-size dumps, sanitizers on the reduced subset, default bench (with the configs block).
TAG=${1:-r3a}
OUT=gpurun_out; mkdir -p $OUT
python -c "import lpvmpc_b200 as lp; print('stale:', lp._native.is_stale())"
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/${TAG}_pytest.log
timeout 600 python tools/dump_fullsize.py $TAG > $OUT/${TAG}_dump.log 2>&1; echo "dump rc=$?"; tail -2 $OUT/${TAG}_dump.log
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-400 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
for TOOL in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $TOOL --print-limit 40 python tools/sanitize_subset.py > $OUT/${TAG}_sanitizer_$TOOL.txt 2>&1
  echo "$TOOL rc=$?"; tail -4 $OUT/${TAG}_sanitizer_$TOOL.txt
done
