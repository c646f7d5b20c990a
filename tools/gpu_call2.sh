#!/bin/bash
# GPU call 2: tests with the new polish, per-case synccheck (one process per case: a launch failure is sticky), the
# tcgen05.alloc cross-check of the synccheck report, bench.
TAG=${1:-r3b}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q -s > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/${TAG}_pytest.log
timeout 600 python tools/dump_fullsize.py $TAG > $OUT/${TAG}_dump.log 2>&1; echo "dump rc=$?"
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-300 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
: > $OUT/${TAG}_sanitizer_synccheck.txt
for C in v1 v5 v6 v7 v5n20 v5n100 v7n160 p1 p5 p7 schedule fleet planfleet; do
  echo "##### case $C" >> $OUT/${TAG}_sanitizer_synccheck.txt
  timeout 300 compute-sanitizer --tool synccheck --print-limit 5 python tools/sanitize_subset.py $C >> $OUT/${TAG}_sanitizer_synccheck.txt 2>&1
  echo "synccheck $C rc=$?"
done
echo "##### tmem_probe (tcgen05.alloc / ld / st only, no mbarrier anywhere)" >> $OUT/${TAG}_sanitizer_synccheck.txt
timeout 120 compute-sanitizer --tool synccheck --print-limit 5 tools/tmem_probe >> $OUT/${TAG}_sanitizer_synccheck.txt 2>&1; echo "tmem_probe synccheck rc=$?"
grep -E "#####|ERROR SUMMARY|Barrier error|ok" $OUT/${TAG}_sanitizer_synccheck.txt | head -60
