#!/bin/bash
# compute-sanitizer over the reduced subset, one process per case (a launch failure is sticky): usage gpu_sanitize.sh <tag> [cases...]
TAG=${1:-r3u}; shift
CASES=${@:-v8 v5n100 p5 p5n39 schedule aux fleet v6 v5 v1 v7}
OUT=gpurun_out; mkdir -p $OUT
for TOOL in memcheck racecheck synccheck; do
  : > $OUT/${TAG}_sanitizer_$TOOL.txt
  for C in $CASES; do
    echo "##### case $C" >> $OUT/${TAG}_sanitizer_$TOOL.txt
    timeout 400 compute-sanitizer --tool $TOOL --print-limit 8 python tools/sanitize_subset.py $C >> $OUT/${TAG}_sanitizer_$TOOL.txt 2>&1
    echo "$TOOL $C rc=$?"
  done
  grep -E "#####|ERROR SUMMARY|RACECHECK SUMMARY|ok" $OUT/${TAG}_sanitizer_$TOOL.txt | paste - - - | head -20
done
