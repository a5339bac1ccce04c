#!/bin/bash
# compute-sanitizer over a small cycle through every kernel family.  gpurun --timeout 700 -- bash tools/gpu_sanitize.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
OUT=gpurun_out/sanitizer_$TAG.txt
: > $OUT
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool python tools/sanitize_cycle.py" >> $OUT
  timeout 200 compute-sanitizer --tool $tool python tools/sanitize_cycle.py 2>&1 | grep -v "^=========$" | tail -6 >> $OUT
  echo "rc=$?" >> $OUT
done
cat $OUT
