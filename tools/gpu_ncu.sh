#!/bin/bash
# ncu full capture of one kernel inside a short bench run.  gpurun --timeout 900 -- bash tools/gpu_ncu.sh <tag> <kernel regex> [skip] [workload]
TAG=$1; KRE=$2; SKIP=${3:-4}; WL=${4:-C3}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s $SKIP -c 1 -f -o gpurun_out/prof_$TAG \
    python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log | cut -c1-300
