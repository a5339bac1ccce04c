#!/bin/bash
# Kernel-tuning builds against the default library: converged-cloud update time.  gpurun -- bash tools/gpu_variants.sh <tag> lib1.so ...
TAG=$1; shift
mkdir -p gpurun_out
for lib in libauvpf.so "$@"; do
  AUVPF_LIB=$PWD/smarc_navigation_b200/$lib timeout 300 python bench.py --cloud converged --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_${lib}.log 2>&1
  python - <<PY
import json
f='gpurun_out/bench_${TAG}_${lib}.log'
l=[x for x in open(f) if x.startswith('{')]
if not l: print('$lib', open(f).read()[-800:])
else:
    c=json.loads(l[-1])['cloud_only']['converged']
    print('$lib', 'update_ms %.4f'%c['update_ms'], 'resample_ms %.4f'%c['resample_ms'])
PY
done
