#!/bin/bash
# Compare kernel-tuning builds: gpurun -- bash tools/gpu_variants.sh <tag> lib1.so lib2.so ...   (paths relative to smarc_navigation_b200/)
TAG=$1; shift
mkdir -p gpurun_out
if [ -z "$NOTEST" ]; then python -m pytest tests -m gpu -x -q -k "raycast or logw or c3_full" > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest(default lib) rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log; fi
for lib in "$@"; do
  for wl in C3 C2; do
  AUVPF_LIB=$PWD/smarc_navigation_b200/$lib python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_${lib}_$wl.log 2>&1
  python -c "
import json,sys
f='gpurun_out/bench_${TAG}_${lib}_$wl.log'
l=[x for x in open(f) if x.startswith('{')]
if not l: print(open(f).read()[-1500:]); sys.exit()
d=json.loads(l[-1]); print('$lib $wl', 'ms/step %.3f'%d['ms_per_step'], 'evals/s %.3e'%d['value'], {k:round(v,4) for k,v in d['kernel_ms_per_cycle'].items()}, 'err %.4f'%d['track_error_m'])
"
  done
done
