#!/bin/bash
# Select-kernel tuning builds: default library through the full quick pass, then each variant through the
# resample sweep and a C3 bench.  gpurun -- bash tools/gpu_sel_variants.sh <tag> libauvpf_sb5.so ...
TAG=$1; shift
timeout 400 bash tools/gpu_quick.sh $TAG
show() { python -c "
import json,sys
for l in open('$1'):
    if l.startswith('{'):
        d=json.loads(l); print('$2', d['particles'], {k:round(v,4) for k,v in d['ms'].items()}, round(d['frac'],3))
"; }
timeout 120 python tools/bench_resample.py --reps 10 --sizes 65536,1048576,2097152,16777216 > gpurun_out/rs_sweep_$TAG.log 2>&1; show gpurun_out/rs_sweep_$TAG.log default
for lib in "$@"; do
  AUVPF_LIB=$PWD/smarc_navigation_b200/$lib timeout 120 python tools/bench_resample.py --reps 10 --sizes 1048576,16777216 > gpurun_out/rs_sweep_${TAG}_$lib.log 2>&1; show gpurun_out/rs_sweep_${TAG}_$lib.log $lib
  AUVPF_LIB=$PWD/smarc_navigation_b200/$lib timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_${TAG}_$lib.log 2>&1
  python -c "
import json
l=[x for x in open('gpurun_out/bench_c3_${TAG}_$lib.log') if x.startswith('{')]
d=json.loads(l[-1]); print('$lib C3', {k:round(v,4) for k,v in d['kernel_ms_per_cycle'].items()}, 'rs_frac %.3f'%d['roofline_resample']['frac'])
"
done
