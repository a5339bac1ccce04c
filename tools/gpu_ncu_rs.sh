#!/bin/bash
# ncu full capture of the resample kernels (one instance each).  gpurun -- bash tools/gpu_ncu_rs.sh <tag> [workload]
TAG=$1; WL=${2:-C3}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_normalise|k_quotients|k_select_gather" -s 9 -c 3 -f -o gpurun_out/prof_rs_$TAG \
    python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_rs_$TAG.log 2>&1
tail -2 gpurun_out/ncu_rs_$TAG.log | cut -c1-200
python bench.py --workload $WL --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_rs_$TAG.log 2>&1
python -c "
import json,sys
f='gpurun_out/bench_rs_$TAG.log'
l=[x for x in open(f) if x.startswith('{')]
d=json.loads(l[-1]); print('$WL', 'ms/step %.3f'%d['ms_per_step'], 'evals/s %.3e'%d['value'], {k:round(v,4) for k,v in d['kernel_ms_per_cycle'].items()}, 'rs_frac %.3f'%d['roofline_resample']['frac'])
"
