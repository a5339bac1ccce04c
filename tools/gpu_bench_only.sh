#!/bin/bash
# bench C3 only.  gpurun --timeout 600 -- bash tools/gpu_bench_only.sh <tag> [extra bench args]
TAG=${1:-q}; shift
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_c3_$TAG.log 2>&1; echo "bench rc=$?"
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_c3_$TAG.log') if x.startswith('{')]
if not l:
    print(open('gpurun_out/bench_c3_$TAG.log').read()[-3000:])
else:
    d=json.loads(l[-1])
    print('ms/step %.4f value %.4e e2e %.4e parity %s'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['parity_check']))
    print({k:round(v,4) for k,v in d['kernel_ms_per_cycle'].items()}, 'rs frac %.3f'%d['roofline_resample']['frac'])
    for k,v in d.get('clouds',{}).items(): print(k, 'update %.3f ms  %.3e evals/s  flagged %.4f  resample %.4f ms  vs conv %.3f'%(v['update_ms'],v['evals_per_s'],v['flagged_particle_frac'],v['resample_ms'],v['vs_converged']))
PY
