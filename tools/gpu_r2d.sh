#!/bin/bash
# Round-2 pass D: the two-kernel resample.  gpurun --timeout 1200 -- bash tools/gpu_r2d.sh <tag>
TAG=${1:-r2d}
mkdir -p gpurun_out
rm -f gpurun_out/fragile_report.jsonl
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_$TAG.log 2>&1; echo "bench rc=$?"
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
timeout 300 python tools/bench_resample.py > gpurun_out/resample_sweep_$TAG.jsonl 2>&1; cat gpurun_out/resample_sweep_$TAG.jsonl | cut -c1-330
