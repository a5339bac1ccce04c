#!/bin/bash
# Quick GPU pass: parity tests + short bench of C3 and C2.  gpurun --timeout 900 -- bash tools/gpu_quick.sh [tag] [pytest -k expr]
TAG=${1:-q}
KEXPR=${2:-}
mkdir -p gpurun_out
if [ -n "$KEXPR" ]; then python -m pytest tests -m gpu -x -q -k "$KEXPR" > gpurun_out/pytest_gpu_$TAG.log 2>&1; else python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; fi
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_$TAG.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_$TAG.log 2>&1; echo "bench rc=$?"
python bench.py --workload C2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_$TAG.log 2>&1
python - <<'PY'
import json,sys,glob,os
tag=os.environ.get("TAG_","")
PY
for f in gpurun_out/bench_c3_$TAG.log gpurun_out/bench_c2_$TAG.log; do python -c "
import json,sys
l=[x for x in open('$f') if x.startswith('{')]
if not l: print(open('$f').read()[-1500:]); sys.exit()
d=json.loads(l[-1]); print('$f', 'ms/step %.3f'%d['ms_per_step'], 'evals/s %.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'], {k:round(v,4) for k,v in d['kernel_ms_per_cycle'].items()}, 'rs_frac %.3f'%d['roofline_resample']['frac'], 'err %.4f'%d['track_error_m'])
"; done
if [ -n "$C5" ]; then python bench.py --workload C5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_$TAG.log 2>&1; python -c "
import json,sys
f='gpurun_out/bench_c5_$TAG.log'
l=[x for x in open(f) if x.startswith('{')]
if not l: print(open(f).read()[-2500:]); sys.exit()
d=json.loads(l[-1]); print(f, 'ms/step %.3f'%d['ms_per_step'], 'evals/s %.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'], {k:round(v,4) for k,v in d['kernel_ms_per_cycle'].items()}, 'rs_frac %.3f'%d['roofline_resample']['frac'], 'err %.4f'%d['track_error_m'])
"; fi
