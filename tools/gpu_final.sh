#!/bin/bash
# Short evidence pass after a resample-kernel change: parity tests, smoke, bench C3/C2/C5, resample sweep,
# ncu launch list + full capture of the resample kernels.  gpurun --timeout 600 -- bash tools/gpu_final.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_$TAG.log
python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$TAG.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c3_$TAG.log 2>&1; echo "bench rc=$?"
python bench.py --workload C2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_$TAG.log 2>&1
python bench.py --workload C5 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_$TAG.log 2>&1
timeout 100 python tools/bench_resample.py --reps 10 > gpurun_out/rs_sweep_$TAG.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}_c3.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_normalise|k_quotients|k_select_gather|k_propagate" -s 12 -c 4 -f -o gpurun_out/prof_resample_$TAG \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full2_$TAG.log 2>&1
for w in c3 c2 c5; do python -c "
import json
l=[x for x in open('gpurun_out/bench_${w}_$TAG.log') if x.startswith('{')]
d=json.loads(l[-1]); print('$w', 'ms/step %.3f'%d['ms_per_step'], 'evals/s %.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'], 'replay %.3e'%d['e2e_replay']['value'], {k:round(v,4) for k,v in d['kernel_ms_per_cycle'].items()}, 'rs_frac %.3f'%d['roofline_resample']['frac'])
"; done
python -c "
import json
for l in open('gpurun_out/rs_sweep_$TAG.log'):
    if l.startswith('{'):
        d=json.loads(l); print(d['particles'], {k:round(v,4) for k,v in d['ms'].items()}, round(d['frac'],3))
"
