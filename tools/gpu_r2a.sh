#!/bin/bash
# Round-2 pass A: parity tests at HEAD, smoke, bench with the cloud sweep, ncu of the update kernel per cloud shape.
# gpurun --timeout 1500 -- bash tools/gpu_r2a.sh <tag>
TAG=${1:-r2a}
mkdir -p gpurun_out
rm -f gpurun_out/fragile_report.jsonl
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_$TAG.log
cp gpurun_out/fragile_report.jsonl gpurun_out/fragile_report_$TAG.jsonl 2>/dev/null
python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$TAG.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c3_$TAG.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_c3_$TAG.log | cut -c1-300
for C in converged diffuse kidnapped; do
ncu --set full --clock-control none --import-source on -k regex:k_update_mbes_profile -s 2 -c 1 -f -o gpurun_out/prof_mbes_${C}_$TAG \
    python bench.py --cloud $C --steps 2 > gpurun_out/ncu_cloud_${C}_$TAG.log 2>&1
tail -1 gpurun_out/ncu_cloud_${C}_$TAG.log | cut -c1-400
done
echo done
