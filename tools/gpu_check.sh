#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench (C3 default + C2), ncu launch list + full capture of the raycast kernel.
# Run as: gpurun --timeout 1500 -- bash tools/gpu_check.sh [tag]
TAG=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$TAG.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c3_$TAG.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_c3_$TAG.log | cut -c1-600
python bench.py --workload C2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_$TAG.log 2>&1; tail -1 gpurun_out/bench_c2_$TAG.log | cut -c1-400
python bench.py --workload C5 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_$TAG.log 2>&1; tail -1 gpurun_out/bench_c5_$TAG.log | cut -c1-400
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2>&1; tail -1 gpurun_out/bench_ref_$TAG.log | cut -c1-600
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}_c3.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_update_mbes_profile -s 4 -c 1 -f -o gpurun_out/prof_mbes_$TAG \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_normalise|k_quotients|k_select_gather|k_propagate" -s 12 -c 8 -f -o gpurun_out/prof_resample_$TAG \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full2_$TAG.log 2>&1
echo done
