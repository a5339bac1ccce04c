#!/bin/bash
# ncu: the update kernel on the converged cloud + the two resample kernels at 1 M and 2 M.  gpurun --timeout 900 -- bash tools/gpu_ncu_all.sh <tag>
TAG=${1:-n}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_update_mbes_profile -s 2 -c 1 -f -o gpurun_out/prof_mbes_$TAG \
    python bench.py --cloud converged --steps 2 > gpurun_out/ncu_mbes_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_weights|k_select" -s 6 -c 4 -f -o gpurun_out/prof_rs_$TAG \
    python tools/bench_resample.py --sizes 1048576,2097152 --reps 3 > gpurun_out/ncu_rs_$TAG.log 2>&1
tail -2 gpurun_out/ncu_rs_$TAG.log | cut -c1-200
