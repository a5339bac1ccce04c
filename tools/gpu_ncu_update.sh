#!/bin/bash
# ncu: the profile update kernel on the converged cloud (full set, source on).  gpurun --timeout 900 -- bash tools/gpu_ncu_update.sh <tag>
TAG=${1:-n}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_update_mbes_profile -s 2 -c 1 -f -o gpurun_out/prof_mbes_$TAG \
    python bench.py --cloud converged --steps 2 > gpurun_out/ncu_mbes_$TAG.log 2>&1
tail -3 gpurun_out/ncu_mbes_$TAG.log | cut -c1-300
