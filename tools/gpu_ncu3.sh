#!/bin/bash
# ncu of the two resample kernels (1 M and 16 M) + the sharded tests on one GPU.  gpurun --timeout 900 -- bash tools/gpu_ncu3.sh <tag>
TAG=${1:-n}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/pytest_sharded1_$TAG.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_sharded1_$TAG.log
ncu --set full --clock-control none --import-source on -k regex:"k_weights|k_select" -s 6 -c 4 -f -o gpurun_out/prof_rs_$TAG \
    python tools/bench_resample.py --sizes 1048576,16777216 --reps 2 > gpurun_out/ncu_rs_$TAG.log 2>&1
tail -2 gpurun_out/ncu_rs_$TAG.log | cut -c1-200
