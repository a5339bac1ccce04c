#!/bin/bash
TAG=${1:-r2e}
mkdir -p gpurun_out
timeout 120 python tools/debug_peers.py 2 5000 > gpurun_out/debug_peers_$TAG.log 2>&1; echo "debug rc=$?"; tail -15 gpurun_out/debug_peers_$TAG.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_parity_$TAG.log 2>&1; echo "pytest parity rc=$?"; tail -5 gpurun_out/pytest_parity_$TAG.log
bash tools/gpu_bench_only.sh $TAG
timeout 200 python tools/bench_resample.py --sizes 1048576,2097152,16777216 > gpurun_out/resample_sweep_$TAG.jsonl 2>&1; cat gpurun_out/resample_sweep_$TAG.jsonl | cut -c1-260
