#!/bin/bash
# Round-2 pass F: tests (parity + sharded on one GPU), bench, resample sweep with variants.
TAG=${1:-r2f}
mkdir -p gpurun_out
rm -f gpurun_out/fragile_report.jsonl
timeout 120 python tools/debug_peers.py 4 70001 > gpurun_out/debug_peers_$TAG.log 2>&1; echo "debug rc=$?"; tail -4 gpurun_out/debug_peers_$TAG.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_$TAG.log
bash tools/gpu_bench_only.sh $TAG
echo "--- sweep default"; timeout 200 python tools/bench_resample.py --sizes 1048576,2097152,16777216 > gpurun_out/resample_sweep_$TAG.jsonl 2>&1; cut -c1-420 gpurun_out/resample_sweep_$TAG.jsonl
echo "--- sweep no PDL"; AUVPF_NO_PDL=1 timeout 200 python tools/bench_resample.py --sizes 1048576,2097152 2>&1 | cut -c1-130
echo "--- sweep MINB=3"; AUVPF_LIB=$PWD/smarc_navigation_b200/variants/libauvpf_minb3.so timeout 200 python tools/bench_resample.py --sizes 1048576,2097152,16777216 2>&1 | cut -c1-130
