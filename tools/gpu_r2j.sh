#!/bin/bash
# Round-2 pass J: full GPU test-suite at HEAD, smoke, sanitizer, bench (default), resample sweep.
TAG=${1:-r2j}
mkdir -p gpurun_out
rm -f gpurun_out/fragile_report.jsonl
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_$TAG.log
cp gpurun_out/fragile_report.jsonl gpurun_out/fragile_report_$TAG.jsonl 2>/dev/null
python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$TAG.log
timeout 700 bash tools/gpu_sanitize.sh $TAG | grep -E "==|ERROR SUMMARY|hazard|rc=|ok"
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c3_$TAG.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_c3_$TAG.log | cut -c1-250
timeout 300 python tools/bench_resample.py > gpurun_out/resample_sweep_$TAG.jsonl 2>&1; cut -c1-200 gpurun_out/resample_sweep_$TAG.jsonl
