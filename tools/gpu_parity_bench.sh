#!/bin/bash
# Round-2 pass L: beam loop of the profile march rewritten (float32 FMA gap, address counter, no-graze instantiation):
# parity tests, fragile report, bench (default, with cloud shapes).
TAG=${1:-r2l}
mkdir -p gpurun_out
rm -f gpurun_out/fragile_report.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_$TAG.log
cp gpurun_out/fragile_report.jsonl gpurun_out/fragile_report_$TAG.jsonl 2>/dev/null
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c3_$TAG.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/bench_c3_$TAG.log | python -c "
import json,sys; d=json.loads(sys.stdin.read())
print({k:d.get(k) for k in ('track_error_m','track_error_m_final','ms_per_step','value')}, d.get('e2e',{}).get('value'))
print('phases', d.get('phases_ms') or d.get('phase_ms'))
print('clouds', d.get('clouds'))
print('parity', d.get('parity_check'))"
