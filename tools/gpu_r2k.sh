#!/bin/bash
# Round-2 pass K: wave-shaped select tile -- resample/sharded tests, sweep with the shaped tile and with 1024 pinned, bench (default).
TAG=${1:-r2k}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python tools/bench_resample.py > gpurun_out/resample_sweep_$TAG.jsonl 2>&1; cut -c1-220 gpurun_out/resample_sweep_$TAG.jsonl
AUVPF_SEL_TILE=1024 timeout 300 python tools/bench_resample.py > gpurun_out/resample_sweep_${TAG}_t1024.jsonl 2>&1; cut -c1-220 gpurun_out/resample_sweep_${TAG}_t1024.jsonl
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c3_$TAG.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_c3_$TAG.log | cut -c1-400
tail -1 gpurun_out/bench_c3_$TAG.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d.get(k) for k in ('track_error_m','track_error_m_final','ms_per_step','value')}, d.get('e2e'))"
