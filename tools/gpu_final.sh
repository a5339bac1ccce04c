#!/bin/bash
# Round-2 final pass: full GPU test-suite, smoke, sanitizer, default bench (record), launch list, full ncu capture, resample sweep.
TAG=${1:-r2z}
mkdir -p gpurun_out
rm -f gpurun_out/fragile_report.jsonl
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
cp gpurun_out/fragile_report.jsonl gpurun_out/fragile_report_$TAG.jsonl 2>/dev/null
python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$TAG.log
timeout 700 bash tools/gpu_sanitize.sh $TAG | grep -E "ERROR SUMMARY|hazard|rc="
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c3_$TAG.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/bench_c3_$TAG.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), '%.4g'%d['value'], '%.4g'%d['e2e']['value'], d['e2e']['latency_ms'], '%.4g'%d['e2e_replay']['value'], d['clocks'], d['parity_check'], {k:round(v['update_ms'],3) for k,v in d['clouds'].items()})"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2>&1; tail -1 gpurun_out/bench_ref_$TAG.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_update_mbes_profile -s 2 -c 1 -f -o gpurun_out/prof_mbes_$TAG python bench.py --cloud converged --steps 2 > gpurun_out/ncu_mbes_$TAG.log 2>&1; echo "ncu rc=$?"
timeout 300 python tools/bench_resample.py > gpurun_out/resample_sweep_$TAG.jsonl 2>&1; cut -c1-120 gpurun_out/resample_sweep_$TAG.jsonl
