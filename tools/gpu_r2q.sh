#!/bin/bash
# Round-2 pass Q: default bench (the record for profiles/), then the ncu launch list of the same command and one full capture of the update kernel.
TAG=${1:-r2q}
mkdir -p gpurun_out
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c3_$TAG.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/bench_c3_$TAG.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['e2e']['latency_ms'], round(d['e2e_replay']['ms_per_step'],3), d['clocks'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_update_mbes_profile -s 2 -c 1 -f -o gpurun_out/prof_mbes_$TAG python bench.py --cloud converged --steps 2 > gpurun_out/ncu_mbes_$TAG.log 2>&1; echo "ncu rc=$?"
