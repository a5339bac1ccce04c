#!/bin/bash
# Multi-GPU pass: the sharded parity tests on real GPUs (one process per GPU, NCCL + CUDA IPC) for the given world
# sizes, then bench.py at those sizes.  gpurun --gpus N --timeout 1500 -- bash tools/gpu_multi.sh <tag> "<worlds>" [bench args]
TAG=$1; WORLDS=${2:-2}; shift 2
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for W in $WORLDS; do
  [ -n "$NOTEST" ] && continue
  AUVPF_TEST_WORLD=$W timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k "on_gpus" > gpurun_out/pytest_sharded_${TAG}_w$W.log 2>&1
  echo "world $W pytest rc=$?"; tail -4 gpurun_out/pytest_sharded_${TAG}_w$W.log
done
for W in $WORLDS; do
  [ -n "$NOBENCH" ] && continue
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $W --steps 10 --warmup 3 "$@" > gpurun_out/bench_${TAG}_g$W.log 2>&1
  python - <<PY
import json
f='gpurun_out/bench_${TAG}_g$W.log'
l=[x for x in open(f) if x.startswith('{')]
if not l:
    print('gpus $W FAILED'); print(open(f).read()[-2500:])
else:
    d=json.loads(l[-1])
    print('gpus $W', 'ms/step %.3f'%d['ms_per_step'], 'evals/s %.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'], 'parity', d['parity_check'], 'clocks', d['clocks'], {k:round(v,4) for k,v in d['phase_ms_per_cycle'].items()}, 'rs frac %.3f'%d['roofline_resample']['frac'])
    for k in ('c4','c5'):
        if k in d: print('   ', k, 'ms/step %.3f'%d[k]['ms_per_step'], 'evals/s %.3e'%d[k]['value'], 'e2e %.3e'%d[k]['e2e']['value'], 'parity', d[k]['parity_check'], 'rs frac %.3f'%d[k]['roofline_resample']['frac'])
PY
done
