#!/bin/bash
# Multi-GPU pass: gpurun --gpus N --timeout 1200 -- bash tools/gpu_multi.sh <tag> <N>
TAG=$1; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/pytest_sharded_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_sharded_$TAG.log
for n in 1 $N; do
  if [ "$n" = "1" ]; then python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_g1.log 2>&1
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_g$n.log 2>&1; fi
  python -c "
import json,sys
f='gpurun_out/bench_${TAG}_g$n.log'
l=[x for x in open(f) if x.startswith('{')]
if not l: print(open(f).read()[-2500:]); sys.exit()
d=json.loads(l[-1]); print('gpus $n', 'ms/step %.3f'%d['ms_per_step'], 'evals/s %.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'], {k:round(v,4) for k,v in d['kernel_ms_per_cycle'].items()}, 'err %.4f'%d['track_error_m'])
"
done
