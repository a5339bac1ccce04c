#!/bin/bash
# 8-GPU pass: gpurun --gpus 8 --timeout 1500 -- bash tools/gpu_multi8.sh <tag>
TAG=$1
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/pytest_sharded_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_sharded_$TAG.log
run() { # name nproc args...
  name=$1; n=$2; shift 2
  if [ "$n" = "1" ]; then python bench.py --gpus 1 "$@" > gpurun_out/bench_${TAG}_$name.log 2>&1
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n "$@" > gpurun_out/bench_${TAG}_$name.log 2>&1; fi
  python -c "
import json,sys
f='gpurun_out/bench_${TAG}_$name.log'
l=[x for x in open(f) if x.startswith('{')]
if not l: print('$name FAILED'); print(open(f).read()[-2000:]); sys.exit()
d=json.loads(l[-1]); print('$name', 'ms/step %.3f'%d['ms_per_step'], 'evals/s %.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'], {k:round(v,4) for k,v in d['kernel_ms_per_cycle'].items()}, 'err %.4f'%d['track_error_m'])
"
}
run c3_g1 1 --steps 10 --warmup 3 --no-cpu-baseline
run c3_g2 2 --steps 10 --warmup 3
run c3_g4 4 --steps 10 --warmup 3
run c3_g8 8 --steps 10 --warmup 3
run c4_g8 8 --workload C4 --steps 10 --warmup 3
run c5_g8 8 --workload C5 --steps 10 --warmup 3
