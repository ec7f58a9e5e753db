#!/usr/bin/env bash
# N-GPU weak-scaling check (one rank per GPU, torchrun, NCCL only for the timing reduction)
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for impl in reference ours; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 4 --impl $impl --cpu-sample 0 --no-profile > gpurun_out/bench_${impl}_n$N.json 2> gpurun_out/bench_${impl}_n$N.err
  echo "== $impl N=$N exit $?"; tail -n 2 gpurun_out/bench_${impl}_n$N.err | cut -c1-300
  python -c "
import json
d=json.loads(open('gpurun_out/bench_${impl}_n$N.json').read().strip().splitlines()[-1])
print('n_gpus', d['n_gpus'], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']))"
done
