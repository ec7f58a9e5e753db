#!/usr/bin/env bash
# Round 2 multi-GPU pass: bash scripts/gpu_r2_multi.sh N   (N = 1, 2, 4, 8 GPUs of one box)
#   BASELINE.json configs[3] (Waymo 8 x 65536 per GPU), both arms; strong scaling of one 128-scene KITTI host batch, both arms;
#   at N = 8 additionally the KITTI weak-scaling e2e timeline (H2D / forward / D2H per step) with fp32 and fp16 outputs.
set -u
N=${1:-2}
mkdir -p gpurun_out
run() {  # name, args...
  local name=$1; shift
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --gpus 1 "$@" > gpurun_out/r2_${name}_n${N}.json 2> gpurun_out/r2_${name}_n${N}.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) \
      bench.py --gpus $N "$@" > gpurun_out/r2_${name}_n${N}.json 2> gpurun_out/r2_${name}_n${N}.err
  fi
  echo "== $name N=$N exit $?"; grep -h '^{' gpurun_out/r2_${name}_n${N}.json | cut -c1-420; tail -1 gpurun_out/r2_${name}_n${N}.err | cut -c1-300
}
COMMON="--cpu-sample 0 --no-profile"
run scale_waymo --workload waymo --steps 12 --warmup 3 $COMMON
run scale_waymo_ref --workload waymo --impl reference --steps 4 --warmup 3 $COMMON
run strong_kitti --scaling strong --total-scenes 128 --steps $((4 * N + 4)) --warmup 3 --no-verify $COMMON
run strong_kitti_ref --scaling strong --total-scenes 128 --impl reference --steps 2 --warmup 3 $COMMON
if [ "$N" = "8" ]; then
  run e2e_kitti_timeline --steps 30 --warmup 5 --timeline --no-verify --no-depth1 $COMMON
  run e2e_kitti_out16 --steps 30 --warmup 5 --timeline --out16 --no-verify --no-depth1 $COMMON
  run e2e_kitti_pin --steps 30 --warmup 5 --pin --no-verify --no-depth1 $COMMON
fi
