#!/usr/bin/env bash
# Round 2, last single-GPU pass on the final kernels: the round-1/2 pass (scripts/gpu_r2_final.sh: full -m gpu suite, smoke, bench
# kitti / spsnet both arms, fps / sa_mma counters, precision diag, ncu launch list + traffic + --set full summaries) plus the
# Waymo-shaped line (both arms), the training micro-benchmark and an ncu launch list of one training forward.
set -u
bash scripts/gpu_r2_final.sh
timeout 400 python bench.py --workload waymo --steps 10 --warmup 3 --cpu-sample 0 > gpurun_out/r2f_bench_waymo.json 2> gpurun_out/r2f_bench_waymo.err
timeout 400 python bench.py --workload waymo --impl reference --steps 5 --warmup 3 > gpurun_out/r2f_bench_waymo_reference.json 2> gpurun_out/r2f_bench_waymo_reference.err
for f in r2f_bench_waymo r2f_bench_waymo_reference; do echo "== $f"; grep -h '^{' gpurun_out/$f.json | cut -c1-330; done
timeout 400 python scripts/bench_train.py --batch 8 --out gpurun_out/r2f_train_mlp.json > gpurun_out/r2f_bench_train.log 2>&1; echo "bench_train exit $?"; grep "^{" gpurun_out/r2f_bench_train.log | cut -c1-400
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2f_ncu_train_launches.csv python scripts/train_forward_once.py > gpurun_out/r2f_ncu_train.log 2>&1; echo "ncu train exit $?"
