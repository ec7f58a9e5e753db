#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_train_fused.py -m gpu -q -s -p no:cacheprovider --timeout 120 > gpurun_out/r2t_train.log 2>&1
echo "pytest exit: $?" >> gpurun_out/r2t_train.log; grep "\[stats\]\|\[train vs\|passed\|failed\|exit" gpurun_out/r2t_train.log
timeout 400 python scripts/bench_train.py --batch 8 --out gpurun_out/r02_train_mlp.json > gpurun_out/r2t_bench_train.log 2>&1; echo "bench_train exit $?"; tail -8 gpurun_out/r2t_bench_train.log
