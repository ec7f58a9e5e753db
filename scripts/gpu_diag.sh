#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python scripts/diag_precision.py > gpurun_out/diag_precision.log 2>&1
tail -12 gpurun_out/diag_precision.log
CMD="python bench.py --steps 2 --warmup 3 --depth 1 --no-graph --no-profile --cpu-sample 0 --pool 2"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sa_mma_kernel -s 8 -c 8 -o gpurun_out/prof_sa_mma_v5 -f $CMD > gpurun_out/ncu_mma.log 2>&1
echo "ncu mma exit: $?"
