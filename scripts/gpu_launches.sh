#!/usr/bin/env bash
# ncu launch list of one eager step (cold-cache, serialised: shares only)
set -u
mkdir -p gpurun_out
NCU_CMD="python bench.py --steps 2 --warmup 3 --depth 1 --no-graph --no-profile --cpu-sample 0 --pool 2"
timeout 300 $NCU_CMD > gpurun_out/ncu_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv $NCU_CMD > gpurun_out/ncu_run.log 2>&1
echo "ncu exit: $?"
