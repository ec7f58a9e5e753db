#!/usr/bin/env bash
# ncu captures of the top kernels (one gpurun call; plain run first, as B200_PROFILING.md requires)
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --depth 1 --no-graph --no-profile --cpu-sample 0 --pool 2"
timeout 300 $CMD > gpurun_out/ncu_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sa_mma_kernel -s 8 -c 8 -o gpurun_out/prof_sa_mma -f $CMD > gpurun_out/ncu_mma.log 2>&1
echo "ncu mma exit: $?" >> gpurun_out/ncu_mma.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fps_pruned_kernel -s 2 -c 2 -o gpurun_out/prof_fps -f $CMD > gpurun_out/ncu_fps.log 2>&1
echo "ncu fps exit: $?" >> gpurun_out/ncu_fps.log
tail -3 gpurun_out/ncu_mma.log gpurun_out/ncu_fps.log; ls -la gpurun_out/*.ncu-rep
