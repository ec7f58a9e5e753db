#!/usr/bin/env bash
# ncu evidence for one eager step (one gpurun call; the plain run first, as B200_PROFILING.md requires):
#   1. launch list  (gpu__time_duration.sum, cold-cache + serialised: shares only)
#   2. --set full captures of the step's top kernels (with source), for the profiles/ summaries and `traffic`
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --depth 1 --no-graph --no-profile --cpu-sample 0 --pool 2"
timeout 300 $CMD > gpurun_out/ncu_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_run.log 2>&1
echo "launch list exit: $?"
# last step only: skip the warm-up launches of each kernel class
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sa_mma_kernel -s 32 -c 8 -o gpurun_out/prof_sa_mma -f $CMD > gpurun_out/ncu_mma.log 2>&1
echo "ncu sa_mma exit: $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fps_pruned_kernel -s 8 -c 2 -o gpurun_out/prof_fps -f $CMD > gpurun_out/ncu_fps.log 2>&1
echo "ncu fps exit: $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bq_grid_query_kernel|pw_mma_kernel|score_topk_kernel" -s 52 -c 13 -o gpurun_out/prof_misc -f $CMD > gpurun_out/ncu_misc.log 2>&1
echo "ncu misc exit: $?"
ls -la gpurun_out/*.ncu-rep
