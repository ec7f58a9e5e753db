#!/usr/bin/env bash
# source-level ncu of one sa_mma launch per chain class: hot lines (where the warps stall)
set -u
mkdir -p gpurun_out
for name in l0s2 l5s2; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:sa_mma_kernel -s 6 -c 1 -o /tmp/prof_$name -f python scripts/bench_sa_mma.py $name > gpurun_out/r2_hot_$name.log 2>&1
  echo "ncu $name exit $?"
  python scripts/ncu_hot_lines.py /tmp/prof_$name.ncu-rep 0 28 > gpurun_out/r2_hot_lines_$name.txt 2>&1
  python scripts/ncu_summary.py /tmp/prof_$name.ncu-rep gpurun_out/r2_hot_summary_$name.txt > /dev/null 2>&1
  head -c 5000 gpurun_out/r2_hot_lines_$name.txt
done
