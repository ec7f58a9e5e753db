#!/usr/bin/env bash
# Source-level ncu profile of the 16384 -> 4096 FPS kernel; hot lines extracted on the box.
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fps_pruned_kernel -s 2 -c 1 -o /tmp/prof_fps -f python scripts/bench_fps.py --bench-data > gpurun_out/ncu_fps.log 2>&1
echo "ncu exit: $?"
python scripts/ncu_hot_lines.py /tmp/prof_fps.ncu-rep 0 60 > gpurun_out/fps_hot_lines.txt 2>&1; echo "hot lines exit $?"
python scripts/ncu_summary.py /tmp/prof_fps.ncu-rep gpurun_out/fps_summary.txt > /dev/null 2>&1
head -c 6000 gpurun_out/fps_hot_lines.txt
