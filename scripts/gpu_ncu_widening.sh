#!/usr/bin/env bash
# ncu --set full evidence for the kernels of the widening rows (head / NMS, surface features, dense-scene ball query):
# 15 launches in total; the summaries are produced ON the box (the reports are too large to travel back).
set -u
mkdir -p gpurun_out
timeout 300 python scripts/ncu_targets.py > gpurun_out/ncuw_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncuw_plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"nms_mask_kernel|nms_reduce_kernel|detect_sort_kernel|edge_point_kernel|edge_aggr_kernel|bq_grid_query_kernel" -c 15 \
  -o /tmp/prof_widening -f python scripts/ncu_targets.py > gpurun_out/ncuw.log 2>&1
echo "ncu exit: $?"
python scripts/ncu_summary.py /tmp/prof_widening.ncu-rep gpurun_out/ncu_widening_summary.txt > gpurun_out/ncu_summary.log 2>&1; echo "summary exit $?"
ls -la /tmp/prof_widening.ncu-rep; head -c 3000 gpurun_out/ncu_widening_summary.txt
