#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for st in 2 3 4 6 8; do echo "== max stages $st"; SPSK_SA_MAX_STAGES=$st timeout 100 python scripts/bench_sa_mma.py l2s2 l5s1 l2s1 l1s2; done > gpurun_out/r2_stages.log 2>&1
echo "== l5s2 no lring" >> gpurun_out/r2_stages.log; SPSK_SA_NO_LRING=1 timeout 100 python scripts/bench_sa_mma.py l5s2 >> gpurun_out/r2_stages.log 2>&1
echo "== prof" >> gpurun_out/r2_stages.log; timeout 100 python scripts/bench_sa_mma.py l5s2 l2s2 l0s2 --prof >> gpurun_out/r2_stages.log 2>&1
cat gpurun_out/r2_stages.log
