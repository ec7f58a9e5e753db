#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
: > gpurun_out/r02_sa_mma_ablations_v2.txt
for abl in 0 1 2 8 9; do
  echo "== SPSK_SA_ABL=$abl (1: no weight copies after the first tile, 2: no hidden-epilogue TMEM loads / smem stores, 8: one MMA per schedule entry)" >> gpurun_out/r02_sa_mma_ablations_v2.txt
  SPSK_SA_ABL=$abl timeout 120 python scripts/bench_sa_mma.py l5s2 l2s2 l1s2 --prof 2>&1 | grep -v Warning >> gpurun_out/r02_sa_mma_ablations_v2.txt
done
grep "^==\|profiling-kernel\|^l" gpurun_out/r02_sa_mma_ablations_v2.txt | cut -c1-120
