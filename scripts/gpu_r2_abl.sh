#!/usr/bin/env bash
# ablations of the streaming sa_mma chains on the PROFILING kernel variants (results are wrong by design, only the time matters)
set -u
mkdir -p gpurun_out
: > gpurun_out/r02_sa_mma_ablations.txt
for abl in 0 1 2 3 8 9 10 11; do
  echo "== SPSK_SA_ABL=$abl (1: no weight copies after the first tile, 2: no hidden-epilogue TMEM loads / smem stores, 8: one MMA per weight tile)" >> gpurun_out/r02_sa_mma_ablations.txt
  SPSK_SA_ABL=$abl timeout 120 python scripts/bench_sa_mma.py l5s2 l2s2 l5s1 --prof 2>&1 | grep -v Warning >> gpurun_out/r02_sa_mma_ablations.txt
done
cut -c1-330 gpurun_out/r02_sa_mma_ablations.txt
