#!/usr/bin/env bash
# One GPU-box pass: parity tests, golden fixtures, smoke, bench (ours + reference), ncu launch list.
# Logs go to gpurun_out/.  Usage: bash scripts/gpu_check.sh [quick]
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 2>&1 | tail -120 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python tests/golden/make_golden.py gpurun_out/golden > gpurun_out/golden.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit: $?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
echo "bench ours exit: $?" >> gpurun_out/bench_ours.err
timeout 600 python bench.py --steps 30 --warmup 5 --depth 1 --no-profile --cpu-sample 0 > gpurun_out/bench_ours_d1.json 2> gpurun_out/bench_ours_d1.err
timeout 600 python bench.py --steps 30 --warmup 5 --depth 1 --no-graph --no-profile --cpu-sample 0 > gpurun_out/bench_ours_eager.json 2> gpurun_out/bench_ours_eager.err
if [ "${1:-}" != "quick" ]; then
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "bench ref exit: $?" >> gpurun_out/bench_ref.err
fi
# ncu launch list (cold-cache, serialised: shares only) -- only after the same command exited 0 without ncu
NCU_CMD="python bench.py --steps 2 --warmup 3 --depth 1 --no-graph --no-profile --cpu-sample 0 --pool 2"
timeout 300 $NCU_CMD > gpurun_out/ncu_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_check.csv $NCU_CMD > gpurun_out/ncu_run.log 2>&1
echo "ncu exit: $?" >> gpurun_out/ncu_run.log
tail -5 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/golden.log; tail -2 gpurun_out/smoke.log
for f in bench_ours bench_ours_d1 bench_ours_eager bench_ref; do echo "== $f"; cat gpurun_out/$f.json 2>/dev/null | cut -c1-1500; tail -2 gpurun_out/$f.err 2>/dev/null; done
tail -3 gpurun_out/ncu_run.log
