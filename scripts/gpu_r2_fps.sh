#!/usr/bin/env bash
# FPS iteration work: bit-exact tests, micro-benchmark with phase counters (TMEM-resident minima vs the register kernel).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -p no:cacheprovider -k "fps" --timeout 600 2>&1 | tail -15 > gpurun_out/r2_fps_pytest.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/r2_fps_pytest.log
timeout 300 python scripts/bench_fps.py --prof > gpurun_out/r2_fps_tm.log 2>&1
SPSK_FPS_NOTMEM=1 timeout 300 python scripts/bench_fps.py > gpurun_out/r2_fps_reg.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-verify --cpu-sample 0 > gpurun_out/r2_fps_bench.json 2> gpurun_out/r2_fps_bench.err
tail -4 gpurun_out/r2_fps_pytest.log; cat gpurun_out/r2_fps_tm.log gpurun_out/r2_fps_reg.log; cut -c1-600 gpurun_out/r2_fps_bench.json; tail -2 gpurun_out/r2_fps_bench.err
