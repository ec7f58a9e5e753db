#!/usr/bin/env bash
set -u
mkdir -p gpurun_out/golden
timeout 900 python -m pytest tests/test_gpu_surface.py tests/test_gpu_ops.py -m gpu -q -p no:cacheprovider --timeout 300 -k "surface or ball_query or extractor or pagnet or cin or autograd" ${PYTEST_ARGS:-} 2>&1 | tail -80 > gpurun_out/pytest_surface.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_surface.log
tail -8 gpurun_out/pytest_surface.log
timeout 900 python scripts/bench_surface.py > gpurun_out/surface_ops.json 2> gpurun_out/surface_ops.err; echo "bench exit $?"; cat gpurun_out/surface_ops.json; tail -5 gpurun_out/surface_ops.err
if [ -n "${EXTRA:-}" ]; then bash -c "$EXTRA"; fi
