#!/usr/bin/env bash
# GPU pass for the head + NMS widening: its parity tests (vs oracle and vs the rebuilt reference kernels).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_det.py -m gpu -q -p no:cacheprovider --timeout 300 ${PYTEST_ARGS:-} 2>&1 | tail -120 > gpurun_out/pytest_det.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_det.log
tail -60 gpurun_out/pytest_det.log
if [ -n "${EXTRA:-}" ]; then bash -c "$EXTRA"; fi
