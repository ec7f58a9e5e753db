#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mma.py -m gpu -q -p no:cacheprovider --timeout 120 -x -s 2>&1 | tail -70 > gpurun_out/pytest_mma.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_mma.log
cat gpurun_out/pytest_mma.log | grep -E "\[mma\]|passed|failed|Error|error|exit" | head -60
