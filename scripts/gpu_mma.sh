#!/usr/bin/env bash
# tensor-core kernel tests first (fast failure), then the whole GPU suite + one bench line
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mma.py -m gpu -q -p no:cacheprovider --timeout 120 -x -s 2>&1 | tail -150 > gpurun_out/pytest_mma.log
rc=${PIPESTATUS[0]}
echo "pytest exit: $rc" >> gpurun_out/pytest_mma.log
grep -E "\[mma\]|\[pw\]|passed|failed|Error|error|exit" gpurun_out/pytest_mma.log | head -80
if [ "$rc" = "0" ]; then bash scripts/gpu_quick.sh; fi
