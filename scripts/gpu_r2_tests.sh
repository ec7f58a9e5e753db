#!/usr/bin/env bash
# full -m gpu suite + smoke (bounded) -> gpurun_out/r2_pytest.log
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 -s ${SPSK_PYTEST_ARGS:-} 2>&1 | grep -v "^$" | tail -150 > gpurun_out/r2_pytest.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/r2_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke exit: $?" >> gpurun_out/r2_smoke.log
grep -n "passed\|failed\|FAILED\|Error\|exit" gpurun_out/r2_pytest.log | tail -20; tail -2 gpurun_out/r2_smoke.log
