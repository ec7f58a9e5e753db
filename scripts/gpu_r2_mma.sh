#!/usr/bin/env bash
# sa_mma work: parity (mma tests + many-tile tests vs the oracle), micro-benchmark of every KITTI chain (scout warp on / off).
set -u
mkdir -p gpurun_out
timeout 120 python scripts/bench_sa_mma.py l5s2 l2s2 > gpurun_out/r2_mma_quick.log 2>&1; echo "quick exit $?" >> gpurun_out/r2_mma_quick.log
cat gpurun_out/r2_mma_quick.log
if grep -q "quick exit 0" gpurun_out/r2_mma_quick.log; then
timeout 400 python -m pytest tests/test_gpu_mma.py tests/test_gpu_timed_path.py -m gpu -q -p no:cacheprovider -k "mma or many_tiles" --timeout 100 -x 2>&1 | tail -8 > gpurun_out/r2_mma_pytest.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/r2_mma_pytest.log
timeout 200 python scripts/bench_sa_mma.py > gpurun_out/r2_mma_bench.log 2>&1
SPSK_SA_NO_SCOUT=1 timeout 200 python scripts/bench_sa_mma.py > gpurun_out/r2_mma_bench_noscout.log 2>&1
tail -4 gpurun_out/r2_mma_pytest.log; cat gpurun_out/r2_mma_bench.log; echo "--- no scout"; cat gpurun_out/r2_mma_bench_noscout.log
fi
