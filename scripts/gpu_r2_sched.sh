#!/usr/bin/env bash
# static-ring schedule of the streaming sa_mma chains: parity first (short timeouts), then the micro-benchmark and a quick bench
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mma.py -m gpu -q -x -p no:cacheprovider --timeout 100 2>&1 | tail -8 > gpurun_out/r2s_mma.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/r2s_mma.log; tail -9 gpurun_out/r2s_mma.log
timeout 200 python scripts/bench_sa_mma.py --prof 2>&1 | grep -v Warning > gpurun_out/r2s_bench_mma.log; cut -c1-330 gpurun_out/r2s_bench_mma.log
timeout 400 python -m pytest tests/test_gpu_timed_path.py tests/test_gpu_train_fused.py -m gpu -q -x -p no:cacheprovider --timeout 200 2>&1 | tail -5 > gpurun_out/r2s_timed.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/r2s_timed.log; tail -4 gpurun_out/r2s_timed.log
timeout 300 python bench.py --steps 20 --warmup 5 --cpu-sample 0 > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; echo "bench exit $?"
python - <<'PY'
import json
p=json.loads([l for l in open('gpurun_out/r2s_bench.json') if l.startswith('{')][-1])
print('value', round(p['value']), 'e2e', round(p['e2e']['value']), 'depth1', round(p['depth1']['value']), p['depth1']['ms_per_step'], 'verified', p.get('verified'))
for r in p.get('kernels',[]): print('  ', r['kernel'], round(r['ms_per_step']*1e3,1))
PY
