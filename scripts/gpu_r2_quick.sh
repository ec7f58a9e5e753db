#!/usr/bin/env bash
# quick loop: mma + timed-path tests, sa_mma micro-bench, bench (depth 8 + depth 1)
set -u
mkdir -p gpurun_out
timeout 120 python scripts/bench_sa_mma.py > gpurun_out/r2q_mma.log 2>&1; echo "exit $?" >> gpurun_out/r2q_mma.log
cat gpurun_out/r2q_mma.log
timeout 500 python -m pytest tests/test_gpu_mma.py tests/test_gpu_timed_path.py tests/test_gpu_modules.py -m gpu -q -p no:cacheprovider --timeout 200 -x 2>&1 | tail -5 > gpurun_out/r2q_pytest.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/r2q_pytest.log; tail -3 gpurun_out/r2q_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 --cpu-sample 0 --no-verify > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo "bench exit $?"
python - <<'PY'
import json
p=json.loads([l for l in open('gpurun_out/r2q_bench.json') if l.startswith('{')][-1])
print('value', round(p['value']), 'e2e', round(p['e2e']['value']), 'depth1', round(p['depth1']['value']), p['depth1']['ms_per_step'], 'launches/step', p['config']['launches_per_step'])
for r in p.get('kernels',[]): print('  ', r['kernel'], round(r['ms_per_step']*1e3,1))
PY
