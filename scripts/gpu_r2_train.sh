#!/usr/bin/env bash
# training-mode fused path: its GPU tests first (short timeouts: a hung kernel must not eat the budget), then the
# tensor-core regression tests (the kernel gained a statistics epilogue) and a quick bench for the inference numbers
set -u
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_train_fused.py -m gpu -q -s -p no:cacheprovider --timeout 120 2>&1 | tail -60 > gpurun_out/r2t_train.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/r2t_train.log; tail -40 gpurun_out/r2t_train.log
timeout 400 python -m pytest tests/test_gpu_mma.py tests/test_gpu_timed_path.py -m gpu -q -p no:cacheprovider --timeout 200 -x 2>&1 | tail -5 > gpurun_out/r2t_mma.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/r2t_mma.log; tail -3 gpurun_out/r2t_mma.log
timeout 300 python bench.py --steps 20 --warmup 5 --cpu-sample 0 --no-verify > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo "bench exit $?"
python - <<'PY'
import json
p=json.loads([l for l in open('gpurun_out/r2t_bench.json') if l.startswith('{')][-1])
print('value', round(p['value']), 'e2e', round(p['e2e']['value']), 'depth1', round(p['depth1']['value']), p['depth1']['ms_per_step'])
PY
