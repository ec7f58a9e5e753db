#!/usr/bin/env bash
set -u
mkdir -p gpurun_out/golden
timeout 900 python -m pytest tests/test_gpu_surface.py tests/test_gpu_ops.py -m gpu -q -p no:cacheprovider --timeout 300 ${PYTEST_ARGS:-} 2>&1 | tail -80 > gpurun_out/pytest_surface.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_surface.log
tail -30 gpurun_out/pytest_surface.log
timeout 900 python scripts/bench_surface.py > gpurun_out/surface_ops.json 2> gpurun_out/surface_ops.err; echo "bench exit $?"; cat gpurun_out/surface_ops.json; tail -5 gpurun_out/surface_ops.err
timeout 900 python bench.py --steps 30 --warmup 5 --cpu-sample 0 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_ours.json')); print('kitti value', d['value'], 'e2e', d['e2e']['value'])
for k in d.get('kernels', [])[:12]: print(f"{k['ms_per_step']:8.3f} ms {k['share']*100:5.1f}% x{k['launches_per_step']:.0f} {k['kernel']}")
PY
