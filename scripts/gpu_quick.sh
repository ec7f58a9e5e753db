#!/usr/bin/env bash
# Quick GPU pass: parity tests + one bench line (+ optional extra command).  Logs to gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 -x 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 30 --warmup 5 --cpu-sample 0 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
echo "bench ours exit: $?" >> gpurun_out/bench_ours.err
tail -8 gpurun_out/pytest_gpu.log
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_ours.json'))
    print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'])
    for k in d.get('kernels', []): print(f"{k['ms_per_step']:8.3f} ms {k['share']*100:5.1f}% x{k['launches_per_step']:.0f} {k['kernel']}")
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_ours.err').read()[-1500:])
PY
