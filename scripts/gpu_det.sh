#!/usr/bin/env bash
# GPU pass for the head + NMS widening: parity tests, op micro-benchmarks, detector bench (ours + reference arm).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_det.py -m gpu -q -p no:cacheprovider --timeout 300 ${PYTEST_ARGS:-} 2>&1 | tail -120 > gpurun_out/pytest_det.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_det.log
tail -15 gpurun_out/pytest_det.log
timeout 600 python scripts/bench_det.py > gpurun_out/det_ops.json 2> gpurun_out/det_ops.err; echo "det ops exit $?"; cat gpurun_out/det_ops.json; tail -5 gpurun_out/det_ops.err
timeout 900 python bench.py --workload kitti_det --steps 30 --warmup 5 --cpu-sample 4 > gpurun_out/bench_det.json 2> gpurun_out/bench_det.err; echo "bench det exit $?"; tail -5 gpurun_out/bench_det.err
timeout 900 python bench.py --workload kitti_det --impl reference --steps 10 --warmup 3 > gpurun_out/bench_det_ref.json 2> gpurun_out/bench_det_ref.err; echo "bench det ref exit $?"; tail -5 gpurun_out/bench_det_ref.err
python - <<'PY'
import json
for f in ('gpurun_out/bench_det.json','gpurun_out/bench_det_ref.json'):
    try:
        d=json.load(open(f)); print(f, 'value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches/step', d['config'].get('launches_per_step'))
        for k in d.get('kernels', [])[:14]: print(f"{k['ms_per_step']:8.3f} ms {k['share']*100:5.1f}% x{k['launches_per_step']:.0f} {k['kernel']}")
    except Exception as e: print(f, 'parse failed', e)
PY
if [ -n "${EXTRA:-}" ]; then bash -c "$EXTRA"; fi
