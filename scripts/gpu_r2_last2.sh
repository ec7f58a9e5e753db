#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_train_fused.py -m gpu -q -x -p no:cacheprovider --timeout 100 2>&1 | tail -3
timeout 150 python scripts/bench_train.py --batch 8 --out gpurun_out/r02_train_mlp_final2.json 2>&1 | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r['layer'][:8], 'fused', r['fused'], 'composed fwd+bwd', r['composed']['fwd_bwd_ms'], 'ref', r.get('reference', {}).get('fwd_bwd_ms'))"
