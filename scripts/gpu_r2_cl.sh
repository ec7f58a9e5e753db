#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for cl in 1 0; do
  echo "== SPSK_TRAIN_CHANNELS_LAST=$cl"
  SPSK_TRAIN_CHANNELS_LAST=$cl timeout 200 python scripts/bench_train.py --batch 8 --out gpurun_out/r02_train_mlp_cl$cl.json 2>&1 | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r['layer'][:8], 'fused fwd+bwd', r['fused']['fwd_bwd_ms'], 'composed', r['composed']['fwd_bwd_ms'], 'ref', r.get('reference', {}).get('fwd_bwd_ms'), 'peak', r['fused']['peak_mb'])"
done
SPSK_TRAIN_CHANNELS_LAST=1 timeout 200 python -m pytest tests/test_gpu_train_fused.py -m gpu -q -x -p no:cacheprovider --timeout 100 2>&1 | tail -2
