#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_train_fused.py -m gpu -q -x -p no:cacheprovider --timeout 100 2>&1 | tail -2
for i in 1 2; do timeout 200 python bench.py --workload waymo --steps 20 --warmup 4 --cpu-sample 0 --no-verify --no-depth1 > gpurun_out/r2l_waymo_$i.json 2> gpurun_out/r2l_waymo_$i.err; done
SPSK_SA_NO_DOUBLE=1 timeout 200 python bench.py --workload waymo --steps 20 --warmup 4 --cpu-sample 0 --no-verify --no-depth1 > gpurun_out/r2l_waymo_nodouble.json 2> gpurun_out/r2l_waymo_nodouble.err
python - <<'PY'
import json
for f in ("r2l_waymo_1","r2l_waymo_2","r2l_waymo_nodouble"):
    p=json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1]); print(f, round(p["value"],1), round(p["e2e"]["value"],1), p["ms_per_step"])
PY
