#!/usr/bin/env bash
# Round-end evidence in one gpurun call: ncu captures, full GPU check, reference arm.
set -u
mkdir -p gpurun_out
bash scripts/gpu_ncu.sh > gpurun_out/ncu_sh.log 2>&1; tail -6 gpurun_out/ncu_sh.log
bash scripts/gpu_check.sh quick > gpurun_out/check.log 2>&1; grep -E "passed|failed|smoke|wrote" gpurun_out/check.log | head
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2>/dev/null
python - <<'PY'
import json
for f in ("bench_ours", "bench_ours_d1", "bench_ours_eager", "bench_ref"):
    try:
        d = json.loads(open("gpurun_out/" + f + ".json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"], 3), round(d["e2e"]["value"]))
    except Exception as e:
        print(f, "failed", e)
PY
