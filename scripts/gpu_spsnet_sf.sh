#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python bench.py --workload spsnet_sf --steps 20 --warmup 4 --cpu-sample 0 > gpurun_out/bench_spsnet_sf.json 2> gpurun_out/bench_spsnet_sf.err; echo "ours exit $?"; tail -3 gpurun_out/bench_spsnet_sf.err
timeout 900 python bench.py --workload spsnet_sf --impl reference --steps 6 --warmup 3 > gpurun_out/bench_spsnet_sf_ref.json 2> gpurun_out/bench_spsnet_sf_ref.err; echo "ref exit $?"; tail -3 gpurun_out/bench_spsnet_sf_ref.err
timeout 900 python bench.py --workload spsnet --impl reference --steps 6 --warmup 3 > gpurun_out/bench_spsnet_ref.json 2> gpurun_out/bench_spsnet_ref.err; echo "ref2 exit $?"; tail -3 gpurun_out/bench_spsnet_ref.err
timeout 900 python bench.py --workload spsnet --steps 20 --warmup 4 --cpu-sample 0 > gpurun_out/bench_spsnet.json 2> gpurun_out/bench_spsnet.err; echo "ours2 exit $?"
python - <<'PY'
import json
for f in ('bench_spsnet_sf','bench_spsnet_sf_ref','bench_spsnet','bench_spsnet_ref'):
    try:
        d=json.load(open(f'gpurun_out/{f}.json')); print(f, 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))
        for k in d.get('kernels', [])[:10]: print(f"{k['ms_per_step']:8.3f} ms {k['share']*100:5.1f}% x{k['launches_per_step']:.0f} {k['kernel']}")
    except Exception as e: print(f, 'parse failed', e)
PY
