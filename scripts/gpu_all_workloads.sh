#!/usr/bin/env bash
# One bench line per workload (ours), for the numbers quoted in README / DESIGN.
set -u
mkdir -p gpurun_out
for wl in kitti kitti_det spsnet spsnet_sf spsnet_e2e spsnet_full waymo waymo_det; do
  extra="--cpu-sample 0"; [ "$wl" = "kitti" ] && extra=""
  timeout 900 python bench.py --workload $wl --steps 32 --warmup 8 $extra > gpurun_out/bench_wl_$wl.json 2> gpurun_out/bench_wl_$wl.err
  python -c "
import json
d=json.loads(open('gpurun_out/bench_wl_$wl.json').read().strip().splitlines()[-1]); print('$wl', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],3), 'roofline', d.get('roofline',{}).get('kernel'), round(d.get('roofline',{}).get('frac') or 0,3))"
done
timeout 600 python bench.py --steps 32 --warmup 8 --depth 1 --no-profile --cpu-sample 0 > gpurun_out/bench_wl_kitti_d1.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/bench_wl_kitti_d1.json').read().strip().splitlines()[-1]); print('kitti depth1', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"
