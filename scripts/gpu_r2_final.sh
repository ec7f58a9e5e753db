#!/usr/bin/env bash
# Round 2 final single-GPU pass: full -m gpu suite, smoke, bench (kitti / spsnet, both arms), ncu launch list + per-instance DRAM
# traffic + --set full summaries of the top kernels.  Everything lands in gpurun_out/r2f_*.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/r2f_gpu.txt 2>&1; nproc >> gpurun_out/r2f_gpu.txt
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 -s 2>&1 | grep -v "^$" | tail -200 > gpurun_out/r2f_pytest.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/r2f_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; echo "smoke exit: $?" >> gpurun_out/r2f_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench_kitti.json 2> gpurun_out/r2f_bench_kitti.err; echo "exit $?" >> gpurun_out/r2f_bench_kitti.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2f_bench_kitti_reference.json 2> gpurun_out/r2f_bench_kitti_reference.err
timeout 600 python bench.py --workload spsnet --steps 20 --warmup 5 --cpu-sample 0 > gpurun_out/r2f_bench_spsnet.json 2> gpurun_out/r2f_bench_spsnet.err
timeout 600 python bench.py --workload spsnet --impl reference --steps 10 --warmup 3 > gpurun_out/r2f_bench_spsnet_reference.json 2> gpurun_out/r2f_bench_spsnet_reference.err
timeout 300 python scripts/bench_fps.py --prof > gpurun_out/r2f_fps.log 2>&1
timeout 200 python scripts/bench_sa_mma.py --prof > gpurun_out/r2f_sa_mma.log 2>&1
timeout 200 python scripts/diag_precision.py > gpurun_out/r2f_diag_precision.log 2>&1
# ncu: the plain run first (B200_PROFILING.md), then the launch list with DRAM bytes of the same command
CMD="python bench.py --steps 3 --warmup 3 --depth 1 --no-graph --no-verify --no-depth1 --cpu-sample 0 --pool 2"
SPSK_DUMP_CALLS=gpurun_out/r2f_calls.json timeout 300 $CMD > gpurun_out/r2f_ncu_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2f_ncu_launches.csv $CMD > gpurun_out/r2f_ncu_run.log 2>&1
echo "ncu launch list exit: $?"
python scripts/ncu_traffic.py gpurun_out/r2f_calls.json gpurun_out/r2f_ncu_launches.csv gpurun_out/r2f_traffic.json gpurun_out/r2f_traffic.txt > /dev/null 2> gpurun_out/r2f_traffic.err; echo "traffic exit $?"
# --set full of the last eager step's sa_mma / fps / misc kernels (summaries only; reports stay on the box)
timeout 900 ncu --set full --clock-control none -k regex:sa_mma_kernel -s 40 -c 8 -o /tmp/r2f_sa -f $CMD > gpurun_out/r2f_ncu_sa.log 2>&1; python scripts/ncu_summary.py /tmp/r2f_sa.ncu-rep gpurun_out/r2f_ncu_sa_mma.txt > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none -k regex:fps_pruned_kernel -s 10 -c 2 -o /tmp/r2f_fps -f $CMD > gpurun_out/r2f_ncu_fps.log 2>&1; python scripts/ncu_summary.py /tmp/r2f_fps.ncu-rep gpurun_out/r2f_ncu_fps.txt > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"bq_grid_query_kernel|pw_mma_kernel|score_topk_kernel" -s 65 -c 13 -o /tmp/r2f_misc -f $CMD > gpurun_out/r2f_ncu_miscrun.log 2>&1; python scripts/ncu_summary.py /tmp/r2f_misc.ncu-rep gpurun_out/r2f_ncu_misc.txt > /dev/null 2>&1
tail -3 gpurun_out/r2f_pytest.log; tail -1 gpurun_out/r2f_smoke.log; cat gpurun_out/r2f_traffic.err | tail -2
for f in r2f_bench_kitti r2f_bench_kitti_reference r2f_bench_spsnet r2f_bench_spsnet_reference; do echo "== $f"; grep -h '^{' gpurun_out/$f.json | cut -c1-330; done
