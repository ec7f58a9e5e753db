#!/usr/bin/env bash
# Round 2, GPU pass A (1 GPU): full -m gpu suite (incl. the new timed-path / shim / torch-op tests), smoke, bench (ours with the
# verify gate + depth1, reference), Waymo N=1, FPS phase counters, and the L2-side counters of every sa_mma instance.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep -i "numa\|model name\|socket" >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 -s 2>&1 | grep -v "^$" | tail -250 > gpurun_out/r2a_pytest.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/r2a_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1
echo "smoke exit: $?" >> gpurun_out/r2a_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_ours.json 2> gpurun_out/r2a_bench_ours.err
echo "exit $?" >> gpurun_out/r2a_bench_ours.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err
echo "exit $?" >> gpurun_out/r2a_bench_ref.err
timeout 900 python bench.py --workload waymo --steps 10 --warmup 3 --cpu-sample 0 > gpurun_out/r2a_bench_waymo.json 2> gpurun_out/r2a_bench_waymo.err
echo "exit $?" >> gpurun_out/r2a_bench_waymo.err
timeout 300 python scripts/bench_fps.py --prof > gpurun_out/r2a_fps.log 2>&1
CMD="python bench.py --steps 2 --warmup 3 --depth 1 --no-graph --no-profile --no-verify --no-depth1 --cpu-sample 0 --pool 2"
timeout 300 $CMD > gpurun_out/r2a_ncu_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,lts__t_bytes.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__m_xbar2l1tex_read_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_uniform.sum,sm__inst_executed.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:"sa_mma_kernel|pw_mma_kernel" -s 40 -c 40 --csv --log-file gpurun_out/r2a_ncu_l2.csv $CMD > gpurun_out/r2a_ncu_l2.log 2>&1
echo "ncu exit: $?" >> gpurun_out/r2a_ncu_l2.log
tail -15 gpurun_out/r2a_pytest.log; tail -2 gpurun_out/r2a_smoke.log
for f in r2a_bench_ours r2a_bench_ref r2a_bench_waymo; do echo "== $f"; cut -c1-700 gpurun_out/$f.json 2>/dev/null; tail -2 gpurun_out/$f.err 2>/dev/null; done
cat gpurun_out/r2a_fps.log
