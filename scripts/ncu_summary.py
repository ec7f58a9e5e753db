#!/usr/bin/env python
"""Compact per-launch summary of an .ncu-rep (ncu --set full): python scripts/ncu_summary.py rep [out.md]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
def col(name):
    return hdr.index(name) if name in hdr else None
want = [("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("launch__shared_mem_per_block_dynamic", "smem_dyn"), ("launch__occupancy_limit_shared_mem", "occ_smem"),
        ("gpu__time_duration.sum", "time"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "hmma_inst%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1%"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("lts__t_bytes.sum", "l2_bytes"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "st_long_sb"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
        ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall_membar"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
        ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "stall_sleep"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall_mio"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall_branch"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall_noinst"),
        ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall_dispatch"),
        ("smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio", "stall_tex"),
        ]
out = []
for r in data:
    line = []
    for name, short in want:
        c = col(name)
        if c is None or c >= len(r):
            continue
        v = r[c]
        u = units[c] if c < len(units) else ""
        if short == "kernel":
            v = v.split("(")[0][-40:]
        else:
            try:
                f = float(v.replace(",", ""))
                v = f"{f:.3g}"
            except ValueError:
                pass
        line.append(f"{short}={v}{u if short in ('time','dram_rd','dram_wr','l2_bytes','smem_dyn') else ''}")
    out.append(" ".join(line))
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")
