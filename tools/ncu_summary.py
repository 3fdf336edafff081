#!/usr/bin/env python
"""Condense an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of counters DESIGN.md argues with."""
import csv, subprocess, sys, io
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct",
        "l1tex__t_sector_pipe_lsu_mem_local_op_st_hit_rate.pct", "smsp__sass_average_data_bytes_per_sector_mem_local_op_ld.ratio",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "smsp__average_warp_latency_per_inst_issued.ratio"]
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print(f"# kernel: {name}")
    for i, h in enumerate(hdr):
        if h in KEYS:
            print(f"{h},{units[i]},{r[i]}")
        elif h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
            try:
                if float(r[i]) >= 0.05:
                    print(f"{h},{units[i]},{r[i]}")
            except ValueError:
                pass
