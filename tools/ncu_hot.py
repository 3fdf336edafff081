#!/usr/bin/env python
"""Hot source lines of one kernel in an .ncu-rep (ncu --set full --import-source on): samples, instructions, threads per
instruction, aggregated per source line.  usage: ncu_hot.py report.ncu-rep [top]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "sm__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_active.avg", "sm__cycles_elapsed.max", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__average_warp_latency_per_inst_issued.ratio",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
for i, h in enumerate(hdr):
    if h in want or ("stalled" in h and "per_issue_active" in h):
        try:
            if "stalled" in h and float(vals[i].replace(",", "")) < 0.1: continue
        except ValueError: pass
        print(f"{h:85s} {units[i]:12s} {vals[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
agg = collections.defaultdict(lambda: [0, 0, 0, ""]); cur = None
for row in csv.reader(io.StringIO(src)):
    if len(row) == 2 and row[0] == "File Path": cur = row[1].split("/")[-1]; continue
    if len(row) < 12 or row[0] in ("Line No", ""): continue
    try: line, ins, tins, smp = int(row[0]), int(row[7]), int(row[8]), int(row[4])
    except ValueError: continue
    a = agg[(cur, line)]; a[0] += ins; a[1] += tins; a[2] += smp; a[3] = row[1][:100]
tots = sum(a[2] for a in agg.values()) or 1
print("total instructions", sum(a[0] for a in agg.values()), "samples", tots)
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    print(f"{f}:{l:4d} smp {a[2]:6d} ({100 * a[2] / tots:4.1f}%) inst {a[0]:8d} thr/inst {a[1] / max(a[0], 1):5.1f}  {a[3]}")
