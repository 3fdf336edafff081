#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (csv with gpu__time_duration.sum and dram bytes): count, time, DRAM read/written."""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
per = collections.defaultdict(lambda: collections.defaultdict(float)); cnt = collections.Counter(); order = []
for row in csv.DictReader(lines):
    k = row["Kernel Name"].split("(")[0]
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]; m = row["Metric Name"]
    if m == "gpu__time_duration.sum":
        v = v / 1e3 if u in ("ns", "nsecond") else v
        cnt[k] += 1; order.append((k, v))
    if "bytes" in m:
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    per[k][m] += v
tt = sum(d["gpu__time_duration.sum"] for d in per.values()); tb = sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in per.values())
print(f"total {tt:.1f} us over {sum(cnt.values())} launches, DRAM {tb / 1e9:.3f} GB")
for k, d in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
    print(f"{k:42s} n={cnt[k]:3d} t={d['gpu__time_duration.sum']:9.1f} us  rd={d['dram__bytes_read.sum'] / 1e6:9.1f} MB  wr={d['dram__bytes_write.sum'] / 1e6:9.1f} MB")
if len(sys.argv) > 2:
    import itertools
    for name in sys.argv[2:]:
        print(name, " ".join(f"{v:.0f}" for k, v in order if name in k))
