#!/usr/bin/env python
"""Per source line of one kernel in an .ncu-rep (--set full --import-source on): samples and the stall reasons they were
taken in.  usage: ncu_lines.py report.ncu-rep file-substring first-line last-line"""
import collections, csv, io, subprocess, sys
rep, fsub, l0, l1 = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; hdr = None
agg = collections.defaultdict(lambda: collections.Counter()); text = {}
for row in csv.reader(io.StringIO(src)):
    if len(row) == 2 and row[0] == "File Path": cur = row[1].split("/")[-1]; continue
    if row and row[0] == "Line No": hdr = row; continue
    if hdr is None or len(row) < len(hdr) or cur is None or fsub not in cur: continue
    try: line = int(row[0])
    except ValueError: continue
    if not (l0 <= line <= l1): continue
    text[line] = row[1][:70]
    for h, v in zip(hdr, row):
        if h.startswith("stall_") or h in ("# Samples", "Instructions Executed", "Warp Stall Sampling (All Samples)"):
            try: agg[line][h] += float(v.replace(",", ""))
            except ValueError: pass
if hdr: print("columns:", [h for h in hdr if h.startswith("stall") or "ampl" in h][:40])
tot = collections.Counter()
for line in sorted(agg):
    a = agg[line]; tot.update(a)
    top = sorted(((v, k) for k, v in a.items() if k.startswith("stall_") and v > 0), reverse=True)[:4]
    print(f"{line:5d} smp {int(a.get('# Samples', a.get('Warp Stall Sampling (All Samples)', 0))):6d} inst {int(a.get('Instructions Executed', 0)):8d} | " + " ".join(f"{k[6:]}={int(v)}" for v, k in top) + f" | {text[line]}")
print("TOTAL", {k: int(v) for k, v in tot.most_common(14)})
