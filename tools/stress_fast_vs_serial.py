#!/usr/bin/env python
"""GPU stress: the wavefront path against the exact serial walk (both on the device), many worlds, many steps.
Reports the first (step, world) whose dynamic state differs."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from physicssimulator_b200 import scenes  # noqa: E402
from physicssimulator_b200.simulator import BatchSimulator  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c3")
ap.add_argument("--worlds", type=int, default=32)
ap.add_argument("--first", type=int, default=0)
ap.add_argument("--steps", type=int, default=260)
ap.add_argument("--chunk", type=int, default=10)
ap.add_argument("--path", default="", help="force PS_B200_PATH for the fast batch (wide|fused)")
args = ap.parse_args()
fn = {"c2": scenes.c2_spheres64, "c3": scenes.c3_mixed128}[args.workload]
worlds = [fn(w)[0] for w in range(args.first, args.first + args.worlds)]
cfg = fn(0)[1]
if args.path:
    os.environ['PS_B200_PATH'] = args.path
a = BatchSimulator(worlds, cfg, device=0)
os.environ.pop('PS_B200_PATH', None)
b = BatchSimulator(worlds, cfg, device=0, exact_all_pairs=True)
off = a._offsets
done = 0
bad = None
while done < args.steps and bad is None:
    a.Step(args.chunk); b.Step(args.chunk)
    done += args.chunk
    sa, sb = a.State(), b.State()
    for w in range(args.worlds):
        for k in ("pos", "vel", "rot"):
            x, y = sa[k][off[w]:off[w + 1]], sb[k][off[w]:off[w + 1]]
            if not np.array_equal(x.view(np.uint64), y.view(np.uint64)):
                rows = np.unique(np.argwhere(x.view(np.uint64) != y.view(np.uint64))[:, 0])
                bad = (done, w + args.first, k, rows[:10])
                break
        if bad:
            break
print("stats fast  ", a.Stats())
print("stats serial", b.Stats())
print("FIRST MISMATCH" if bad else "all equal", bad, "after", done, "steps")
