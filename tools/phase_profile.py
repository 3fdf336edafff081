#!/usr/bin/env python
"""Per-world phase cycle counters of the step kernel (clock64 of thread 0): where does a step go, and how
uneven are the worlds?"""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from physicssimulator_b200.simulator import BatchSimulator
ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c3"); ap.add_argument("--worlds", type=int, default=1024)
ap.add_argument("--settle", type=int, default=200); ap.add_argument("--steps", type=int, default=10)
a = ap.parse_args()
built, cfg, joints = bench.build_worlds(a.workload, 0, a.worlds)
sim = BatchSimulator(None, cfg, joints=joints, device=0, _prebuilt=built)
sim.batch.step(a.settle)
c0 = sim.batch.debug_cycles().astype(np.float64)
t = time.perf_counter()
for _ in range(a.steps):
    sim.batch.step(1)
dt = time.perf_counter() - t
c = sim.batch.debug_cycles().astype(np.float64) - c0
ph = c[:, :5] / a.steps
tot = ph.sum(1)
print(f"{a.workload} worlds={a.worlds} wall/step={1e3*dt/a.steps:.3f} ms")
print("mean cycles/step per phase [predict, nearlist, levels+sort, wavefronts, rest]:", np.round(ph.mean(0)))
print("total cycles/step: mean %.0f median %.0f p90 %.0f p99 %.0f max %.0f" % (tot.mean(), np.median(tot), np.percentile(tot, 90), np.percentile(tot, 99), tot.max()))
print("K mean %.1f  levels mean %.1f max %.1f" % ((c[:, 6] / a.steps).mean(), (c[:, 7] / a.steps).mean(), (c[:, 7] / a.steps).max()))
worst = np.argsort(-tot)[:5]
for w in worst:
    print("  world", w, "cycles", np.round(ph[w]), "K", c[w, 6] / a.steps, "levels", c[w, 7] / a.steps)
print("stats", sim.Stats())
