#!/usr/bin/env python
"""ms per step of a workload in windows of steps from creation (device-timed), with the fallback counters per window."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from physicssimulator_b200.simulator import BatchSimulator
ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c3"); ap.add_argument("--worlds", type=int, default=16384)
ap.add_argument("--window", type=int, default=25); ap.add_argument("--steps", type=int, default=400)
a = ap.parse_args()
built, cfg, joints = bench.build_worlds(a.workload, 0, a.worlds, min(16, os.cpu_count() or 1))
sim = BatchSimulator(None, cfg, joints=joints, device=0, _prebuilt=built)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
prev = sim.Stats()
for s0 in range(0, a.steps, a.window):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.window):
        sim.batch.step_async(1, st.cuda_stream)
    e1.record(); torch.cuda.synchronize()
    cur = sim.Stats()
    d = {k: cur[k] - prev[k] for k in ("fallback_steps", "fallback_margin_steps", "fallback_capacity_steps", "fallback_static_steps", "margin_verified_steps", "pairs_tested", "pairs_colliding")}
    prev = cur
    print(f"steps {s0:4d}-{s0 + a.window:4d}: {e0.elapsed_time(e1) / a.window:8.3f} ms/step  fallbacks {d['fallback_steps']:6d} (margin {d['fallback_margin_steps']}, capacity {d['fallback_capacity_steps']}, static {d['fallback_static_steps']}) verified {d['margin_verified_steps']:6d}  tested/ws {d['pairs_tested'] / (a.window * a.worlds):7.1f} hits/ws {d['pairs_colliding'] / (a.window * a.worlds):6.1f}", flush=True)
