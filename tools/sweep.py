#!/usr/bin/env python
"""Time one workload under several (path, margin scale, hit cap) settings from the same settled state.
Each variant: fresh batch, state restored, `--steps` single-step calls timed with CUDA events on torch's stream
is not possible (the library owns its stream), so wall clock around synchronous ps_batch_step calls."""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from physicssimulator_b200.simulator import BatchSimulator

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c3"); ap.add_argument("--worlds", type=int, default=16384)
ap.add_argument("--settle", type=int, default=200); ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--variants", default="wide:2:16,fused:2:16", help="comma list of path:mscale:ncap[:tile]")
ap.add_argument("--prof", action="store_true", help="PS_B200_WIDE_PROF for the wide variants (serialises)")
a = ap.parse_args()
built, cfg, joints = bench.build_worlds(a.workload, 0, a.worlds, min(16, os.cpu_count() or 1))
sim = BatchSimulator(None, cfg, joints=joints, device=0, _prebuilt=built)
sim.batch.step(a.settle)
st = sim.batch.get_state()
sim.close()
keys = ("pairs_tested", "pairs_colliding", "fallback_margin_steps", "margin_verified_steps", "fallback_capacity_steps")
ref = None
for v in a.variants.split(","):
    f = v.split(":")
    path, ms, nc = f[0], f[1], f[2]
    os.environ["PS_B200_PATH"] = path
    os.environ["PS_B200_MSCALE"] = ms
    os.environ["PS_B200_NCAP"] = nc
    if len(f) > 3: os.environ["PS_B200_TILE"] = f[3]
    else: os.environ.pop("PS_B200_TILE", None)
    if a.prof and path == "wide": os.environ["PS_B200_WIDE_PROF"] = "1"
    else: os.environ.pop("PS_B200_WIDE_PROF", None)
    sim = BatchSimulator(None, cfg, joints=joints, device=0, _prebuilt=built)
    sim.batch.set_state(st)
    sim.batch.step(2)  # first step after create takes the static fallback
    s0 = sim.Stats()
    t = time.perf_counter()
    for _ in range(a.steps):
        sim.batch.step(1)
    dt = time.perf_counter() - t
    s1 = sim.Stats()
    out = sim.batch.get_state()
    d = {k: (s1[k] - s0[k]) / (a.steps * a.worlds) for k in keys}
    same = None
    if ref is None: ref = out
    else: same = all(np.array_equal(out[k].view(np.uint64), ref[k].view(np.uint64)) for k in ("pos", "vel", "rot"))
    print(f"{a.workload} worlds={a.worlds} {v:18s} ms/step={1e3 * dt / a.steps:8.3f}  per world-step: " +
          " ".join(f"{k}={d[k]:.4g}" for k in keys) + f"  same_bits_as_first={same}", flush=True)
    sim.close()
