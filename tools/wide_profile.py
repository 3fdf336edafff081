#!/usr/bin/env python
"""Per-stage times of the batch-wide path (PS_B200_WIDE_PROF=1 makes the library time every stage with CUDA events;
that serialises the stream, so the sum is an upper bound of a normal step)."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from physicssimulator_b200.simulator import BatchSimulator
ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c3"); ap.add_argument("--worlds", type=int, default=16384)
ap.add_argument("--settle", type=int, default=200); ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
built, cfg, joints = bench.build_worlds(a.workload, 0, a.worlds, min(16, os.cpu_count() or 1))
os.environ.pop("PS_B200_WIDE_PROF", None)
sim = BatchSimulator(None, cfg, joints=joints, device=0, _prebuilt=built)
sim.batch.step(a.settle)
st = sim.batch.get_state()
sim.close()
os.environ["PS_B200_WIDE_PROF"] = "1"
os.environ["PS_B200_PATH"] = "wide"
sim = BatchSimulator(None, cfg, joints=joints, device=0, _prebuilt=built)
sim.batch.set_state(st)
sim.batch.step(a.steps)
print(sim.Stats())
sim.close()
