#!/usr/bin/env python
"""Config 1 of BASELINE.json: one Stack-style world (ground + 20 cubes + 4 spheres, friction), 1000 steps, on the GPU
stepper (one world = one warp: this is the latency of the path, not its throughput) and on the CPU oracle."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from physicssimulator_b200 import scenes, prototypes as P
from physicssimulator_b200.configuration import to_native_config
from physicssimulator_b200.simulator import BatchSimulator
from oracle.oracle import OracleWorld
protos, cfg = scenes.c1_stack20()
sim = BatchSimulator([protos], cfg, device=0)
sim.Step(10)
sim.Reset()
t = time.perf_counter(); sim.Step(1000); gpu_s = time.perf_counter() - t
o = OracleWorld(P.build_world(protos), to_native_config(cfg))
t = time.perf_counter(); o.step(1000); cpu_s = time.perf_counter() - t
ref, got = o.get_state(), sim.WorldState(0)
same = all(np.array_equal(ref[k][1:].view(np.uint64), got[k][1:].view(np.uint64)) for k in ("pos", "vel", "rot", "ang_mom"))
print(json.dumps({"workload": "c1 stack20, 1000 steps, Dummy order", "gpu_s": gpu_s, "gpu_body_steps_per_s": 24 * 1000 / gpu_s,
                  "cpu_oracle_s": cpu_s, "cpu_body_steps_per_s": 24 * 1000 / cpu_s, "dynamic_state_bit_identical_after_1000_steps": bool(same),
                  "stats": sim.Stats()}))
