#!/usr/bin/env python
"""Generates tests/golden/*.npz from the CPU oracle (oracle/ps_oracle.cpp).

The reference (F#/.NET) cannot be run in this image, so these vectors pin the ORACLE against regressions and
give the GPU tests a fixed target that does not depend on rebuilding the oracle; they are NOT outputs of the
real reference (parity vs the .NET runtime is unpinned).  Re-run after a deliberate oracle change:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from common import P, scenes, Configuration, StepConfiguration, to_native_config  # noqa: E402
from oracle.oracle import OracleWorld  # noqa: E402

CASES = {
    "stack_scene": (lambda: scenes.stack_scene()[0], Configuration(), [1, 10, 100, 400]),
    "domino_scene": (lambda: scenes.domino_scene()[0], Configuration(), [1, 10, 60]),
    "c1_stack20": (lambda: scenes.c1_stack20()[0], Configuration(), [1, 10, 100]),
    "c2_world0": (lambda: scenes.c2_spheres64(0)[0], scenes.c2_spheres64(0)[1], [1, 50, 200]),
    "c3_world0_small": (lambda: scenes.c3_mixed128(0, n_boxes=20, n_spheres=20)[0], Configuration(), [1, 30, 120]),
    "c3_second_order": (lambda: scenes.c3_mixed128(1, n_boxes=12, n_spheres=12)[0],
                        Configuration(StepConfig=StepConfiguration(RigidBodyIntegratorKind=1)), [1, 30, 90]),
}


def generate(name):
    mk, cfg, marks = CASES[name]
    protos = mk()
    w = OracleWorld(P.build_world(protos, cfg.StepConfig.GravitationalAccelerationMagnitude, cfg.StepConfig.GravityDirection), to_native_config(cfg))
    out = {"marks": np.array(marks)}
    done = 0
    for m in marks:
        w.step(m - done)
        done = m
        st = w.get_state()
        c = w.contacts()
        for k, v in st.items():
            out[f"s{m}_{k}"] = v
        out[f"s{m}_pairs"] = c["pairs"]
        out[f"s{m}_first"] = c["first"]
        out[f"s{m}_feature"] = c["feature"]
        out[f"s{m}_penetration"] = c["penetration"]
    return out


if __name__ == "__main__":
    for name in CASES:
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **generate(name))
        print("wrote", name)
