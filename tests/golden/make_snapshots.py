#!/usr/bin/env python
"""Generates tests/golden/*.pstj (format: include/ps_snapshot.h) from the CPU oracle — two worlds per file, a frame
at step 0 and at every mark, with ordered candidate lists, colliding pairs, contacts and feature ids.

Like the .npz fixtures these pin the ORACLE (the reference cannot run in this image); the same files are what the
headless .NET driver of INTEGRATION.md §6 writes on a box with the SDK, so that `tools/compare_snapshots.py` can put a
measured tolerance where DESIGN.md §2 says "parity unpinned".  Re-run after a deliberate oracle change:
    python tests/golden/make_snapshots.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from common import P, scenes, Configuration, to_native_config  # noqa: E402
from oracle.oracle import OracleWorld  # noqa: E402
from physicssimulator_b200 import snapshot  # noqa: E402

# name -> (list of worlds (prototype lists), Configuration, marks)
CASES = {
    "snap_c2_two_worlds": (lambda: [scenes.c2_spheres64(0)[0], scenes.c2_spheres64(1)[0]], scenes.c2_spheres64(0)[1], [1, 2, 20, 60]),
    # the Gui's Stack scene in its OWN configuration (SpatialTree broad phase: candidate order of the persistent 2^3-tree)
    "snap_stack_scene_tree": (lambda: [scenes.stack_scene()[0], scenes.stack_scene()[0]], scenes.stack_scene()[1], [1, 2, 25, 80]),
    "snap_c3_small_two_worlds": (lambda: [scenes.c3_mixed128(0, n_boxes=10, n_spheres=8)[0], scenes.c3_mixed128(3, n_boxes=12, n_spheres=6)[0]],
                                 Configuration(), [1, 2, 15, 40]),
}


def oracle_steppers(worlds, cfg):
    """(bodies SoA of all worlds, offsets, step, get_state, get_contacts, get_candidates) over one OracleWorld per world"""
    sc = cfg.StepConfig
    built = [P.build_world(list(w), sc.GravitationalAccelerationMagnitude, sc.GravityDirection) for w in worlds]
    bodies, off = P.concat_worlds(built)
    ncfg = to_native_config(cfg, record_contacts=True)
    ws = [OracleWorld(b, ncfg) for b in built]
    tc = cfg.BroadPhaseCollisionDetectionKind.SpatialTree
    if tc is not None:  # the oracle's persistent tree (pinned by the reference's known-answer facts)
        for w in ws:
            assert w.use_spatial_tree(tc.LeafCapacity, tc.MaxDepth, tc.SpaceBoundariesMin, tc.SpaceBoundariesMax) == 0

    def step(n):
        for w in ws:
            w.step(n)

    def get_state():
        sts = [w.get_state() for w in ws]
        return {k: np.concatenate([s[k] for s in sts]) for k in sts[0]}

    return bodies, off, ncfg, step, get_state, (lambda w: ws[w].contacts()), (lambda w: ws[w].candidates())


def generate(name):
    mk, cfg, marks = CASES[name]
    bodies, off, ncfg, step, get_state, contacts, cands = oracle_steppers(mk(), cfg)
    return snapshot.record(bodies, off, ncfg, marks, step, get_state, contacts, cands,
                           producer=snapshot.PRODUCER_ORACLE, scene=name)


if __name__ == "__main__":
    for name in CASES:
        snapshot.save(os.path.join(HERE, name + ".pstj"), generate(name))
        print("wrote", name, os.path.getsize(os.path.join(HERE, name + ".pstj")), "bytes")
