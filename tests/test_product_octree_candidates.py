"""The product's host-side SpatialTree (csrc/ps_octree.hpp, compiled into the test harness) must hand out the oracle's
candidate list — same pairs, same ORDER — at every step of a SpatialTree-mode rollout: the list is history dependent
(leaves split and never merge, dynamic ids are re-filed every step), so this walks the scenes for many steps.
The oracle's tree is pinned by the reference's own known-answer facts (tests/test_oracle_octree.py, both trees)."""
import numpy as np
import pytest

from common import P, scenes, Configuration, to_native_config, ProductTree
from oracle.oracle import OracleWorld, aabbs
from physicssimulator_b200.configuration import BroadPhaseCollisionDetectionKind, SpatialTreeConfiguration


def _rollout(protos, cfg, steps, what):
    tc = cfg.BroadPhaseCollisionDetectionKind.SpatialTree
    sc = cfg.StepConfig
    b = P.build_world(protos, sc.GravitationalAccelerationMagnitude, sc.GravityDirection)
    o = OracleWorld(b, to_native_config(cfg))
    assert o.use_spatial_tree(tc.LeafCapacity, tc.MaxDepth, tc.SpaceBoundariesMin, tc.SpaceBoundariesMax) == 0
    n = len(b["mass"])
    static = np.isinf(b["mass"])
    t = ProductTree(tc.LeafCapacity, tc.MaxDepth, list(zip(tc.SpaceBoundariesMin, tc.SpaceBoundariesMax)))

    def feed(state):
        cur = dict(b)
        cur.update(state)
        size, mn = aabbs(cur)
        for i in range(n):
            t.set_extent(i, [(size[i, k], mn[i, k]) for k in range(3)])

    feed({"pos": b["pos"], "rot": b["rot"]})
    for i in range(n):  # BroadPhase.init: every id, ascending
        assert t.insert(i) == 0
    nonempty = 0
    for s in range(steps):
        feed(o.get_state())  # extents at the step-start state
        for i in range(n):   # BroadPhase.update: dynamic ids out, then back in ascending order
            if not static[i]:
                t.remove(i)
        for i in range(n):
            if not static[i]:
                t.insert(i)  # tryInsert: outside the space -> absent
        mine = t.candidates(static)
        o.step(1)
        ref = [tuple(p) for p in o.candidates().tolist()]
        assert mine == ref, f"{what}: candidate list differs at step {s + 1}"
        nonempty += bool(ref)
    assert nonempty > 0


def test_stack_scene_candidates():
    _rollout(*scenes.stack_scene(), 160, "stack")


def test_domino_scene_candidates():
    _rollout(*scenes.domino_scene(), 60, "domino")


def test_c1_in_the_stack_scene_tree():
    protos, _ = scenes.c1_stack20()
    cfg = Configuration(BroadPhaseCollisionDetectionKind=BroadPhaseCollisionDetectionKind(SpatialTreeConfiguration(LeafCapacity=2, MaxDepth=10)))
    _rollout(protos, cfg, 120, "c1")


def test_small_space_drops_leavers():
    """bodies that fall out of the tree's space are silently absent from the candidates (tryInsert, BroadPhase.fs:117-120)"""
    protos, _ = scenes.c3_mixed128(2, n_boxes=10, n_spheres=10)
    cfg = scenes.c3_mixed128(2)[1].With(BroadPhaseCollisionDetectionKind=BroadPhaseCollisionDetectionKind(
        SpatialTreeConfiguration(LeafCapacity=3, MaxDepth=6, SpaceBoundariesMin=(-30.0, -30.0, -3.0), SpaceBoundariesMax=(30.0, 30.0, 30.0))))
    _rollout(protos, cfg, 80, "c3 slice")
