"""CPU check of the DEVICE physics routines: physicssimulator_b200/csrc/ps_physics.cuh compiled with g++
(tests/host_harness) must reproduce the oracle bit for bit — states, colliding sets, contacts and feature ids.
This runs without a GPU; the same comparison through the real kernels is tests/test_gpu_parity.py."""
import math

import numpy as np
import pytest

from common import (P, scenes, Configuration, StepConfiguration, to_native_config, assert_state_bits, assert_contacts_equal,
                    harness_step)
from oracle.oracle import OracleWorld


def _run(protos, cfg, chunks, what):
    b = P.build_world(protos, cfg.StepConfig.GravitationalAccelerationMagnitude, cfg.StepConfig.GravityDirection)
    nc = to_native_config(cfg)
    o = OracleWorld(b, nc)
    cur = {k: v.copy() for k, v in b.items()}
    done = 0
    for n in chunks:
        o.step(n)
        # both forms of the narrow phase: the compact one of the fused kernel and the inlined one of the wide kernels
        st, rec, cnt = harness_step(cur, nc, n)
        st_f, rec_f, cnt_f = harness_step(cur, nc, n, fast=True)
        st_s, rec_s, cnt_s = harness_step(cur, nc, n, mode=1 | 2 | 8)  # + straight-line integration and straight-line solve
        done += n
        for s_, r_, c_, form in ((st, rec, cnt, "compact"), (st_f, rec_f, cnt_f, "inlined"), (st_s, rec_s, cnt_s, "straight-line")):
            assert_state_bits(o.get_state(), s_, f"{what}/{form} @ {done}")
            assert_contacts_equal(o.contacts(), r_, f"{what}/{form} @ {done}")
            assert c_["overflow"] == 0
        cur.update(st)


def test_stack_scene():
    _run(*scenes.stack_scene(), [1, 9, 90, 200], "stack")


def test_domino_scene():
    _run(*scenes.domino_scene(), [1, 9, 60], "domino")


def test_c1_stack20():
    _run(*scenes.c1_stack20(), [1, 24, 100], "c1")


def test_c2_world():
    p, c = scenes.c2_spheres64(3)
    _run(p, c, [1, 49, 150], "c2")


def test_c3_world():
    p, c = scenes.c3_mixed128(1, n_boxes=32, n_spheres=31)
    _run(p, c, [1, 29, 90], "c3")


@pytest.mark.parametrize("integrator", [0, 1])
def test_free_flight(integrator):
    rng = scenes.SplitMix64(5 + integrator)
    protos = []
    for k in range(12):
        p = P.createDefaultBox(rng.uniform(.2, 1), rng.uniform(.2, 1), rng.uniform(.2, 1)) if k % 2 else P.createDefaultSphere(rng.uniform(.2, .5))
        protos.append(p.With(Position=(10.0 * k, 0, 0), Velocity=(rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(-1, 1)),
                             Yaw=rng.uniform(-3, 3), Pitch=rng.uniform(-3, 3), Roll=rng.uniform(-3, 3)))
    cfg = Configuration(StepConfig=StepConfiguration(RigidBodyIntegratorKind=integrator))
    b = P.build_world(protos)
    b["ang_mom"][:] = np.array([[rng.uniform(-1, 1) for _ in range(3)] for _ in range(12)])
    b["ang_mom"][5] = 0.0
    b["torque"][3] = (0.1, -0.2, 0.05)
    nc = to_native_config(cfg)
    o = OracleWorld(b, nc)
    o.step(500)
    st, _, _ = harness_step(b, nc, 500)
    assert_state_bits(o.get_state(), st, "free flight")


def test_random_box_and_sphere_pairs_incl_axis_aligned():
    """many small random worlds: axis-aligned boxes (exact ties, zero components), rotated boxes, spheres —
    exercises the +-axis pairing, the early 'separated' exit and the zero-cross-product skip of the device SAT"""
    rng = scenes.SplitMix64(2024)
    for trial in range(120):
        protos = [P.createDefaultBox(6, 6, 1).With(Mass=P.INFINITE, UseGravity=False, Position=(0, 0, -0.5),
                                                   Yaw=(math.pi if trial % 3 == 0 else 0.0))]
        for k in range(5):
            aligned = (trial % 2 == 0)
            ang = (lambda: 0.0) if aligned else (lambda: rng.uniform(-3.1, 3.1))
            pos = (rng.uniform(-0.8, 0.8), rng.uniform(-0.8, 0.8), rng.uniform(0.3, 1.6))
            if aligned:
                pos = (round(pos[0] * 4) / 4, round(pos[1] * 4) / 4, round(pos[2] * 4) / 4)
            if k == 4 and trial % 4 == 1:
                protos.append(P.createDefaultSphere(rng.uniform(0.2, 0.5)).With(Position=pos, Mass=rng.uniform(1, 9)))
            else:
                protos.append(P.createDefaultBox(rng.uniform(0.3, 1.0), rng.uniform(0.3, 1.0), rng.uniform(0.3, 1.0)).With(
                    Position=pos, Yaw=ang(), Pitch=ang(), Roll=ang(), Mass=rng.uniform(1, 50),
                    Velocity=(rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(-1, 1))))
        _run(protos, Configuration(), [1, 3, 12], f"random trial {trial}")


def test_team_form_of_box_box_and_straight_line_integrate():
    """ps_coop.cuh (8 lanes per box pair, here 8 host threads behind a barrier) must give the literal routine's manifold —
    normals, points, penetrations, feature ids, contact order — or ask for the literal routine; the straight-line
    integration (integrate_a) must give integrate()'s bits.  Stacked axis-aligned boxes put exact ties on the arg-min."""
    rng = scenes.SplitMix64(77)
    total = {"team_runs": 0, "team_literal": 0, "team_mismatch": 0, "pairs_colliding": 0}
    worlds = [scenes.c3_mixed128(1, n_boxes=24, n_spheres=8), scenes.c1_stack20()]
    for trial in range(6):
        protos = [P.createDefaultBox(6, 6, 1).With(Mass=P.INFINITE, UseGravity=False, Position=(0, 0, -0.5))]
        for k in range(6):
            aligned = trial % 2 == 0
            ang = (lambda: 0.0) if aligned else (lambda: rng.uniform(-3.1, 3.1))
            pos = (rng.uniform(-0.7, 0.7), rng.uniform(-0.7, 0.7), rng.uniform(0.3, 1.4))
            if aligned:
                pos = tuple(round(x * 4) / 4 for x in pos)
            protos.append(P.createDefaultBox(rng.uniform(0.3, 1.0), rng.uniform(0.3, 1.0), rng.uniform(0.3, 1.0)).With(
                Position=pos, Yaw=ang(), Pitch=ang(), Roll=ang(), Mass=rng.uniform(1, 50)))
        worlds.append((protos, Configuration()))
    for protos, cfg in worlds:
        b = P.build_world(protos, cfg.StepConfig.GravitationalAccelerationMagnitude, cfg.StepConfig.GravityDirection)
        nc = to_native_config(cfg)
        o = OracleWorld(b, nc)
        o.step(25)
        st, rec, cnt = harness_step(b, nc, 25, mode=1 | 2 | 4)
        assert_state_bits(o.get_state(), st, "team/straight-line")
        assert_contacts_equal(o.contacts(), rec, "team/straight-line")
        for k in total:
            total[k] += cnt[k]
    assert total["team_runs"] > 1000 and total["pairs_colliding"] > 100
    assert total["team_mismatch"] == 0, total
    assert total["team_literal"] * 20 <= total["team_runs"], total
