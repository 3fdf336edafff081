"""Host-side mirror of the reference API: prototypes, configuration, scene builders (no GPU)."""
import math

import numpy as np
import pytest

from common import P, scenes, Configuration, StepConfiguration, to_native_config
from physicssimulator_b200 import sharding


def test_prototype_defaults_match_reference():  # RigidBodyPrototype.fs:29-50
    p = P.createDefaultCube(1.0)
    assert (p.ElasticityCoeff, p.StaticFrictionCoeff, p.KineticFrictionCoeff, p.Mass, p.UseGravity) == (0.1, 0.5, 0.6, 1.0, True)


def test_build_inertia_and_force():
    b = P.build_world([P.createDefaultBox(1, 2, 3).With(Mass=12.0), P.createDefaultSphere(0.5).With(Mass=5.0),
                       P.createDefaultCube(1).With(Mass=P.INFINITE, UseGravity=False)])
    assert np.allclose(b["inertia_inv"][0], [1 / 13.0, 1 / 10.0, 1 / 5.0])  # Box.fs:13-19
    assert np.allclose(b["inertia_inv"][1], 1.0 / (2.0 * 5.0 * 0.25 / 5.0))    # Sphere.fs:11
    assert np.all(b["inertia_inv"][2] == 0.0)                                  # 1/inf
    assert np.array_equal(b["force"][0], [0.0, 0.0, 12.0 * (9.81 * -1.0)])
    assert np.all(b["force"][2] == 0.0)


def test_validation_errors_like_the_reference():
    with pytest.raises(ValueError):  # RigidBodyPrototype.fs:61-67
        P.build_world([P.createDefaultCube(1).With(Mass=P.INFINITE)])
    with pytest.raises(ValueError):  # RigidBody.fs:58-59
        P.build_world([P.createDefaultCube(1).With(ElasticityCoeff=1.5)])
    with pytest.raises(ValueError):  # Box.fs:51-54
        P.Box(-1, 1, 1)


def test_config_mapping():
    cfg = Configuration(StepConfig=StepConfiguration(EnableFriction=False, CollisionSolverIterationCount=7), SimulationStepInterval=0.02)
    n = to_native_config(cfg, record_contacts=True)
    assert (n.dt, n.solver_iterations, n.enable_friction, n.record_contacts, n.restore_joints) == (0.02, 7, 0, 1, 0)
    assert (n.epsilon, n.baumgarte, n.allowed_penetration, n.max_correction_velocity) == (1e-4, 0.2, 0.001, 1.0)


def test_scenes_are_deterministic_and_shaped():
    a, _ = scenes.c2_spheres64(5)
    b, _ = scenes.c2_spheres64(5)
    assert len(a) == 65 and all(x.Position == y.Position and x.Velocity == y.Velocity for x, y in zip(a, b))
    c, _ = scenes.c2_spheres64(6)
    assert any(x.Position != y.Position for x, y in zip(a, c))
    p3, _ = scenes.c3_mixed128(0)
    assert len(p3) == 128 and sum(isinstance(x.Kind, P.Box) for x in p3) == 65
    p4, cfg4, j4 = scenes.c4_chain32(0)
    assert len(p4) == 34 and len(j4) == 32 and cfg4.RestoreJoints
    s, cfg = scenes.stack_scene()
    assert len(s) == 5 and not cfg.BroadPhaseCollisionDetectionKind.IsDummy
    big = scenes.c5_big(1000, 100)
    assert len(big["mass"]) == 1101 and np.isinf(big["mass"][0])


def test_splitmix64_known_values():
    r = scenes.SplitMix64(0)
    assert [r.next_u64() for _ in range(3)] == [0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4, 0x06C45D188009454F]


def test_shard_range_partitions():
    for n in (1, 7, 4096, 16384):
        for g in (1, 2, 4, 8):
            r = [sharding.shard_range(n, k, g) for k in range(g)]
            assert r[0][0] == 0 and r[-1][1] == n and all(r[k][1] == r[k + 1][0] for k in range(g - 1))
    with pytest.raises(ValueError):
        sharding.shard_range(10, 3, 2)


def test_simulator_needs_three_prototypes():  # SimulatorState.fs:63-64
    from physicssimulator_b200.simulator import Simulator
    with pytest.raises(KeyError):
        Simulator([P.createDefaultCube(1), P.createDefaultCube(1)])
