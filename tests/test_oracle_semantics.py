"""The oracle reproduces the reference behaviours SURVEY.md Appendix B lists (read off the F# source): these
checks pin the READING, so that a later "fix" of a quirk shows up as a failure."""
import math

import numpy as np

from common import P, scenes, Configuration, StepConfiguration, to_native_config
from oracle.oracle import OracleWorld


def _world(protos, cfg=None):
    cfg = cfg or Configuration()
    return OracleWorld(P.build_world(protos, cfg.StepConfig.GravitationalAccelerationMagnitude, cfg.StepConfig.GravityDirection), to_native_config(cfg))


_fill = [0]


def _filler():
    """far-away padding bodies (kept apart from each other: coincident spheres are NaN with friction, Q20)"""
    _fill[0] += 1
    return P.createDefaultSphere(0.1).With(Position=(50.0 + 3.0 * _fill[0], 50, 50), UseGravity=False)


def test_free_fall_is_semi_implicit_euler():  # Particle.fs:25-30
    w = _world([P.createDefaultSphere(0.2).With(Position=(0, 0, 10)), _filler(), _filler()])
    w.step(10)
    st = w.get_state()
    v = 0.0
    z = 10.0
    for _ in range(10):
        v = v + 0.01 * (1.0 * (9.81 * -1.0) * (1.0 / 1.0))
        z = z + 0.01 * v
    assert st["vel"][0, 2] == v and st["pos"][0, 2] == z


def test_q3_sphere_sphere_normal_points_second_to_first():  # CollisionDetection.fs:189,195
    a = P.createDefaultSphere(0.5).With(Position=(0, 0, 0), UseGravity=False)
    b = P.createDefaultSphere(0.5).With(Position=(0.9, 0, 0), UseGravity=False)
    w = _world([a, b, _filler()])
    c = w.detect(0, 1)
    assert c["normal"].shape == (1, 3) and c["normal"][0, 0] == -1.0  # from second to first
    assert math.isclose(c["penetration"][0], 0.1, rel_tol=1e-12) and c["penetration"][0] > 0
    # approaching spheres get no impulse: they pass into each other
    a2 = a.With(Velocity=(1, 0, 0)); b2 = b.With(Velocity=(-1, 0, 0))
    w = _world([a2, b2, _filler()], Configuration(StepConfig=StepConfiguration(EnableFriction=False)))
    w.step(1)
    st = w.get_state()
    assert st["vel"][0, 0] == 1.0 and st["vel"][1, 0] == -1.0


def test_q4_sphere_on_plus_z_side_of_a_box_gets_the_wrong_face():  # SAT.fs:114-151, Box.fs:42-49
    box = P.createDefaultBox(4, 4, 1).With(Mass=P.INFINITE, UseGravity=False)
    up = P.createDefaultSphere(0.3).With(Position=(0, 0, 0.75), UseGravity=False)
    down = P.createDefaultSphere(0.3).With(Position=(0, 0, -0.75), UseGravity=False)
    w = _world([box, up, down])
    cu, cd = w.detect(0, 1), w.detect(0, 2)
    # both report face 0 (local -z, the first of the +-z tie) with the unsigned axis as "normal from box"
    assert (cu["feature"][0] >> 4) & 7 == 0 and (cd["feature"][0] >> 4) & 7 == 0
    assert cu["normal"][0, 2] == -1.0 and cd["normal"][0, 2] == -1.0
    # sphere contacts carry POSITIVE penetration -> Baumgarte bias 0 (CollisionResponse.fs:35-41)
    assert cu["penetration"][0] > 0 and cd["penetration"][0] > 0


def test_q5_sphere_aabb_is_a_cube_of_side_radius():  # SimulatorObject.fs:59-64
    from oracle import oracle
    b = P.build_world([P.createDefaultSphere(0.4).With(Position=(1, 2, 3)), _filler(), _filler()])
    size, mn = oracle.aabbs(b, 0)
    assert np.all(size[0] == 0.4) and np.allclose(mn[0], [0.8, 1.8, 2.8])
    size, _ = oracle.aabbs(b, 1)
    assert np.all(size[0] == 0.8)


def test_q1_q2_k_plus_one_integrations_and_static_angular_momentum():
    protos, _ = scenes.stack_scene()
    w = _world(protos)
    w.step(300)
    st = w.get_state()
    assert np.linalg.norm(st["ang_mom"][0]) > 1.0  # the static ground accumulates L (RigidBody.fs:194-197)
    assert np.all(st["vel"][0] == 0) and np.array_equal(st["pos"][0], [0, 0, -0.5])
    s = w.stats()
    assert s["pairs_tested"] == 300 * 10  # Dummy: C(5,2) candidates each step (BroadPhase.fs:33-39)


def test_q11_yaw_pitch_roll_formula_quirk():  # Matrix3.fs:90-108
    R = P.from_yaw_pitch_roll(0.3, 0.5, 0.7)
    cA, sA, cB, sB, cG, sG = math.cos(.3), math.sin(.3), math.cos(.5), math.sin(.5), math.cos(.7), math.sin(.7)
    assert R[1, 1] == cA * sB * sG + cA * cG  # cosA where a rotation matrix has sinA
    assert abs(np.linalg.det(R) - 1.0) > 1e-3  # hence not a rotation in general


def test_q20_coincident_sphere_centres_give_nan_with_friction():  # Friction.fs:57-60
    a = P.createDefaultSphere(0.5).With(Position=(0, 0, 5))
    b = P.createDefaultSphere(0.5).With(Position=(0, 0, 5))
    w = _world([a, b, _filler()])
    w.step(1)
    assert w.stats()["nan"] == 1
    w = _world([a, b, _filler()], Configuration(StepConfig=StepConfiguration(EnableFriction=False)))
    w.step(1)
    assert w.stats()["nan"] == 0


def test_spatial_tree_candidates_are_a_history_dependent_subset():  # BroadPhase.fs:62-70,100-127
    protos, cfg = scenes.stack_scene()
    d = _world(protos)
    t = _world(protos)
    tc = cfg.BroadPhaseCollisionDetectionKind.SpatialTree
    assert t.use_spatial_tree(tc.LeafCapacity, tc.MaxDepth, tc.SpaceBoundariesMin, tc.SpaceBoundariesMax) == 0
    d.step(1); t.step(1)
    cd = {tuple(x) for x in d.candidates()}
    ct = {tuple(x) for x in t.candidates()}
    assert len(cd) == 10 and ct <= cd and len(ct) >= 1
    # an object outside the bounds cannot be constructed (BroadPhase.fs:83-98 throws)
    far = _world(protos[:4] + [protos[4].With(Position=(100, 0, 0))])
    assert far.use_spatial_tree(2, 10, (-15, -15, -15), (15, 15, 15)) == -1


def test_joint_restore_as_written_is_jacobi_last_writer_wins():  # JointsRestorer.fs:9-26
    p, cfg, joints = scenes.c4_chain32(0, n_links=4)
    b = P.build_world(p)
    jl = [(a, bb) + P.ball_socket_local_anchors(b, a, bb, anchor) for (a, bb, anchor) in joints]
    w = OracleWorld(b, to_native_config(cfg), joints=jl)
    w2 = OracleWorld(b, to_native_config(cfg.With(RestoreJoints=False)), joints=jl)
    w.step(3); w2.step(3)
    assert not np.array_equal(w.get_state()["vel"], w2.get_state()["vel"])
