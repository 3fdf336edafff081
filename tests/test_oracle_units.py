"""T5/T6-style unit checks on the oracle's building blocks: the Sutherland-Hodgman clipper with the reference's
quirks (Q9) and the per-pair solver's observable invariants."""
import numpy as np

from common import P, scenes, Configuration, StepConfiguration, to_native_config
from oracle import oracle
from oracle.oracle import OracleWorld

EPS = 1e-4


def test_clip_keeps_inside_polygon():
    sq = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0)]
    planes_n = [(1, 0, 0), (0, 1, 0)]  # keep x >= -5, y >= -5
    out = oracle.clip(EPS, planes_n, [-5.0, -5.0], sq)
    assert np.array_equal(out, np.array(sq, dtype=float)[[0, 1, 2, 3]])  # in/in edges emit the END point: (last,first) first


def test_clip_out_in_emits_end_then_intersection_q9():
    # GraphicsUtils.fs:100: out->in yields [endPoint; intersection] (not the textbook order)
    sq = [(0, 0, 0), (2, 0, 0), (2, 2, 0), (0, 2, 0)]
    out = oracle.clip(EPS, [(1, 0, 0)], [1.0], sq)  # keep x >= 1
    # edges: (v3,v0) out/out -> []; (v0,v1) out/in -> [v1, X(1,0)]; (v1,v2) in/in -> [v2]; (v2,v3) in/out -> [X(1,2)]
    assert np.allclose(out, [(2, 0, 0), (1, 0, 0), (2, 2, 0), (1, 2, 0)])


def test_clip_parallel_edge_drops_intersection_and_distinct():
    tri = [(0, 0, 0), (1, 0, 0), (1, 0, 0)]  # duplicate vertex: List.distinct removes it
    out = oracle.clip(EPS, [(0, 0, 1)], [-1.0], tri)
    assert len(out) == 2
    # an edge parallel to the plane (|n.ab| <= eps) yields no intersection point
    quad = [(0, 0, 0), (1, 0, 0), (1, 0, 1e-9), (0, 0, 1e-9)]
    out = oracle.clip(EPS, [(0, 0, 1)], [5e-10], quad)
    assert all(p[2] >= 5e-10 for p in out)


def test_clip_everything_outside_is_empty():
    sq = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0)]
    out = oracle.clip(EPS, [(1, 0, 0)], [10.0], sq)
    assert len(out) == 0


def _pair(a, b, cfg=None):
    filler = P.createDefaultSphere(0.1).With(Position=(40, 40, 40), UseGravity=False)
    cfg = cfg or Configuration()
    return OracleWorld(P.build_world([a, b, filler]), to_native_config(cfg))


def test_solver_momentum_is_conserved_between_two_dynamic_bodies():
    # equal and opposite impulses at the same point (CollisionResponse.fs:98-113): linear momentum is conserved
    a = P.createDefaultCube(1.0).With(Position=(0, 0, 0), Velocity=(1.0, 0.2, 0), Mass=3.0, UseGravity=False)
    b = P.createDefaultCube(1.0).With(Position=(0.95, 0.1, 0.05), Velocity=(-1.0, 0, 0.1), Mass=5.0, UseGravity=False, Yaw=0.2)
    w = _pair(a, b)
    p0 = 3.0 * np.array([1.0, 0.2, 0]) + 5.0 * np.array([-1.0, 0, 0.1])
    w.step(1)
    st = w.get_state()
    assert len(w.contacts()["pairs"]) == 1
    p1 = 3.0 * st["vel"][0] + 5.0 * st["vel"][1]
    assert np.allclose(p0, p1, atol=1e-12)


def test_solver_static_partner_keeps_zero_velocity_but_gains_angular_momentum_q2():
    g = P.createDefaultBox(10, 10, 1).With(Mass=P.INFINITE, UseGravity=False, Position=(0, 0, -0.5))
    c = P.createDefaultCube(1.0).With(Position=(1.0, 0.5, 0.49), Velocity=(0.5, 0, -1.0), Mass=2.0)
    w = _pair(g, c)
    w.step(1)
    st = w.get_state()
    assert np.all(st["vel"][0] == 0.0) and np.any(st["ang_mom"][0] != 0.0)
    assert st["vel"][1][2] > -1.0  # the normal impulse acted on the cube


def test_friction_off_leaves_tangential_velocity():
    g = P.createDefaultBox(10, 10, 1).With(Mass=P.INFINITE, UseGravity=False, Position=(0, 0, -0.5))
    c = P.createDefaultCube(1.0).With(Position=(0, 0, 0.49), Velocity=(0.7, 0, -0.2), Mass=2.0, UseGravity=False)
    on = _pair(g, c)
    off = _pair(g, c, Configuration(StepConfig=StepConfiguration(EnableFriction=False)))
    on.step(1); off.step(1)
    assert off.get_state()["vel"][1][0] == 0.7
    assert on.get_state()["vel"][1][0] < 0.7


def test_iteration_count_zero_means_no_impulse():
    g = P.createDefaultBox(10, 10, 1).With(Mass=P.INFINITE, UseGravity=False, Position=(0, 0, -0.5))
    c = P.createDefaultCube(1.0).With(Position=(0, 0, 0.49), Velocity=(0, 0, -1.0), Mass=2.0, UseGravity=False)
    w = _pair(g, c, Configuration(StepConfig=StepConfiguration(CollisionSolverIterationCount=0)))
    w.step(1)
    assert w.get_state()["vel"][1][2] == -1.0 and len(w.contacts()["pairs"]) == 1
