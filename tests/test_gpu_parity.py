"""Parity tests proper: the CUDA path, called through the C ABI (libps_b200.so), against the CPU oracle on
the same seeded inputs.  Bar: candidate/colliding pair sets and contact-feature ids bit-exact; state of every
DYNAMIC body bit-exact (same bits, not a tolerance); angular momentum of STATIC bodies (reference quirk Q2)
bit-exact on the exact serial path and within rel 1e-9 on the wavefront path (its pair sums are folded in
pair order but associated per pair)."""
import numpy as np
import pytest

from common import (P, scenes, Configuration, StepConfiguration, to_native_config, same_bits, assert_state_bits,
                    assert_contacts_equal, bits)

from physicssimulator_b200.simulator import BatchSimulator

pytestmark = pytest.mark.gpu

STATIC_L_RTOL = 1e-9


def _oracle_worlds(worlds, cfg, joints=None):
    from oracle.oracle import OracleWorld
    out = []
    for w, protos in enumerate(worlds):
        b = P.build_world(protos, cfg.StepConfig.GravitationalAccelerationMagnitude, cfg.StepConfig.GravityDirection)
        jl = None
        if joints:
            jl = []
            for (a, bb, anchor) in joints[w]:
                la, lb = P.ball_socket_local_anchors(b, a, bb, anchor)
                jl.append((a, bb, la, lb))
        o = OracleWorld(b, to_native_config(cfg), joints=jl)
        tc = cfg.BroadPhaseCollisionDetectionKind.SpatialTree
        if tc is not None:  # the oracle's persistent tree (pinned by the reference's known-answer facts)
            assert o.use_spatial_tree(tc.LeafCapacity, tc.MaxDepth, tc.SpaceBoundariesMin, tc.SpaceBoundariesMax) == 0
        out.append(o)
    return out


def _compare(sim, oracles, what, exact_static_L):
    st = sim.State()
    off = sim._offsets
    for w, o in enumerate(oracles):
        ref = o.get_state()
        mass = sim._worlds[w]["mass"]
        dyn = ~np.isinf(mass)
        got = {k: st[k][off[w]:off[w + 1]] for k in st}
        for k in ("pos", "vel", "rot"):
            assert same_bits(got[k], ref[k]), f"{what}: world {w} {k} differs"
        assert same_bits(got["ang_mom"][dyn], ref["ang_mom"][dyn]), f"{what}: world {w} L of dynamic bodies differs"
        if exact_static_L:
            assert same_bits(got["ang_mom"], ref["ang_mom"]), f"{what}: world {w} L of static bodies differs"
        else:
            np.testing.assert_allclose(got["ang_mom"][~dyn], ref["ang_mom"][~dyn], rtol=STATIC_L_RTOL, atol=1e-9,
                                       err_msg=f"{what}: world {w} static L")


def _run(worlds, cfg, chunks, what, joints=None, exact=False, contacts=True):
    from physicssimulator_b200.simulator import BatchSimulator
    sim = BatchSimulator(worlds, cfg, joints=joints, device=0, record_contacts=contacts, exact_all_pairs=exact)
    oracles = _oracle_worlds(worlds, cfg, joints)
    done = 0
    for n in chunks:
        sim.Step(n)
        for o in oracles:
            o.step(n)
        done += n
        _compare(sim, oracles, f"{what} @step {done}", exact)
        if contacts:
            for w, o in enumerate(oracles):
                assert_contacts_equal(o.contacts(), sim.Collisions(w), f"{what} world {w} @step {done}")
                assert np.array_equal(o.candidates(), sim.Candidates(w)), f"{what}: candidate list differs"
    stats = sim.Stats()
    assert stats["contact_overflow"] == 0
    sim.close()
    return stats


def test_t1_free_flight_both_integrators():
    """T1: K1 integration on random x, v, R, L incl. static bodies and L = 0 — bit-exact over 1000 steps."""
    rng = scenes.SplitMix64(11)
    for integ in (0, 1):
        protos = []
        for k in range(40):
            pos = (k * 5.0, rng.uniform(-1, 1), rng.uniform(-1, 1))  # far apart: no contacts
            vel = (rng.uniform(-2, 2), rng.uniform(-2, 2), rng.uniform(-2, 2))
            if k % 3 == 0:
                p = P.createDefaultSphere(rng.uniform(0.2, 0.5))
            else:
                p = P.createDefaultBox(rng.uniform(0.2, 1), rng.uniform(0.2, 1), rng.uniform(0.2, 1))
            p = p.With(Position=pos, Velocity=vel, Yaw=rng.uniform(-3, 3), Pitch=rng.uniform(-3, 3), Roll=rng.uniform(-3, 3),
                       Mass=rng.uniform(0.5, 20))
            if k % 7 == 0:
                p = p.With(Mass=P.INFINITE, UseGravity=False, Velocity=(0, 0, 0))
            protos.append(p)
        cfg = Configuration(StepConfig=StepConfiguration(RigidBodyIntegratorKind=integ))
        from physicssimulator_b200.simulator import BatchSimulator
        from oracle.oracle import OracleWorld
        sim = BatchSimulator([protos], cfg, device=0)
        b = P.build_world(protos)
        # spin: give every body an angular momentum through the public impulse call
        o = OracleWorld(b, to_native_config(cfg))
        for k in range(40):
            J, off = (rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(-1, 1)), (rng.uniform(-.3, .3), rng.uniform(-.3, .3), rng.uniform(-.3, .3))
            if k % 5 != 4:  # leave some with L = 0
                sim.ApplyImpulse(0, k, J, off)
                o.apply_impulse(k, J, off)
        for n in (1, 9, 90, 900):
            sim.Step(n)
            o.step(n)
            assert_state_bits(o.get_state(), sim.WorldState(0), f"free flight integrator {integ}")
        sim.close()


def test_t2_sphere_sphere():
    """T2: touching / overlapping / approaching spheres (Q3), friction on and off."""
    for friction in (True, False):
        protos = [P.createDefaultBox(40, 40, 1).With(Mass=P.INFINITE, UseGravity=False, Position=(0, 0, -0.5), Yaw=np.pi)]
        protos += [P.createDefaultSphere(0.5).With(Position=(-0.6, 0, 3), Velocity=(1, 0, 0)),
                   P.createDefaultSphere(0.4).With(Position=(0.6, 0.05, 3), Velocity=(-1, 0, 0), Mass=3.0),
                   P.createDefaultSphere(0.3).With(Position=(5, 0, 0.3)),            # resting on the ground
                   P.createDefaultSphere(0.3).With(Position=(5.55, 0, 0.3)),         # overlapping its neighbour
                   P.createDefaultSphere(0.3).With(Position=(8, 0, 0.3)), P.createDefaultSphere(0.3).With(Position=(8.6, 0, 0.3))]  # exactly touching
        cfg = Configuration(StepConfig=StepConfiguration(EnableFriction=friction))
        _run([protos], cfg, [1, 1, 3, 45, 150], f"sphere-sphere friction={friction}")


def test_t3_sphere_box_six_sides():
    """T3: a sphere against each side of a box, both id orders (Q4 face selection, normal inversion)."""
    dirs = [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)]
    worlds = []
    for order in (0, 1):
        for d in dirs:
            box = P.createDefaultBox(1.0, 1.2, 0.8).With(Mass=P.INFINITE, UseGravity=False, Position=(0, 0, 0), Yaw=0.3, Pitch=-0.2)
            sph = P.createDefaultSphere(0.3).With(Position=tuple(0.95 * x for x in d), Velocity=tuple(-0.5 * x for x in d), UseGravity=False)
            filler = P.createDefaultSphere(0.1).With(Position=(30, 30, 30), UseGravity=False)
            worlds.append([box, sph, filler] if order == 0 else [sph, box, filler])
    _run(worlds, Configuration(), [1, 2, 10, 40], "sphere-box sides")


def test_t4_box_box_cases():
    """T4: axis-aligned exact ties (Face1 wins), rotated face-face, edge-edge, stacked."""
    worlds = []
    g = P.createDefaultBox(15, 15, 1.01).With(Mass=P.INFINITE, UseGravity=False, Position=(0, 0, -0.5))
    worlds.append([g, P.createDefaultCube(1.0).With(Position=(0, 0, 0.5), Mass=10), P.createDefaultCube(1.0).With(Position=(0.2, 0.1, 1.5), Mass=10)])
    worlds.append([g, P.createDefaultCube(1.0).With(Position=(0, 0, 0.6), Yaw=0.4, Pitch=0.3, Roll=0.2, Mass=10),
                   P.createDefaultBox(0.5, 2.0, 0.3).With(Position=(0.1, 0.0, 1.6), Roll=0.7, Mass=4)])
    # edge-edge: two boxes rotated 45 deg about different axes, approaching
    worlds.append([g, P.createDefaultCube(1.0).With(Position=(0, 0, 3), Yaw=np.pi / 4, UseGravity=False, Velocity=(0.3, 0, 0)),
                   P.createDefaultCube(1.0).With(Position=(1.38, 0, 3.0), Pitch=np.pi / 4, UseGravity=False, Velocity=(-0.3, 0, 0))])
    _run(worlds, Configuration(), [1, 1, 8, 40, 100], "box-box")


def test_t7_scenes_ordered_world_step():
    """T7: Stack / Domino / C1 worlds: k+1 integrations (Q1), write-back order, Collisions content (Q18)."""
    s, cfg = scenes.stack_scene()
    d, _ = scenes.domino_scene()
    c1, _ = scenes.c1_stack20()
    cfg = Configuration()
    _run([s, d, c1], cfg, [1, 4, 45, 150, 300], "scenes")


def test_t7_exact_serial_path_bit_exact_including_static_L():
    s, _ = scenes.stack_scene()
    c1, _ = scenes.c1_stack20()
    _run([s, c1], Configuration(), [1, 9, 90], "serial", exact=True)


def test_c2_c3_slices():
    """BASELINE configs 2 and 3 on a slice of worlds, bit-exact vs the oracle."""
    w2 = [scenes.c2_spheres64(w)[0] for w in range(6)]
    _run(w2, scenes.c2_spheres64(0)[1], [1, 19, 180], "C2 slice")
    w3 = [scenes.c3_mixed128(w)[0] for w in range(3)]
    _run(w3, scenes.c3_mixed128(0)[1], [1, 19, 100], "C3 slice")


def test_t8_joints_chain():
    """T8: C4 chains with the 'restore' as written (Q14): Jacobi over joints, last writer wins."""
    worlds, joints = [], []
    for w in range(3):
        p, cfg, j = scenes.c4_chain32_blowup(w, n_links=8)
        worlds.append(p)
        joints.append(j)
    # with light links the "restore" as written is an unstable feedback (Q14): these chains blow up to NaN within
    # ~35 steps in the oracle too; parity is checked before and through the blow-up (NaN == NaN)
    _run(worlds, cfg, [1, 5, 14, 30], "C4 chains (light links, blow-up)", joints=joints)


def test_t8_joints_chain_32_links_full_size_worlds():
    """the BASELINE configs[3] world as benched: 32 links of 1000 kg (finite states), joints + box-box contacts with
    friction while the chain falls, lands and piles up"""
    worlds, joints = [], []
    for w in range(4):
        p, cfg, j = scenes.c4_chain32(w)
        worlds.append(p)
        joints.append(j)
    _run(worlds, cfg, [1, 9, 60, 50], "C4 chains, 32 links", joints=joints)


def test_mutations_apply_impulse_add_object_reset():
    """Simulator.ApplyImpulse / AddObject / Reset (Simulator.fs:91-121) through the facade."""
    from physicssimulator_b200.simulator import Simulator
    from oracle.oracle import OracleWorld
    protos, cfg = scenes.stack_scene()
    cfg = Configuration()  # Dummy order
    sim = Simulator(protos, cfg, device=0)
    o = OracleWorld(P.build_world(protos), to_native_config(cfg))
    sim.Step(20); o.step(20)
    sim.ApplyImpulse(4, (0.0, 200.0, 0.0), (0.0, 0.0, 0.0)); o.apply_impulse(4, (0.0, 200.0, 0.0), (0.0, 0.0, 0.0))
    sim.Step(30); o.step(30)
    new = P.createDefaultCube(0.5).With(Mass=50.0, Position=(0.0, -3.0, 3.0))
    sim.AddObject(new)
    one = P.empty_bodies(1); P.build_into(one, 0, new)
    assert o.add_body(one) == 5
    sim.Step(60); o.step(60)
    got, ref = sim._b.WorldState(0), o.get_state()
    dyn = np.array([False, True, True, True, True, True])
    for k in ("pos", "vel", "rot"):
        assert same_bits(got[k], ref[k])
    assert same_bits(got["ang_mom"][dyn], ref["ang_mom"][dyn])
    assert sim.ObjectsIdentifiers == set(range(6))
    sim.Reset()
    assert sim.ObjectsIdentifiers == set(range(5))
    o2 = OracleWorld(P.build_world(protos), to_native_config(cfg))
    sim.Step(10); o2.step(10)
    got, ref = sim._b.WorldState(0), o2.get_state()
    assert same_bits(got["pos"], ref["pos"])
    with pytest.raises(KeyError):
        sim.ApplyImpulse(77, (0, 0, 1), (0, 0, 0))
    sim.Dispose()


def test_wavefront_equals_serial_on_device():
    """The conservative cull + wavefront path and the exact all-pairs serial walk give the same bits."""
    from physicssimulator_b200.simulator import BatchSimulator
    worlds = [scenes.c3_mixed128(w, n_boxes=24, n_spheres=24)[0] for w in range(8)]
    cfg = Configuration()
    a = BatchSimulator(worlds, cfg, device=0)
    b = BatchSimulator(worlds, cfg, device=0, exact_all_pairs=True)
    for n in (1, 30, 120):
        a.Step(n); b.Step(n)
        sa, sb = a.State(), b.State()
        for k in ("pos", "vel", "rot"):
            assert same_bits(sa[k], sb[k])
    a.close(); b.close()


def test_errors_are_loud():
    from physicssimulator_b200 import _native
    from physicssimulator_b200.simulator import BatchSimulator
    protos, _ = scenes.stack_scene()
    sim = BatchSimulator([protos], Configuration(), device=0)
    with pytest.raises(KeyError):
        sim.batch.get_state(0, 5)
    with pytest.raises(RuntimeError):
        sim.Collisions(0)  # record_contacts was off
    bad = [p for p in protos]
    bad[1] = bad[1].With(ElasticityCoeff=1.5)
    with pytest.raises(ValueError):
        BatchSimulator([bad], Configuration(), device=0)
    sim.close()


def test_tile_widths_and_multi_step_launches_agree(monkeypatch):
    """the lanes-per-world choice (4/8/16/32) and n steps per launch vs n launches are pure scheduling: same bits"""
    from physicssimulator_b200.simulator import BatchSimulator
    worlds = [scenes.c3_mixed128(w, n_boxes=20, n_spheres=20)[0] for w in range(9)]  # 9: a partly filled last warp
    ref = None
    # both register builds of the step kernel (small batches default to the 255-register one)
    for tile, minb in (("4", "8"), ("8", "8"), ("16", "8"), ("32", "8"), ("4", "16"), ("8", "16"), ("16", "16"), ("32", "16")):
        monkeypatch.setenv("PS_B200_TILE", tile)
        monkeypatch.setenv("PS_B200_MINB", minb)
        sim = BatchSimulator(worlds, Configuration(), device=0)
        if tile == "8":
            for _ in range(60):
                sim.Step(1)
        else:
            sim.Step(60)
        st = sim.State()
        sim.close()
        if ref is None:
            ref = st
        else:
            for k in ("pos", "vel", "rot"):
                assert same_bits(st[k], ref[k]), f"tile {tile}, {minb} CTAs per SM: {k}"


@pytest.mark.parametrize("path", ["wide", "fused"])
def test_both_step_paths_against_the_oracle(monkeypatch, path):
    """the batch-wide (breadth-first) path and the fused per-world kernel, forced explicitly, on the scenes, the C2/C3
    slices and the box cases: same bits as the oracle (small batches default to the fused kernel, so the wide path
    would otherwise only be exercised by bench-size runs)"""
    monkeypatch.setenv("PS_B200_PATH", path)
    s, _ = scenes.stack_scene()
    d, _ = scenes.domino_scene()
    c1, _ = scenes.c1_stack20()
    _run([s, d, c1], Configuration(), [1, 4, 45, 150], f"scenes/{path}")
    w3 = [scenes.c3_mixed128(w)[0] for w in range(3)]
    _run(w3, scenes.c3_mixed128(0)[1], [1, 19, 100], f"C3/{path}")
    w2 = [scenes.c2_spheres64(w)[0] for w in range(4)]
    _run(w2, scenes.c2_spheres64(0)[1], [1, 19, 120], f"C2/{path}")


def test_wide_path_second_order_integrator_and_many_worlds(monkeypatch):
    monkeypatch.setenv("PS_B200_PATH", "wide")
    cfg = Configuration(StepConfig=StepConfiguration(RigidBodyIntegratorKind=1))
    worlds = [scenes.c3_mixed128(w, n_boxes=12, n_spheres=12)[0] for w in range(40)]
    _run(worlds, cfg, [1, 30, 90], "wide/second order", contacts=False)


def test_joints_in_chunks_of_the_tile_width(monkeypatch):
    """more joints per world than lanes per world: results are staged and committed after every joint has read"""
    monkeypatch.setenv("PS_B200_TILE", "4")
    worlds, joints = [], []
    for w in range(5):
        p, cfg, j = scenes.c4_chain32_blowup(w, n_links=11)
        worlds.append(p)
        joints.append(j)
    _run(worlds, cfg, [1, 5, 12], "C4 chains, 4 lanes per world", joints=joints)


@pytest.mark.parametrize("path,tile", [("fused", "4"), ("fused", "16"), ("fused", "32"), ("wide", "8")])
def test_tight_margins_exercise_the_verification_and_the_retry(monkeypatch, path, tile):
    """a motion margin far too small: bodies leave their cull spheres all the time, so the culled pairs are re-examined
    (phase V / wide_verify) and some world-steps restart with the wider margin — the result is still the oracle's"""
    from physicssimulator_b200.simulator import BatchSimulator
    monkeypatch.setenv("PS_B200_PATH", path)
    monkeypatch.setenv("PS_B200_TILE", tile)
    monkeypatch.setenv("PS_B200_MSCALE", "0.02")
    monkeypatch.setenv("PS_B200_NCAP", "0")
    worlds = [scenes.c3_mixed128(w, n_boxes=24, n_spheres=24)[0] for w in range(7)]
    cfg = scenes.c3_mixed128(0)[1]
    sim = BatchSimulator(worlds, cfg, device=0)
    oracles = _oracle_worlds(worlds, cfg)
    for n in (1, 30, 90):
        sim.Step(n)
        for o in oracles:
            o.step(n)
        _compare(sim, oracles, f"tight margins {path}/{tile} +{n}", False)
    st = sim.Stats()
    assert st["margin_verified_steps"] > 0, st
    sim.close()


# ---- BroadPhaseCollisionDetectionKind.SpatialTree: the candidate ORDER of the persistent tree ---------------------
def _tree_cfg(**kw):
    from physicssimulator_b200.configuration import BroadPhaseCollisionDetectionKind, SpatialTreeConfiguration
    return Configuration(BroadPhaseCollisionDetectionKind=BroadPhaseCollisionDetectionKind(SpatialTreeConfiguration(**kw)))


@pytest.mark.parametrize("tile", ["", "4", "16"])
def test_spatial_tree_mode_gui_scenes(monkeypatch, tile):
    """the Gui scenes in the configuration they ship with (Gui/Scenes/StackScene.fs:36-42, DominoScene.fs:45-55): states,
    colliding sets, contacts, feature ids AND the ordered candidate list of every step chunk equal the oracle's"""
    if tile:
        monkeypatch.setenv("PS_B200_TILE", tile)
    s, cfg_s = scenes.stack_scene()
    _run([s], cfg_s, [1, 4, 45, 150], f"stack/tree/{tile}")
    d, cfg_d = scenes.domino_scene()
    _run([d], cfg_d, [1, 4, 45, 100], f"domino/tree/{tile}")


def test_spatial_tree_mode_batch_of_worlds_and_exact_walk():
    """one tree per world; several worlds per launch; the exact serial walk follows the tree's list too"""
    cfg = _tree_cfg(LeafCapacity=2, MaxDepth=10)
    c1, _ = scenes.c1_stack20()
    s, _ = scenes.stack_scene()
    _run([c1, s, c1], cfg, [1, 9, 60, 130], "c1+stack/tree")
    _run([s, c1], cfg, [1, 9, 60], "tree/serial", exact=True)
    cfg3 = scenes.c3_mixed128(0)[1].With(BroadPhaseCollisionDetectionKind=_tree_cfg(
        LeafCapacity=3, MaxDepth=6, SpaceBoundariesMin=(-30.0, -30.0, -3.0), SpaceBoundariesMax=(30.0, 30.0, 30.0)).BroadPhaseCollisionDetectionKind)
    w3 = [scenes.c3_mixed128(w, n_boxes=14, n_spheres=14)[0] for w in range(5)]
    _run(w3, cfg3, [1, 19, 60], "c3 slices/tree")


def test_spatial_tree_mode_mutations_and_errors():
    """AddObject joins the existing tree (BroadPhase.onNewObjectAdded), Reset rebuilds it; a body outside the space at
    construction is an ArgumentException in the reference (SpatialTree.fs:171-172)"""
    from physicssimulator_b200.simulator import Simulator, BatchSimulator
    from oracle.oracle import OracleWorld
    protos, cfg = scenes.stack_scene()
    tc = cfg.BroadPhaseCollisionDetectionKind.SpatialTree
    sim = Simulator(protos, cfg, device=0)
    o = OracleWorld(P.build_world(protos), to_native_config(cfg))
    assert o.use_spatial_tree(tc.LeafCapacity, tc.MaxDepth, tc.SpaceBoundariesMin, tc.SpaceBoundariesMax) == 0
    sim.Step(25); o.step(25)
    new = P.createDefaultCube(0.5).With(Mass=50.0, Position=(0.0, -3.0, 3.0))
    sim.AddObject(new)
    one = P.empty_bodies(1); P.build_into(one, 0, new)
    assert o.add_body(one) == 5
    sim.Step(70); o.step(70)
    got, ref = sim._b.WorldState(0), o.get_state()
    for k in ("pos", "vel", "rot"):
        assert same_bits(got[k], ref[k]), k
    sim.Reset()
    o2 = OracleWorld(P.build_world(protos), to_native_config(cfg))
    assert o2.use_spatial_tree(tc.LeafCapacity, tc.MaxDepth, tc.SpaceBoundariesMin, tc.SpaceBoundariesMax) == 0
    sim.Step(40); o2.step(40)
    assert same_bits(sim._b.WorldState(0)["pos"], o2.get_state()["pos"])
    sim.Dispose()
    far = protos + [P.createDefaultCube(1.0).With(Position=(100.0, 0.0, 0.0))]
    with pytest.raises(ValueError):
        BatchSimulator([far], cfg, device=0)


@pytest.mark.gpu
@pytest.mark.parametrize("chunks", ["1", "3", "4", "64"])
def test_step_host_equals_set_step_get(monkeypatch, chunks):
    """ps_batch_step_host (copies pipelined against the step over chunks of worlds) == the three calls in sequence"""
    monkeypatch.setenv("PS_B200_HOST_CHUNKS", chunks)
    worlds = [scenes.c2_spheres64(w)[0] for w in range(37)]  # not a multiple of anything
    cfg = scenes.c2_spheres64(0)[1]
    a = BatchSimulator(worlds, cfg, device=0)
    b = BatchSimulator(worlds, cfg, device=0)
    try:
        a.Step(30); b.Step(30)
        st = a.State()
        for _ in range(3):
            a.batch.set_state(st); a.batch.step(2); ref = a.batch.get_state()
            got = b.StepHost(st, 2)
            assert_state_bits(ref, got, f"step_host chunks={chunks}")
            st = {k: v.copy() for k, v in got.items()}
        sa, sb = a.Stats(), b.Stats()
        assert sa["pairs_colliding"] == sb["pairs_colliding"] and sa["contacts"] == sb["contacts"] and sa["steps"] == sb["steps"]
        inplace = {k: v.copy() for k, v in st.items()}
        b.StepHost(inplace, 1, out=inplace)   # in == out
        a.batch.set_state(st); a.batch.step(1)
        assert_state_bits(a.batch.get_state(), inplace, "step_host in place")
    finally:
        a.close(); b.close()


@pytest.mark.gpu
def test_step_host_on_the_spatial_tree_path_and_errors():
    protos, cfg = scenes.stack_scene()
    a = BatchSimulator([protos, protos], cfg, device=0)
    b = BatchSimulator([protos, protos], cfg, device=0)
    try:
        st = a.State()
        a.batch.set_state(st); a.batch.step(5)
        assert_state_bits(a.batch.get_state(), b.StepHost(st, 5), "step_host, tree mode")
        with pytest.raises(ValueError):  # PS_ERR_INVALID_ARGUMENT -> ArgumentException class
            b.batch.step_host(st, -1)
    finally:
        a.close(); b.close()


@pytest.mark.gpu
def test_world_reordering_changes_nothing(monkeypatch):
    """worlds are re-ordered inside their segments by the passes of their last step (every 16 steps): which warp a
    world runs in must not show in any bit — single-step calls, multi-step launches and ps_batch_step_host"""
    worlds = [scenes.c2_spheres64(w)[0] for w in range(70)] + [scenes.c3_mixed128(w, n_boxes=6, n_spheres=6)[0] for w in range(9)]
    cfg = Configuration()
    ref = BatchSimulator(worlds, cfg, device=0)
    a = BatchSimulator(worlds, cfg, device=0)
    b = BatchSimulator(worlds, cfg, device=0)
    try:
        monkeypatch.setenv("PS_B200_PERM", "0")
        ref.Step(75)
        monkeypatch.setenv("PS_B200_PERM", "1")
        for _ in range(40):
            a.Step(1)
        a.Step(35)
        assert_state_bits(ref.State(), a.State(), "re-ordered worlds, step")
        st = b.State()
        for n in (1, 20, 17, 37):
            st = b.StepHost(st, n)
        assert_state_bits(ref.State(), st, "re-ordered worlds, step_host")
        ra, rb_, rr = a.Stats(), b.Stats(), ref.Stats()
        for k in ("pairs_colliding", "contacts", "steps", "body_steps"):
            assert ra[k] == rr[k] == rb_[k], k
    finally:
        ref.close(); a.close(); b.close()
