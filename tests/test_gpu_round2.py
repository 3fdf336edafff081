"""GPU tests added in round 2: parity at the size where each path is the DEFAULT one (full C2 batch, >= 4096 C3 worlds on
the batch-wide path), the same worlds on one handle vs split over two (T10), stream ordering of
ps_batch_step_async, the limits ps_batch_add_body must check before it changes anything, conservative SpatialTree
extents (aabb_mode), input validation."""
import os

import numpy as np
import pytest

from common import (P, scenes, Configuration, to_native_config, same_bits, assert_state_bits)

from physicssimulator_b200.simulator import BatchSimulator, Simulator

pytestmark = pytest.mark.gpu


def _built(fn, w0, n, **kw):
    out, cfg = [], None
    for w in range(w0, w0 + n):
        r = fn(w, **kw)
        cfg = r[1]
        out.append(P.build_world(r[0], cfg.StepConfig.GravitationalAccelerationMagnitude, cfg.StepConfig.GravityDirection))
    return out, cfg


def _check_sample_against_oracle(sim, built, cfg, n_steps, sample, what):
    from oracle.oracle import OracleWorld, batch_step
    oracles = [OracleWorld(built[w], to_native_config(cfg)) for w in sample]
    batch_step(oracles, n_steps, min(len(oracles), os.cpu_count() or 1))
    for w, o in zip(sample, oracles):
        got, ref = sim.batch.get_state(w, w + 1), o.get_state()
        dyn = ~np.isinf(built[w]["mass"])
        for k in ("pos", "vel", "rot", "ang_mom"):
            assert same_bits(got[k][dyn], ref[k][dyn]), f"{what}: world {w} field {k} differs from the oracle after {n_steps} steps"


def test_full_c2_batch_default_path_against_the_oracle():
    """BASELINE configs[1] at full size (4096 worlds x 64 spheres), the path the library picks by itself, from creation;
    16 worlds spread over the batch are held against the oracle bit for bit"""
    built, cfg = _built(scenes.c2_spheres64, 0, 4096)
    sim = BatchSimulator(None, cfg, device=0, _prebuilt=built)
    n = 0
    for chunk in (1, 39, 60):  # free fall, first landings, sphere-sphere contacts
        sim.Step(chunk)
        n += chunk
    sample = sorted({(q * 4096) // 16 + (q * 37) % 200 for q in range(16)})
    _check_sample_against_oracle(sim, built, cfg, n, sample, "full C2")
    st = sim.Stats()
    assert st["nan_worlds"] == 0 and st["contact_overflow"] == 0 and st["pairs_colliding"] > 4096
    sim.close()


def test_4096_c3_worlds_wide_path_by_default_against_the_oracle():
    """4096 worlds x 128 mixed bodies = 524 288 bodies: above the size where the batch-wide path is the default (no
    environment override here), many-CTA wide_levels / wide_scatter, 4096-world bucket tables; from creation through the
    landing (box-box manifolds, friction); 16 worlds against the oracle bit for bit"""
    assert "PS_B200_PATH" not in os.environ
    built, cfg = _built(scenes.c3_mixed128, 0, 4096)
    sim = BatchSimulator(None, cfg, device=0, _prebuilt=built)
    assert sim.batch.schedule()["path"] == "wide"
    n = 0
    for chunk in (1, 29, 40):
        sim.Step(chunk)
        n += chunk
    sample = sorted({(q * 4096) // 16 + (q * 53) % 250 for q in range(16)})
    _check_sample_against_oracle(sim, built, cfg, n, sample, "4096 C3 worlds (wide)")
    st = sim.Stats()
    assert st["nan_worlds"] == 0 and st["contact_overflow"] == 0 and st["pairs_colliding"] > 4096
    sim.close()


def test_t10_same_worlds_on_one_handle_and_split_over_two():
    """T10: 64 worlds stepped by one handle, and the same worlds as two shards of 32 on two handles (two
    devices when the box has them): every body identical bit for bit, and the shard statistics sum to the whole"""
    import torch
    from physicssimulator_b200.sharding import shard_range, SUM_KEYS
    built, cfg = _built(scenes.c3_mixed128, 100, 64, n_boxes=20, n_spheres=20)
    one = BatchSimulator(None, cfg, device=0, _prebuilt=built)
    ndev = torch.cuda.device_count()
    shards = []
    for r in range(2):
        w0, w1 = shard_range(64, r, 2)
        shards.append((w0, w1, BatchSimulator(None, cfg, device=(r if ndev >= 2 else 0), _prebuilt=built[w0:w1])))
    for chunk in (1, 49, 70):
        one.Step(chunk)
        for _, _, s in shards:
            s.Step(chunk)
    whole = one.State()
    off = one._offsets
    for w0, w1, s in shards:
        part = s.State()
        for k in ("pos", "vel", "rot", "ang_mom"):
            assert same_bits(part[k], whole[k][off[w0]:off[w1]]), f"T10: worlds [{w0},{w1}) field {k}"
    tot, parts = one.Stats(), [s.Stats() for _, _, s in shards]
    for k in SUM_KEYS:
        assert tot[k] == sum(p[k] for p in parts), k
    assert tot["max_penetration"] == max(p["max_penetration"] for p in parts)
    one.close()
    for _, _, s in shards:
        s.close()


@pytest.mark.parametrize("path", ["fused", "wide"])
def test_step_async_on_a_foreign_stream_is_ordered_with_the_handle(monkeypatch, path):
    """ps_batch_step_async launches on the caller's stream; get_state / apply_impulse / step on the handle's own stream
    right after it (no synchronise in between) must see its result, and a later async step must see theirs"""
    import torch
    from oracle.oracle import OracleWorld
    monkeypatch.setenv("PS_B200_PATH", path)
    built, cfg = _built(scenes.c3_mixed128, 7, 6, n_boxes=16, n_spheres=16)
    sim = BatchSimulator(None, cfg, device=0, _prebuilt=built)
    oracles = [OracleWorld(b, to_native_config(cfg)) for b in built]
    stream = torch.cuda.Stream()
    J, offv = (3.0, -2.0, 5.0), (0.1, 0.0, -0.05)
    for rnd in range(3):
        sim.batch.step_async(25, stream.cuda_stream)
        st = sim.batch.get_state()                      # handle's stream, no synchronise before
        for o in oracles:
            o.step(25)
        for w, o in enumerate(oracles):
            ref = o.get_state()
            dyn = ~np.isinf(built[w]["mass"])
            a, b = sim._offsets[w], sim._offsets[w + 1]
            for k in ("pos", "vel", "rot"):
                assert same_bits(st[k][a:b][dyn], ref[k][dyn]), f"async round {rnd} world {w} {k}"
        sim.batch.step_async(3, stream.cuda_stream)
        sim.ApplyImpulse(2, 5, J, offv)                 # handle's stream: must come after the 3 async steps
        for o in oracles:
            o.step(3)
        oracles[2].apply_impulse(5, J, offv)
    sim.batch.step_async(4, stream.cuda_stream)         # must come after the impulse
    for o in oracles:
        o.step(4)
    sim.batch.synchronize()
    st = sim.batch.get_state()
    for w, o in enumerate(oracles):
        ref = o.get_state()
        dyn = ~np.isinf(built[w]["mass"])
        a, b = sim._offsets[w], sim._offsets[w + 1]
        for k in ("pos", "vel", "rot"):
            assert same_bits(st[k][a:b][dyn], ref[k][dyn]), f"async tail world {w} {k}"
    sim.close()


def test_step_async_under_graph_capture():
    """the fused path has no host round trip: a captured step replays; a SpatialTree-order batch refuses to be captured"""
    import torch
    built, cfg = _built(scenes.c2_spheres64, 0, 8)
    sim = BatchSimulator(None, cfg, device=0, _prebuilt=built)
    ref = BatchSimulator(None, cfg, device=0, _prebuilt=built)
    sim.Step(30); ref.Step(30)
    s = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    torch.cuda.synchronize()
    with torch.cuda.stream(s):
        g.capture_begin()
        sim.batch.step_async(2, s.cuda_stream)
        g.capture_end()
    torch.cuda.synchronize()
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    ref.Step(10)
    assert_state_bits(ref.State(), sim.State(), "graph replay of the fused step")
    sim.close(); ref.close()
    # the batch-wide path: its rounds are counted on the device (no host round trip), so a step can be captured too
    import os as _os
    _os.environ["PS_B200_PATH"] = "wide"
    try:
        built3, cfg3 = _built(scenes.c3_mixed128, 0, 12, n_boxes=12, n_spheres=12)
        wsim = BatchSimulator(None, cfg3, device=0, _prebuilt=built3)
        wref = BatchSimulator(None, cfg3, device=0, _prebuilt=built3)
        assert wsim.batch.schedule()["path"] == "wide"
        wsim.Step(40); wref.Step(40)
        gw = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        with torch.cuda.stream(s):
            gw.capture_begin()
            wsim.batch.step_async(1, s.cuda_stream)
            gw.capture_end()
        torch.cuda.synchronize()
        for _ in range(6):
            gw.replay()
        torch.cuda.synchronize()
        wref.Step(6)
        assert_state_bits(wref.State(), wsim.State(), "graph replay of the batch-wide step")
        wsim.close(); wref.close()
    finally:
        del _os.environ["PS_B200_PATH"]
    protos, tcfg = scenes.stack_scene()
    tree = BatchSimulator([protos], tcfg, device=0)
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        g2.capture_begin()
        with pytest.raises(RuntimeError):
            tree.batch.step_async(1, s.cuda_stream)
        g2.capture_end()
    tree.Step(3)
    tree.close()


def test_add_body_refuses_before_it_changes_anything():
    """per-world size limit and (SpatialTree order) the tree's space are checked first: after the refusal the batch steps
    on exactly as if the call had never been made (the handle used to be left half-built)"""
    from oracle.oracle import OracleWorld
    # 1) 1024 bodies is the limit of a world
    protos = [P.createDefaultBox(400, 400, 1).With(Mass=P.INFINITE, UseGravity=False, Position=(0, 0, -0.5), Yaw=np.pi)]
    for k in range(1023):
        protos.append(P.createDefaultSphere(0.3).With(Position=(((k % 32) - 16) * 3.0, ((k // 32) - 16) * 3.0, 0.5), Mass=2.0))
    small, _ = scenes.c2_spheres64(1, n_spheres=8)
    sim = BatchSimulator([small, protos], Configuration(), device=0)
    sim.Step(3)
    before = sim.State()
    with pytest.raises(ValueError):
        sim.AddObject(1, P.createDefaultSphere(0.2).With(Position=(0, 0, 9.0)))
    assert_state_bits(before, sim.State(), "state after a refused add_body")
    sim.Step(2)
    ok_id = sim.AddObject(0, P.createDefaultSphere(0.2).With(Position=(0, 0, 9.0)))  # the other world still takes one
    assert ok_id == 9
    sim.Step(2)
    assert sim.Stats()["nan_worlds"] == 0
    sim.close()
    # 2) SpatialTree order: a body outside the tree's space
    sp, cfg = scenes.stack_scene()
    tc = cfg.BroadPhaseCollisionDetectionKind.SpatialTree
    s2 = Simulator(sp, cfg, device=0)
    o = OracleWorld(P.build_world(sp), to_native_config(cfg))
    assert o.use_spatial_tree(tc.LeafCapacity, tc.MaxDepth, tc.SpaceBoundariesMin, tc.SpaceBoundariesMax) == 0
    s2.Step(10); o.step(10)
    with pytest.raises(ValueError):
        s2.AddObject(P.createDefaultCube(0.5).With(Position=(500.0, 0.0, 3.0)))
    s2.Step(30); o.step(30)
    got, ref = s2._b.WorldState(0), o.get_state()
    assert len(got["pos"]) == len(ref["pos"]) == len(sp)
    for k in ("pos", "vel", "rot"):
        assert same_bits(got[k], ref[k]), k
    s2.Dispose()


def test_spatial_tree_conservative_extents():
    """ps_step_config.aabb_mode = 1: spheres enter the tree with side 2r instead of the reference's r (Q5); candidate
    lists, contacts and states equal the oracle's tree run in the same mode"""
    from oracle.oracle import OracleWorld
    from physicssimulator_b200.configuration import BroadPhaseCollisionDetectionKind, SpatialTreeConfiguration
    cfg = Configuration(BroadPhaseCollisionDetectionKind=BroadPhaseCollisionDetectionKind(SpatialTreeConfiguration(
        LeafCapacity=3, MaxDepth=6, SpaceBoundariesMin=(-30.0, -30.0, -3.0), SpaceBoundariesMax=(30.0, 30.0, 30.0))))
    tc = cfg.BroadPhaseCollisionDetectionKind.SpatialTree
    worlds = [scenes.c3_mixed128(w, n_boxes=10, n_spheres=18)[0] for w in range(3)]
    lists_differ = False
    for mode in (0, 1):
        sim = BatchSimulator(worlds, cfg, device=0, record_contacts=True, aabb_mode=mode)
        oracles = []
        for protos in worlds:
            o = OracleWorld(P.build_world(protos), to_native_config(cfg))
            assert o.use_spatial_tree(tc.LeafCapacity, tc.MaxDepth, tc.SpaceBoundariesMin, tc.SpaceBoundariesMax, mode) == 0
            oracles.append(o)
        for n in (1, 39, 40):
            sim.Step(n)
            for w, o in enumerate(oracles):
                o.step(n)
                assert np.array_equal(o.candidates(), sim.Candidates(w)), f"aabb_mode {mode}: candidate list of world {w}"
                got, ref = sim.WorldState(w), o.get_state()
                for k in ("pos", "vel", "rot"):
                    assert same_bits(got[k], ref[k]), f"aabb_mode {mode} world {w} {k}"
        if mode == 0:
            first = [sim.Candidates(w).copy() for w in range(3)]
        else:
            lists_differ = any(len(first[w]) != len(sim.Candidates(w)) or not np.array_equal(first[w], sim.Candidates(w)) for w in range(3))
        sim.close()
    assert lists_differ, "the conservative extents should give longer candidate lists than the reference's r-sided boxes"


def test_mass_and_static_inertia_are_validated():
    protos, _ = scenes.stack_scene()
    b = P.build_world(protos)
    off = np.array([0, len(b["mass"])], dtype=np.int32)
    from physicssimulator_b200 import _native
    for field, idx, val in (("mass", 1, 0.0), ("mass", 1, -3.0), ("mass", 1, float("nan"))):
        bad = {k: v.copy() for k, v in b.items()}
        bad[field][idx] = val
        with pytest.raises(ValueError):
            _native.Batch(bad, off, to_native_config(Configuration()), device=0)
    stat = int(np.flatnonzero(np.isinf(b["mass"]))[0])
    bad = {k: v.copy() for k, v in b.items()}
    bad["inertia_inv"][stat] = (0.0, 0.5, 0.0)
    with pytest.raises(ValueError):
        _native.Batch(bad, off, to_native_config(Configuration()), device=0)


def _add_to_both(sim, oracles, world, proto):
    one = P.empty_bodies(1)
    P.build_into(one, 0, proto)
    new_id = sim.AddObject(world, proto)
    assert oracles[world].add_body(one) == new_id
    return new_id


def test_add_body_is_incremental_and_exact():
    """Simulator.AddObject (SimulatorState.fs:89-107) on a large batch: a world with a spare slot takes the body as a
    record write (100 bodies into a 4096-world batch in a few milliseconds, not a re-upload each), a full world triggers
    ONE re-layout; either way the batch steps on exactly like the oracle worlds that got the same bodies"""
    import time
    from oracle.oracle import OracleWorld
    built, cfg = _built(scenes.c2_spheres64, 0, 4096)
    sim = BatchSimulator(None, cfg, device=0, _prebuilt=built)
    sim.Step(20)
    touched = [17 * q % 4096 for q in range(100)]          # 100 different worlds
    oracles = {w: OracleWorld(built[w], to_native_config(cfg)) for w in set(touched) | {5}}
    for o in oracles.values():
        o.step(20)
    rng = scenes.SplitMix64(99)
    protos = [P.createDefaultSphere(rng.uniform(0.2, 0.4)).With(Position=(rng.uniform(-2, 2), rng.uniform(-2, 2), 6.0 + rng.uniform(0, 1)), Mass=rng.uniform(1, 5))
              for _ in touched]
    t0 = time.perf_counter()
    for w, p in zip(touched, protos):
        sim.AddObject(w, p)
    dt = time.perf_counter() - t0
    for w, p in zip(touched, protos):
        one = P.empty_bodies(1); P.build_into(one, 0, p)
        assert oracles[w].add_body(one) == 65
    assert dt < 0.25, f"100 AddObject calls took {dt * 1e3:.1f} ms"   # (a re-upload of this batch per call is ~100x that)
    # fill world 5 beyond its spare slots: the fifth addition lays the batch out again
    for k in range(7):
        _add_to_both(sim, oracles, 5, P.createDefaultCube(0.4).With(Position=(0.3 * k - 1.0, 0.5, 7.0 + 0.6 * k), Mass=3.0))
    sim.Step(60)
    st = sim.Stats()
    assert st["nan_worlds"] == 0
    for w, o in oracles.items():
        o.step(60)
        got, ref = sim.WorldState(w), o.get_state()
        assert len(got["pos"]) == len(ref["pos"])
        for k in ("pos", "vel", "rot"):
            assert same_bits(got[k], ref[k]), f"world {w} {k} after AddObject"
    assert sim.batch.num_bodies()[0] == 4096 * 65 + 107
    sim.close()
    print(f"100 AddObject calls on 4096 worlds: {dt * 1e3:.2f} ms")


def test_add_body_without_spare_slots_and_on_the_wide_path(monkeypatch):
    """PS_B200_SPARE=0: every addition is a re-layout (the general route); the batch-wide path sees added bodies too"""
    from oracle.oracle import OracleWorld
    monkeypatch.setenv("PS_B200_SPARE", "0")
    monkeypatch.setenv("PS_B200_PATH", "wide")
    worlds = [scenes.c3_mixed128(w, n_boxes=6, n_spheres=6)[0] for w in range(9)]   # 13 bodies: N / 16 = 0 spare slots
    cfg = scenes.c3_mixed128(0)[1]
    sim = BatchSimulator(worlds, cfg, device=0)
    oracles = [OracleWorld(P.build_world(p), to_native_config(cfg)) for p in worlds]
    sim.Step(15)
    for o in oracles:
        o.step(15)
    _add_to_both(sim, oracles, 3, P.createDefaultCube(0.5).With(Position=(0.0, 0.0, 5.0), Mass=20.0))
    _add_to_both(sim, oracles, 3, P.createDefaultSphere(0.3).With(Position=(0.2, 0.1, 6.5), Mass=2.0))
    _add_to_both(sim, oracles, 8, P.createDefaultBox(3.0, 3.0, 0.2).With(Position=(0.0, 0.0, 3.0), Mass=P.INFINITE, UseGravity=False))  # a second static body
    monkeypatch.setenv("PS_B200_SPARE", "4")
    _add_to_both(sim, oracles, 8, P.createDefaultSphere(0.3).With(Position=(0.5, 0.5, 8.0), Mass=2.0))  # re-layout, now with spare slots
    _add_to_both(sim, oracles, 8, P.createDefaultSphere(0.25).With(Position=(-0.5, 0.5, 9.0), Mass=2.0))  # record write
    sim.Step(80)
    for w, o in enumerate(oracles):
        o.step(80)
        got, ref = sim.WorldState(w), o.get_state()
        dyn = ~np.isinf(np.concatenate([sim._worlds[w]["mass"]]))
        for k in ("pos", "vel", "rot"):
            assert same_bits(got[k][dyn], ref[k][dyn]), f"world {w} {k}"
    sim.close()


def test_joints_given_as_hi_lo_swap_the_bodies_like_the_reference():
    """Q15 (SimulatorState.fs:153-160): changePhysicalObjects zips the id-SORTED objects with the joint-ORDERED results, so a
    joint given as (a > b) hands each body the other's restored state; reproduced, not rejected.  A self joint throws."""
    from oracle.oracle import OracleWorld
    worlds, joints = [], []
    for w in range(4):
        p, cfg, j = scenes.c4_chain32(w, n_links=9)
        j = [(b, a, anchor) if k % 3 == 1 else (a, b, anchor) for k, (a, b, anchor) in enumerate(j)]  # every third joint reversed
        worlds.append(p)
        joints.append(j)
    sim = BatchSimulator(worlds, cfg, joints=joints, device=0)
    oracles = []
    for p, js in zip(worlds, joints):
        b = P.build_world(p)
        oracles.append(OracleWorld(b, to_native_config(cfg), joints=[(a, bb) + P.ball_socket_local_anchors(b, a, bb, anchor) for (a, bb, anchor) in js]))
    for n in (1, 7, 40):
        sim.Step(n)
        for w, o in enumerate(oracles):
            o.step(n)
            got, ref = sim.WorldState(w), o.get_state()
            for k in ("pos", "vel", "rot"):
                assert same_bits(got[k], ref[k]), f"reversed joints: world {w} {k} after +{n}"
    sim.close()
    p, cfg, j = scenes.c4_chain32(0, n_links=4)
    with pytest.raises(ValueError):
        BatchSimulator([p], cfg, joints=[[(2, 2, j[0][2])]], device=0)


def test_wide_path_on_ragged_worlds(monkeypatch):
    """the batch-wide path on worlds of very different sizes in one batch: empty, one body, a few, more than one CTA's rows
    (150) and more than the bit-mask rows of the near-list kernel hold (300: the two-walk route) — against the oracle"""
    from oracle.oracle import OracleWorld
    monkeypatch.setenv("PS_B200_PATH", "wide")
    rng = scenes.SplitMix64(4242)

    def pile(n, boxes):
        protos = [P.createDefaultBox(60, 60, 1).With(Mass=P.INFINITE, UseGravity=False, Position=(0, 0, -0.5), Yaw=np.pi)]
        side = int(np.ceil(np.sqrt(max(n, 1))))
        for k in range(n):
            pos = ((k % side - side / 2) * 0.9 + rng.uniform(-0.05, 0.05), (k // side - side / 2) * 0.9 + rng.uniform(-0.05, 0.05), 0.6 + 0.5 * (k % 3))
            if boxes and k % 4 == 0:
                protos.append(P.createDefaultBox(0.5, 0.4, 0.45).With(Position=pos, Mass=8.0, Yaw=rng.uniform(-1, 1)))
            else:
                protos.append(P.createDefaultSphere(0.3 + 0.1 * (k % 2)).With(Position=pos, Mass=3.0))
        return protos

    worlds = [[], pile(0, False), pile(1, False), pile(7, True), pile(149, True), pile(299, False), pile(40, True)]
    cfg = Configuration()
    sim = BatchSimulator(worlds, cfg, device=0, record_contacts=True)
    assert sim.batch.schedule()["path"] == "wide"
    oracles = [OracleWorld(P.build_world(p), to_native_config(cfg)) if p else None for p in worlds]
    for n in (1, 14, 25):
        sim.Step(n)
        for w, o in enumerate(oracles):
            if o is None:
                assert len(sim.WorldState(w)["pos"]) == 0
                continue
            o.step(n)
            got, ref = sim.WorldState(w), o.get_state()
            dyn = ~np.isinf(P.build_world(worlds[w])["mass"])
            for k in ("pos", "vel", "rot"):
                assert same_bits(got[k][dyn], ref[k][dyn]), f"ragged wide batch: world {w} ({len(worlds[w])} bodies) {k} after +{n}"
            c_gpu, c_ref = sim.Collisions(w), o.contacts()
            assert np.array_equal(c_gpu["pairs"], c_ref["pairs"]) and np.array_equal(c_gpu["feature"], c_ref["feature"]), f"world {w}: contact sets"
    st = sim.Stats()
    assert st["nan_worlds"] == 0 and st["contact_overflow"] == 0
    sim.close()


@pytest.mark.parametrize("path", ["fused", "wide"])
@pytest.mark.parametrize("integrator,friction", [(0, True), (1, True), (0, False)])
def test_random_worlds_both_paths(monkeypatch, path, integrator, friction):
    """120 small random worlds in one batch — boxes and spheres at random poses and velocities, axis-aligned and rotated,
    interpenetrating at creation, a second static body in some — both step paths, both integrators, friction on and off:
    states, colliding sets and feature ids against the oracle"""
    from common import StepConfiguration
    from oracle.oracle import OracleWorld
    monkeypatch.setenv("PS_B200_PATH", path)
    rng = scenes.SplitMix64(1000 + 10 * integrator + (1 if friction else 0))
    worlds = []
    for w in range(120):
        protos = [P.createDefaultBox(8, 8, 1).With(Mass=P.INFINITE, UseGravity=False, Position=(0, 0, -0.5), Yaw=(np.pi if w % 2 else 0.0))]
        if w % 7 == 3:
            protos.append(P.createDefaultBox(1.0, 3.0, 0.6).With(Mass=P.INFINITE, UseGravity=False, Position=(1.5, 0.0, 0.3), Yaw=0.4))
        aligned = w % 3 == 0
        for k in range(3 + w % 6):
            ang = (lambda: 0.0) if aligned else (lambda: rng.uniform(-3.1, 3.1))
            pos = (rng.uniform(-1.2, 1.2), rng.uniform(-1.2, 1.2), rng.uniform(0.2, 2.2))
            if aligned:
                pos = tuple(round(x * 4) / 4 for x in pos)
            vel = (rng.uniform(-2, 2), rng.uniform(-2, 2), rng.uniform(-3, 1))
            if (k + w) % 3 == 0:
                protos.append(P.createDefaultSphere(rng.uniform(0.15, 0.5)).With(Position=pos, Velocity=vel, Mass=rng.uniform(0.5, 9)))
            else:
                protos.append(P.createDefaultBox(rng.uniform(0.2, 1.0), rng.uniform(0.2, 1.0), rng.uniform(0.2, 1.0)).With(
                    Position=pos, Velocity=vel, Yaw=ang(), Pitch=ang(), Roll=ang(), Mass=rng.uniform(0.5, 40),
                    ElasticityCoeff=rng.uniform(0.0, 0.9), KineticFrictionCoeff=rng.uniform(0.0, 1.0)))
        worlds.append(protos)
    cfg = Configuration(StepConfig=StepConfiguration(RigidBodyIntegratorKind=integrator, EnableFriction=friction))
    sim = BatchSimulator(worlds, cfg, device=0, record_contacts=True)
    assert sim.batch.schedule()["path"] == path
    oracles = [OracleWorld(P.build_world(p), to_native_config(cfg)) for p in worlds]
    for n in (1, 4, 25):
        sim.Step(n)
        st = sim.State()
        for w, o in enumerate(oracles):
            o.step(n)
            ref = o.get_state()
            a, b = sim._offsets[w], sim._offsets[w + 1]
            dyn = ~np.isinf(sim._worlds[w]["mass"])
            for k in ("pos", "vel", "rot", "ang_mom"):
                assert same_bits(st[k][a:b][dyn], ref[k][dyn]), f"random worlds {path}/{integrator}/{friction}: world {w} {k} after +{n}"
            c_gpu, c_ref = sim.Collisions(w), o.contacts()
            assert np.array_equal(c_gpu["pairs"], c_ref["pairs"]) and np.array_equal(c_gpu["feature"], c_ref["feature"]), f"world {w}: contacts after +{n}"
    stt = sim.Stats()
    assert stt["contact_overflow"] == 0 and stt["pairs_colliding"] > 500
    sim.close()


@pytest.mark.parametrize("chunks", ["1", "3", "4"])
def test_step_host_on_the_wide_path_in_chunks(monkeypatch, chunks):
    """ps_batch_step_host on the batch-wide path: chunks of worlds, each with its own chain of rounds on its own stream and
    its own slice of the pair / hit / contact lists, must give the bits of set_state + step + get_state on the whole batch —
    ragged worlds, a world that falls back to the fused kernel (more statics than the wide path takes), several steps per call,
    in == out, and the whole-batch step and the chunked one alternating on ONE handle."""
    monkeypatch.setenv("PS_B200_PATH", "wide")
    monkeypatch.setenv("PS_B200_WIDE_HOST_CHUNKS", chunks)
    worlds = []
    for w in range(131):  # not a multiple of 32
        if w % 17 == 5: worlds.append(scenes.c2_spheres64(w)[0])
        elif w % 29 == 3: worlds.append([])
        else: worlds.append(scenes.c3_mixed128(w, n_boxes=5 + w % 7, n_spheres=4 + w % 5)[0])
    many_statics = [P.createDefaultBox(4, 4, 1).With(Mass=P.INFINITE, UseGravity=False, Position=(5.0 * k, 0, -0.5)) for k in range(6)]
    many_statics += [P.createDefaultSphere(0.4).With(Position=(5.0 * k, 0, 0.6), Mass=2.0) for k in range(6)]
    worlds[40] = many_statics
    cfg = Configuration()
    a = BatchSimulator(worlds, cfg, device=0)
    b = BatchSimulator(worlds, cfg, device=0)
    try:
        assert b.batch.schedule()["path"] == "wide"
        a.Step(40); b.Step(40)
        st = a.State()
        for n in (1, 3, 2):
            a.batch.set_state(st); a.batch.step(n); ref = a.batch.get_state()
            got = b.StepHost(st, n)
            assert_state_bits(ref, got, f"wide step_host, {chunks} chunks, {n} steps")
            st = {k: v.copy() for k, v in got.items()}
            b.Step(1); a.Step(1)  # the whole-batch pass in between uses the same lists
            assert_state_bits(a.State(), b.State(), "whole-batch step after a chunked one")
            st = b.State()
        sa, sb = a.Stats(), b.Stats()
        for k in ("pairs_tested", "pairs_colliding", "contacts", "steps", "fallback_steps"):
            assert sa[k] == sb[k], (k, sa[k], sb[k])
        inplace = {k: v.copy() for k, v in st.items()}
        b.StepHost(inplace, 1, out=inplace)
        a.batch.set_state(st); a.batch.step(1)
        assert_state_bits(a.batch.get_state(), inplace, "wide step_host in place")
        # a new body (a dynamic one: a record write; a static one: the batch is laid out again and the chunks with it)
        for proto in (P.createDefaultSphere(0.3).With(Position=(0.2, 0.1, 4.0), Mass=2.0),
                      P.createDefaultBox(1, 1, 1).With(Mass=P.INFINITE, UseGravity=False, Position=(3.0, 3.0, 0.5))):
            for s_ in (a, b): s_.AddObject(9, proto)
            st = a.State()
            a.batch.set_state(st); a.batch.step(2); ref = a.batch.get_state()
            assert_state_bits(ref, b.StepHost(st, 2), "wide step_host after add_body")
            b.batch.set_state(ref)
    finally:
        a.close(); b.close()
