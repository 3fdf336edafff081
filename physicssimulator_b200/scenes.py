"""Scene builders: mirrors of the reference's Gui/Scenes fixtures and the synthetic configs C1-C5 of
BASELINE.json.  Everything is seeded with SplitMix64; doubles are (u>>11)*2^-53.

Reference (relative to /root/reference/): Gui/Scenes/StackScene.fs:10-44, Gui/Scenes/DominoScene.fs:13-67.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np

from . import prototypes as P
from .configuration import (Configuration, StepConfiguration, SpatialTreeConfiguration,
                            BroadPhaseCollisionDetectionKind)

MASK = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed: int):
        self.s = seed & MASK

    def next_u64(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & MASK
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK
        return z ^ (z >> 31)

    def uniform(self, lo: float = 0.0, hi: float = 1.0) -> float:
        return lo + (hi - lo) * ((self.next_u64() >> 11) * (1.0 / 9007199254740992.0))


def scene_seed(config: int, world: int) -> int:
    return (0x5EED0000 + (config << 32) + world) & MASK


def _ground(x: float, y: float, z: float, pos=(0.0, 0.0, -0.5), yaw: float = 0.0) -> P.RigidBodyPrototype:
    return P.createDefaultBox(x, y, z).With(Mass=P.INFINITE, UseGravity=False, Position=pos, Yaw=yaw)


# ---------------------------------------------------------------------------------------------
# Gui/Scenes mirrors
# ---------------------------------------------------------------------------------------------
def create_vertical_stack(position, height: int, box_size: float, mass: float) -> List[P.RigidBodyPrototype]:
    """StackSceneData.createVerticalStack (Gui/Scenes/StackScene.fs:13-19)."""
    out = []
    for i in range(height):
        dz = float(i) * (box_size + 0.01)
        out.append(P.createDefaultCube(box_size).With(
            ElasticityCoeff=0.1, Mass=mass,
            Position=(position[0] + 0.0 * (box_size + 0.01), position[1] + 0.0 * (box_size + 0.01), position[2] + dz)))
    return out


def stack_scene() -> Tuple[List[P.RigidBodyPrototype], Configuration]:
    """StackSceneData.prototypes / configuration (Gui/Scenes/StackScene.fs:21-44)."""
    protos = [_ground(15.0, 15.0, 1.01)]
    protos += create_vertical_stack((0.0, 0.0, 1.5), 3, 1.0, 100.0)
    protos += [P.createDefaultCube(0.5).With(Mass=50.0, UseGravity=True, Position=(0.0, -5.0, 0.0))]
    cfg = Configuration(BroadPhaseCollisionDetectionKind=BroadPhaseCollisionDetectionKind(
        SpatialTreeConfiguration(LeafCapacity=2, MaxDepth=10)))
    return protos, cfg


def domino_scene() -> Tuple[List[P.RigidBodyPrototype], Configuration]:
    """DominoSceneData (Gui/Scenes/DominoScene.fs:13-67), SpatialTree bounds from SpatialTreeBoundaries.fromPrototypeAABBs
    as there (:41-55)."""
    count, thick, width, height, mass = 15, 0.08, 0.35, 1.2, 100.0
    spacing = thick + 0.4
    origin_y = -(float(count - 1) * spacing) / 2.0
    protos = [_ground(12.0, 12.0, 1.01)]
    for i in range(count):
        protos.append(P.createDefaultBox(width, thick, height).With(
            Mass=mass, ElasticityCoeff=0.05, StaticFrictionCoeff=0.7, KineticFrictionCoeff=0.55,
            Position=(0.0, origin_y + float(i) * spacing, height / 2.0)))
    impulse_strength = 120.0
    tip_reach = height + width  # "AABB grows to ~height when a domino tips over"
    margin = tip_reach + impulse_strength / mass * 2.0
    mn, mx = P.SpatialTreeBoundaries.fromPrototypeAABBs(protos, 2.0, (margin, margin, margin))
    cfg = Configuration(BroadPhaseCollisionDetectionKind=BroadPhaseCollisionDetectionKind(
        SpatialTreeConfiguration(LeafCapacity=4, SpaceBoundariesMin=mn, SpaceBoundariesMax=mx, MaxDepth=12)))
    return protos, cfg


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs
# ---------------------------------------------------------------------------------------------
def c1_stack20() -> Tuple[List[P.RigidBodyPrototype], Configuration]:
    """C1: StackScene-style ground + 20 unit cubes in 4 columns of 5 + 4 dropped spheres, friction on."""
    protos = [_ground(15.0, 15.0, 1.01)]
    cols = [(1.5, 1.5), (1.5, -1.5), (-1.5, 1.5), (-1.5, -1.5)]
    for (cx, cy) in cols:
        protos += create_vertical_stack((cx, cy, 1.5), 5, 1.0, 100.0)
    for k, (cx, cy) in enumerate(cols):
        protos.append(P.createDefaultSphere(0.4).With(Mass=20.0, Position=(cx, cy, 8.0 + float(k))))
    return protos, Configuration()


def c2_spheres64(world: int, n_spheres: int = 64) -> Tuple[List[P.RigidBodyPrototype], Configuration]:
    """C2: ground (Yaw=pi so its local -z face is up, Q4) + 64 spheres on a jittered 4x4x4 lattice; no friction."""
    rng = SplitMix64(scene_seed(2, world))
    protos = [_ground(40.0, 40.0, 1.0, yaw=math.pi)]
    side = max(1, round(n_spheres ** (1.0 / 3.0)))
    for k in range(n_spheres):
        ix, iy, iz = k % side, (k // side) % side, k // (side * side)
        r = rng.uniform(0.3, 0.5)
        pos = ((ix - (side - 1) / 2.0) * 1.2 + rng.uniform(-0.1, 0.1),
               (iy - (side - 1) / 2.0) * 1.2 + rng.uniform(-0.1, 0.1),
               1.0 + iz * 1.2 + rng.uniform(-0.1, 0.1))
        vel = (rng.uniform(-1.0, 1.0), rng.uniform(-1.0, 1.0), rng.uniform(-1.0, 1.0))
        mass = 4.0 / 3.0 * math.pi * r * r * r * 1000.0 * 1e-3
        protos.append(P.createDefaultSphere(r).With(Mass=mass, Position=pos, Velocity=vel))
    cfg = Configuration(StepConfig=StepConfiguration(EnableFriction=False))
    return protos, cfg


def c3_mixed128(world: int, n_boxes: int = 64, n_spheres: int = 63) -> Tuple[List[P.RigidBodyPrototype], Configuration]:
    """C3: ground + 64 boxes + 63 spheres on a jittered 8x4x4 lattice (pitch 1.3), friction on."""
    rng = SplitMix64(scene_seed(3, world))
    protos = [_ground(40.0, 40.0, 1.0, yaw=math.pi)]
    n = n_boxes + n_spheres
    slots = []
    iz = 0
    while len(slots) < n:
        for iy in range(4):
            for ix in range(8):
                slots.append((ix, iy, iz))
        iz += 1
    for k in range(n):
        ix, iy, iz = slots[k]
        pos = ((ix - 3.5) * 1.3 + rng.uniform(-0.1, 0.1), (iy - 1.5) * 1.3 + rng.uniform(-0.1, 0.1),
               1.2 + iz * 1.3 + rng.uniform(-0.1, 0.1))
        if k < n_boxes:
            ex, ey, ez = rng.uniform(0.4, 1.0), rng.uniform(0.4, 1.0), rng.uniform(0.4, 1.0)
            yaw, pitch = rng.uniform(-math.pi, math.pi), rng.uniform(-math.pi, math.pi)
            protos.append(P.createDefaultBox(ex, ey, ez).With(Mass=ex * ey * ez * 500.0, Position=pos, Yaw=yaw, Pitch=pitch))
        else:
            r = rng.uniform(0.2, 0.5)
            protos.append(P.createDefaultSphere(r).With(Mass=4.0 / 3.0 * math.pi * r * r * r * 500.0, Position=pos))
    return protos, Configuration()


def c4_chain32(world: int, n_links: int = 32):
    """C4: ground 80x80 + static anchor cube + 32 box links 0.2x0.2x1.0 in a nearly horizontal chain (pitch 1.35 +- 0.1,
    any yaw) whose lowest link starts ~1 m above the ground; ball-socket joints (1,2),(2,3)...; restore_joints=1.
    Joint.restore as written (Joints.fs:55-94) applies K.(-vRel) N times: a velocity feedback of gain ~ N.|K|^2 with
    K ~ 1/mass.  With the links at 1000 kg the gain stays below 2 and the chains fall, land, pile up and come to rest
    with finite states (oracle: 800 steps, ~53 colliding pairs and ~167 contacts per world-step at rest); lighter links
    diverge (`c4_chain32_blowup`, kept for parity through the blow-up).
    Returns (prototypes, configuration, joints[(a, b, anchor_world)])."""
    rng = SplitMix64(scene_seed(4, world))
    protos = [_ground(80.0, 80.0, 1.01)]
    yaw = rng.uniform(-math.pi, math.pi)
    pitch = 1.35 + rng.uniform(-0.1, 0.1)
    # link long axis = local z in the reference's yaw/pitch/roll convention (Matrix3.fs:90-108)
    R = P.from_yaw_pitch_roll(yaw, pitch, 0.0)
    axis = R[:, 2]
    if axis[2] < 0.0:
        axis = -axis
    span = axis * float(n_links)
    top = (rng.uniform(-0.5, 0.5) + span[0] / 2.0, rng.uniform(-0.5, 0.5) + span[1] / 2.0, 1.5 + span[2])
    protos.append(P.createDefaultCube(0.2).With(Mass=P.INFINITE, UseGravity=False, Position=top))
    joints = []
    end = np.array(top)
    for k in range(n_links):
        centre = end - axis * 0.5
        protos.append(P.createDefaultBox(0.2, 0.2, 1.0).With(Mass=1000.0, Position=tuple(centre), Yaw=yaw, Pitch=pitch))
        joints.append((1 + k, 2 + k, tuple(end)))
        end = end - axis * 1.0
    cfg = Configuration(RestoreJoints=True)
    return protos, cfg, joints


def c4_chain32_blowup(world: int, n_links: int = 32):
    """The first C4 scene (round 1): 5 kg links hanging diagonally (pitch 0.3) from an anchor at z = 12 over a 15x15
    ground.  With light links Joint.restore as written is an unstable feedback: the chains reach NaN within ~35
    steps, in the reference semantics too.  Kept as a parity-only case (NaN = NaN through the blow-up)."""
    rng = SplitMix64(scene_seed(4, world))
    protos = [_ground(15.0, 15.0, 1.01)]
    top = (rng.uniform(-0.5, 0.5), rng.uniform(-0.5, 0.5), 12.0)
    protos.append(P.createDefaultCube(0.2).With(Mass=P.INFINITE, UseGravity=False, Position=top))
    pitch = 0.3
    R = P.from_yaw_pitch_roll(0.0, pitch, 0.0)
    axis = R[:, 2]
    joints = []
    end = np.array(top)
    for k in range(n_links):
        centre = end - axis * 0.5
        protos.append(P.createDefaultBox(0.2, 0.2, 1.0).With(Mass=5.0, Position=tuple(centre), Pitch=pitch))
        joints.append((1 + k, 2 + k, tuple(end)))
        end = end - axis * 1.0
    cfg = Configuration(RestoreJoints=True)
    return protos, cfg, joints


def c5_big(n_spheres: int = 1_000_000, n_boxes: int = 100_000, seed_world: int = 0) -> dict:
    """C5: one world, ground 1100x1100x1 + spheres r in U[0.1,0.3] + boxes edges U[0.2,1.0] with random
    orientation, centres uniform in [-500,500]^2 x [0.5,20].  Returns the SoA arrays the broad phase needs
    (shape, dims, mass, pos, rot).  Uses numpy's PCG64 seeded from the SplitMix64 scene seed (1.1 M bodies
    in pure Python would take minutes)."""
    n = 1 + n_spheres + n_boxes
    rs = np.random.Generator(np.random.PCG64(scene_seed(5, seed_world)))
    shape = np.zeros(n, dtype=np.int32)
    dims = np.zeros((n, 3))
    mass = np.ones(n)
    pos = np.zeros((n, 3))
    rot = np.tile(np.eye(3).reshape(9), (n, 1))
    shape[0] = P.BOX
    dims[0] = (1100.0, 1100.0, 1.0)
    mass[0] = np.inf
    pos[0] = (0.0, 0.0, -0.5)
    s = slice(1, 1 + n_spheres)
    dims[s, 0] = rs.uniform(0.1, 0.3, n_spheres)
    b = slice(1 + n_spheres, n)
    shape[b] = P.BOX
    dims[b] = rs.uniform(0.2, 1.0, (n_boxes, 3))
    scale = (n - 1) / 1_100_000.0  # keep the density of the full-size scene when subsampling
    half = 500.0 * math.sqrt(max(scale, 1e-12))
    pos[1:, 0] = rs.uniform(-half, half, n - 1)
    pos[1:, 1] = rs.uniform(-half, half, n - 1)
    pos[1:, 2] = rs.uniform(0.5, 20.0, n - 1)
    yaw = rs.uniform(-math.pi, math.pi, n_boxes)
    pitch = rs.uniform(-math.pi, math.pi, n_boxes)
    cA, sA, cB, sB = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch)
    cG, sG = np.ones(n_boxes), np.zeros(n_boxes)
    R = np.stack([cB * cG, sA * sB * cG - cA * sG, cA * sB * cG + sA * sG,
                  cB * sG, cA * sB * sG + cA * cG, cA * sB * sG - sA * cG,
                  -sB, sA * cB, cA * cB], axis=1)
    rot[b] = R
    return {"shape": shape, "dims": np.ascontiguousarray(dims), "mass": mass,
            "pos": np.ascontiguousarray(pos), "rot": np.ascontiguousarray(rot)}
