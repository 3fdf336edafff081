"""Host-side mirror of the reference's scene-description types.

The reference keeps these in F# and keeps computing them on the host even after the swap
(SURVEY.md §8(b) "What F# keeps computing"): validation, yaw/pitch/roll -> R, inertia and its
inverse, gravity force per body.  The native side only ever sees the evaluated arrays.

Reference (relative to /root/reference/PhysicsSimulatorCore/):
  RigidBodyPrototype.fs:10-90, Entities/Base.fs:13-20, Entities/Box.fs:12-20,
  Entities/Sphere.fs:10-12, Utilities/Matrix3.fs:90-108, SimulatorState.fs:26-36.
"""
from __future__ import annotations

import dataclasses
import math
import sys
from dataclasses import dataclass, field
from typing import Sequence, Tuple, Union

import numpy as np

INFINITE = math.inf  # Mass.Infinite (Entities/Base.fs:13-20): static <=> mass == +inf

SPHERE = 0
BOX = 1


@dataclass(frozen=True)
class Sphere:
    """Entities/Sphere.fs:5-17"""

    Radius: float

    def __post_init__(self):
        if self.Radius < 0.0:  # throwIfNegative, Sphere.fs:15
            raise ValueError("radius: Must be positive")

    def CreateRotationalInertia(self, mass: float) -> Tuple[float, float, float]:
        i = 2.0 * mass * self.Radius * self.Radius / 5.0  # Sphere.fs:11
        return (i, i, i)


@dataclass(frozen=True)
class Box:
    """Entities/Box.fs:6-20"""

    XSize: float
    YSize: float
    ZSize: float

    def __post_init__(self):
        for name in ("XSize", "YSize", "ZSize"):
            if getattr(self, name) < 0.0:  # Box.fs:51-54
                raise ValueError(f"{name}: Must be positive")

    def CreateRotationalInertia(self, mass: float) -> Tuple[float, float, float]:
        m = mass / 12.0  # Box.fs:13-19
        a2 = self.XSize * self.XSize
        b2 = self.YSize * self.YSize
        c2 = self.ZSize * self.ZSize
        return (m * (b2 + c2), m * (a2 + c2), m * (a2 + b2))


Kind = Union[Sphere, Box]


def from_yaw_pitch_roll(yaw: float, pitch: float, roll: float) -> np.ndarray:
    """RotationMatrix3D.fromYawPitchRoll (Matrix3.fs:90-108), quirk Q11 included:
    element [1,1] really is cosA*sinB*sinG + cosA*cosG."""
    cosA, sinA = math.cos(yaw), math.sin(yaw)
    cosB, sinB = math.cos(pitch), math.sin(pitch)
    cosG, sinG = math.cos(roll), math.sin(roll)
    return np.array(
        [
            [cosB * cosG, sinA * sinB * cosG - cosA * sinG, cosA * sinB * cosG + sinA * sinG],
            [cosB * sinG, cosA * sinB * sinG + cosA * cosG, cosA * sinB * sinG - sinA * cosG],
            [-sinB, sinA * cosB, cosA * cosB],
        ],
        dtype=np.float64,
    )


def _smul(s: float, v: Sequence[float]) -> Tuple[float, float, float]:
    """MathNet scalar*vector: 1 -> copy, 0 -> zeros."""
    if s == 1.0:
        return (float(v[0]), float(v[1]), float(v[2]))
    if s == 0.0:
        return (0.0, 0.0, 0.0)
    return (s * v[0], s * v[1], s * v[2])


EARTH_GRAVITY_MAGNITUDE = 9.81  # StepConfiguration.earthGravityAccelMagnitude (Configuration.fs:43)
EARTH_GRAVITY_DIRECTION = (0.0, 0.0, -1.0)  # Configuration.fs:45


def gravity_force(mass: float, magnitude: float, direction: Sequence[float]):
    """SimulatorStateBuilder.gravityForce (SimulatorState.fs:26-27): mass * (mag * dir)."""
    return _smul(mass, _smul(magnitude, direction))


@dataclass
class RigidBodyPrototype:
    """RigidBodyPrototype record (RigidBodyPrototype.fs:10-26). Collider always equals Kind in the
    reference's constructors (:31-36); it is kept as a separate field like there."""

    Kind: Kind
    Position: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    Velocity: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    Mass: float = 1.0
    Yaw: float = 0.0
    Pitch: float = 0.0
    Roll: float = 0.0
    AngularVelocity: Tuple[float, float, float] = (0.0, 0.0, 0.0)  # ignored by build (Q12)
    ElasticityCoeff: float = 0.1  # RigidBodyPrototype.fs:29
    StaticFrictionCoeff: float = 0.5  # :30
    KineticFrictionCoeff: float = 0.5 + 0.1  # :50
    UseGravity: bool = True
    Collider: Kind = None  # type: ignore[assignment]

    def __post_init__(self):
        if self.Collider is None:
            self.Collider = self.Kind

    def With(self, **kw) -> "RigidBodyPrototype":
        """F# `{ proto with Field = value }`"""
        return dataclasses.replace(self, **kw)


def createDefaultBox(xSize: float, ySize: float, zSize: float) -> RigidBodyPrototype:
    return RigidBodyPrototype(Kind=Box(xSize, ySize, zSize))  # RigidBodyPrototype.fs:52-53


def createDefaultCube(size: float) -> RigidBodyPrototype:
    return RigidBodyPrototype(Kind=Box(size, size, size))  # :55-56


def createDefaultSphere(radius: float) -> RigidBodyPrototype:
    return RigidBodyPrototype(Kind=Sphere(radius))  # :58-59


def validate(proto: RigidBodyPrototype) -> None:
    """RigidBodyPrototype.validate (:61-67) + RigidBody.create checks (RigidBody.fs:58-65)."""
    if math.isinf(proto.Mass) and proto.UseGravity:
        raise ValueError("UseGravity: objects with infinite mass cannot use gravity")
    if proto.ElasticityCoeff < 0.0 or proto.ElasticityCoeff > 1.0:
        raise ValueError("elasticityCoeff: Invalid value")
    if proto.StaticFrictionCoeff < 0.0:
        raise ValueError("staticFrictionCoeff: Invalid value")
    if proto.KineticFrictionCoeff < 0.0:
        raise ValueError("staticFrictionCoeff: Invalid value")


BODY_FIELDS = (
    ("shape", np.int32, ()),
    ("dims", np.float64, (3,)),
    ("mass", np.float64, ()),
    ("inertia_inv", np.float64, (3,)),
    ("pos", np.float64, (3,)),
    ("vel", np.float64, (3,)),
    ("rot", np.float64, (9,)),
    ("ang_mom", np.float64, (3,)),
    ("force", np.float64, (3,)),
    ("torque", np.float64, (3,)),
    ("elasticity", np.float64, ()),
    ("mu_kinetic", np.float64, ()),
    ("mu_static", np.float64, ()),
)


def empty_bodies(n: int) -> dict:
    return {name: np.zeros((n,) + shape, dtype=dt) for name, dt, shape in BODY_FIELDS}


def build_into(arr: dict, i: int, proto: RigidBodyPrototype,
               gravity_magnitude: float = EARTH_GRAVITY_MAGNITUDE,
               gravity_direction: Sequence[float] = EARTH_GRAVITY_DIRECTION) -> None:
    """RigidBodyPrototype.build (:69-90) + default external force (SimulatorState.fs:29-36),
    written into row i of the SoA arrays the C ABI takes (include/ps_b200.h ps_bodies_view)."""
    validate(proto)
    kind = proto.Collider
    if isinstance(kind, Sphere):
        arr["shape"][i] = SPHERE
        arr["dims"][i] = (kind.Radius, 0.0, 0.0)
    else:
        arr["shape"][i] = BOX
        arr["dims"][i] = (kind.XSize, kind.YSize, kind.ZSize)
    inertia = proto.Kind.CreateRotationalInertia(proto.Mass)
    for k in range(3):
        d = inertia[k]
        if d == 0.0:  # DiagonalMatrix.Inverse throws on a zero diagonal (RigidBody.fs:76)
            raise ValueError("principalRotationalInertia: Matrix must not be singular")
        arr["inertia_inv"][i, k] = 1.0 / d
    arr["mass"][i] = proto.Mass
    arr["pos"][i] = proto.Position
    arr["vel"][i] = proto.Velocity
    arr["rot"][i] = from_yaw_pitch_roll(proto.Yaw, proto.Pitch, proto.Roll).reshape(9)
    arr["ang_mom"][i] = (0.0, 0.0, 0.0)  # RigidBody.fs:69 (Q12)
    arr["force"][i] = (
        gravity_force(proto.Mass, gravity_magnitude, gravity_direction) if proto.UseGravity else (0.0, 0.0, 0.0)
    )
    arr["torque"][i] = (0.0, 0.0, 0.0)
    arr["elasticity"][i] = proto.ElasticityCoeff
    arr["mu_kinetic"][i] = proto.KineticFrictionCoeff
    arr["mu_static"][i] = proto.StaticFrictionCoeff


def build_world(prototypes: Sequence[RigidBodyPrototype], gravity_magnitude: float = EARTH_GRAVITY_MAGNITUDE,
                gravity_direction: Sequence[float] = EARTH_GRAVITY_DIRECTION) -> dict:
    """SimulatorStateBuilder.fromPrototypes + withConfiguration (SimulatorState.fs:38-83) for one world:
    ids are the list index (:41-42); non-zero default forces are recomputed with the configured gravity
    (:67-83)."""
    arr = empty_bodies(len(prototypes))
    for i, p in enumerate(prototypes):
        build_into(arr, i, p, gravity_magnitude, gravity_direction)
    return arr


def concat_worlds(worlds: Sequence[dict]):
    """Concatenate per-world SoA dicts; returns (arrays, world_offset int32[n+1])."""
    offs = np.zeros(len(worlds) + 1, dtype=np.int32)
    for k, w in enumerate(worlds):
        offs[k + 1] = offs[k] + len(w["mass"])
    out = {name: np.ascontiguousarray(np.concatenate([w[name] for w in worlds], axis=0)) for name, _, _ in BODY_FIELDS}
    return out, offs


def ball_socket_local_anchors(world: dict, a: int, b: int, anchor_world: Sequence[float]):
    """Joint.createBallSocket (Joints.fs:43-53): local anchor = R^T * (anchor_world - x), evaluated
    like MathNet's dense M*v (sum over k ascending from 0.0)."""
    out = []
    for body in (a, b):
        R = world["rot"][body].reshape(3, 3)
        d = np.asarray(anchor_world, dtype=np.float64) - world["pos"][body]
        loc = []
        for i in range(3):
            s = 0.0
            for k in range(3):
                s += R[k, i] * d[k]  # (R^T)[i][k] = R[k][i]
            loc.append(s)
        out.append(tuple(loc))
    return out[0], out[1]


def getAABoundingBox(proto: RigidBodyPrototype):
    """RigidBodyPrototype.getAABoundingBox (RigidBodyPrototype.fs:92-95) = SimulatorObject.getAABoundingBox of the built
    body (SimulatorObject.fs:58-71): (centre, size).  A sphere's box has side = RADIUS (quirk Q5); a box takes
    Box.worldAxisAlignedSize (Box.fs:69-85), evaluated like MathNet's dense M*v (zero-started sums)."""
    arr = empty_bodies(1)
    build_into(arr, 0, proto)
    c = tuple(float(v) for v in arr["pos"][0])
    if int(arr["shape"][0]) == SPHERE:
        r = float(arr["dims"][0, 0])
        return c, (r, r, r)
    R = arr["rot"][0].reshape(3, 3)
    h = [float(arr["dims"][0, k]) / 2.0 for k in range(3)]

    def to_world(local):  # R * local + 0, dense M*v summed from 0.0 in index order
        out = []
        for i in range(3):
            acc = 0.0
            for q in range(3):
                acc = acc + float(R[i, q]) * local[q]
            out.append(acc + 0.0)
        return out

    ex, ey, ez = to_world([h[0], 0.0, 0.0]), to_world([0.0, h[1], 0.0]), to_world([0.0, 0.0, h[2]])
    size = tuple((abs(ex[i]) + abs(ey[i]) + abs(ez[i])) * 2.0 for i in range(3))
    return c, size


class SpatialTreeBoundaries:
    """SpatialTreeBoundaries module (RigidBodyPrototype.fs:97-132)."""

    @staticmethod
    def fromPrototypeAABBs(prototypes: Sequence[RigidBodyPrototype], padding: float, motionMargin: Sequence[float]):
        """Union of the prototype AABBs grown by `padding` (static geometry) and `motionMargin` (dynamic travel):
        returns (Min, Max) of the SpaceBoundaries record (Configuration.fs:9-11)."""
        expand = [padding + float(m) for m in motionMargin]
        big = sys.float_info.max
        mn, mx = [big, big, big], [-big, -big, -big]
        for p in prototypes:
            c, s = getAABoundingBox(p)
            for k in range(3):
                half = s[k] * (1.0 / 2.0)  # Vector3D `/`: multiply by 1/s (Vector3D.fs:34)
                mn[k] = min(c[k] - half, mn[k])
                mx[k] = max(c[k] + half, mx[k])
        return tuple(mn[k] - expand[k] for k in range(3)), tuple(mx[k] + expand[k] for k in range(3))
