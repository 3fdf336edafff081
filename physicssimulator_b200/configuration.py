"""Configuration records, mirroring the reference (relative to /root/reference/PhysicsSimulatorCore/):
Configuration.fs:5-55 (StepConfiguration, SpatialTreeConfiguration, BroadPhaseCollisionDetectionKind,
RigidBodyIntegratorKind) and Simulator.fs:16-27 (Configuration)."""
from __future__ import annotations

import ctypes
import dataclasses
from dataclasses import dataclass, field
from typing import Optional, Tuple

FirstOrder = 0  # RigidBodyIntegratorKind (Configuration.fs:5-7)
AugmentedSecondOrder = 1


@dataclass(frozen=True)
class SpatialTreeConfiguration:
    """Configuration.fs:13-23"""

    LeafCapacity: int = 4
    SpaceBoundariesMin: Tuple[float, float, float] = (-15.0, -15.0, -15.0)
    SpaceBoundariesMax: Tuple[float, float, float] = (15.0, 15.0, 15.0)
    MaxDepth: int = 10


@dataclass(frozen=True)
class BroadPhaseCollisionDetectionKind:
    """Dummy | SpatialTree of SpatialTreeConfiguration (Configuration.fs:24-26)."""

    SpatialTree: Optional[SpatialTreeConfiguration] = None

    @property
    def IsDummy(self) -> bool:
        return self.SpatialTree is None


Dummy = BroadPhaseCollisionDetectionKind(None)


@dataclass(frozen=True)
class StepConfiguration:
    """Configuration.fs:28-55 (defaults = StepConfiguration.getDefault)."""

    Epsilon: float = 0.0001
    GravitationalAccelerationMagnitude: float = 9.81
    GravityDirection: Tuple[float, float, float] = (0.0, 0.0, -1.0)
    BaumgarteTerm: float = 0.2
    AllowedPenetration: float = 0.001
    MaxCorrectionVelocity: float = 1.0
    CollisionSolverIterationCount: int = 10
    EnableFriction: bool = True
    RigidBodyIntegratorKind: int = FirstOrder

    def With(self, **kw) -> "StepConfiguration":
        return dataclasses.replace(self, **kw)


@dataclass(frozen=True)
class Configuration:
    """Simulator.fs:16-27 (defaults = Configuration.getDefault). SimulationStepInterval in seconds."""

    StepConfig: StepConfiguration = field(default_factory=StepConfiguration)
    SimulationSpeedMultiplier: float = 1.0
    SimulationStepInterval: float = 0.010
    BroadPhaseCollisionDetectionKind: BroadPhaseCollisionDetectionKind = Dummy
    # additive (not in the reference): the line the reference keeps commented out (Simulator.fs:55)
    RestoreJoints: bool = False

    def With(self, **kw) -> "Configuration":
        return dataclasses.replace(self, **kw)


class ps_step_config(ctypes.Structure):
    """include/ps_b200.h ps_step_config"""

    _fields_ = [
        ("epsilon", ctypes.c_double),
        ("baumgarte", ctypes.c_double),
        ("allowed_penetration", ctypes.c_double),
        ("max_correction_velocity", ctypes.c_double),
        ("dt", ctypes.c_double),
        ("solver_iterations", ctypes.c_int32),
        ("enable_friction", ctypes.c_int32),
        ("integrator_kind", ctypes.c_int32),
        ("broadphase_kind", ctypes.c_int32),
        ("restore_joints", ctypes.c_int32),
        ("record_contacts", ctypes.c_int32),
        ("exact_all_pairs", ctypes.c_int32),
        ("tree_leaf_capacity", ctypes.c_int32),
        ("tree_max_depth", ctypes.c_int32),
        ("aabb_mode", ctypes.c_int32),
        ("tree_min", ctypes.c_double * 3),
        ("tree_max", ctypes.c_double * 3),
    ]


def to_native_config(cfg: Configuration, record_contacts: bool = False, exact_all_pairs: bool = False, aabb_mode: int = 0) -> ps_step_config:
    s = cfg.StepConfig
    t = cfg.BroadPhaseCollisionDetectionKind.SpatialTree
    tree = t or SpatialTreeConfiguration()
    return ps_step_config(
        epsilon=s.Epsilon,
        baumgarte=s.BaumgarteTerm,
        allowed_penetration=s.AllowedPenetration,
        max_correction_velocity=s.MaxCorrectionVelocity,
        dt=cfg.SimulationStepInterval,
        solver_iterations=s.CollisionSolverIterationCount,
        enable_friction=1 if s.EnableFriction else 0,
        integrator_kind=int(s.RigidBodyIntegratorKind),
        broadphase_kind=0 if t is None else 1,
        tree_leaf_capacity=int(tree.LeafCapacity),
        tree_max_depth=int(tree.MaxDepth),
        tree_min=(ctypes.c_double * 3)(*[float(v) for v in tree.SpaceBoundariesMin]),
        tree_max=(ctypes.c_double * 3)(*[float(v) for v in tree.SpaceBoundariesMax]),
        restore_joints=1 if cfg.RestoreJoints else 0,
        record_contacts=1 if record_contacts else 0,
        exact_all_pairs=1 if exact_all_pairs else 0,
        aabb_mode=int(aabb_mode),
    )
