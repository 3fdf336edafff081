"""Simulator / BatchSimulator: the reference's public API for the step path, re-routed to the C ABI.

Mirrors (relative to /root/reference/PhysicsSimulatorCore/):
  Simulator.fs:29-166   Simulator class: ObjectsIdentifiers, CollisionsIdentifiers, PhysicalObject(id),
                        AllPhysicalObjects, Collision(idSet), AllCollisions, Configuration, AddObject,
                        ApplyImpulse, Reset, Start/Stop/PauseResume (pacing shell, host side only)
  SimulatorState.fs:38-107 SimulatorStateBuilder.fromPrototypes / withConfiguration / withPrototype

The only thing that changed is the body of updateSimulation (Simulator.fs:47-62): it is now
"marshal -> ps_batch_step -> unmarshal".  BatchSimulator is the additive multi-world type SURVEY.md §8(b)
names; Simulator is a one-world BatchSimulator.

The async timer loop (Simulator.Start, Simulator.fs:123-145) is pacing, not hot path: `Step(n)` is the
headless equivalent of n timer ticks.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _native
from . import prototypes as P
from .configuration import Configuration, to_native_config


@dataclass
class PhysicalObjectView:
    """What Simulator.PhysicalObject(id) exposes of a RigidBody (Entities/RigidBody.fs:6-17)."""

    Id: int
    Position: np.ndarray
    Velocity: np.ndarray
    Orientation: np.ndarray  # 3x3 row-major
    AngularMomentum: np.ndarray
    Mass: float
    IsStatic: bool


@dataclass
class ContactPoint:
    """CollisionDetection.fs:9-17 (+ the derived feature id, include/ps_b200.h)"""

    Normal: np.ndarray
    Position: np.ndarray
    Penetration: float
    FeatureId: int


class BatchSimulator:
    """Many independent worlds stepped together on one GPU.

    worlds: sequence of prototype lists (ids = list index, SimulatorState.fs:41-42)
    joints: optional per-world lists of (idA, idB, anchorWorldPosition) ball-socket joints (Joints.fs:43-53);
            only used when configuration.RestoreJoints is set (the reference keeps that line commented out).
    """

    def __init__(self, worlds: Sequence[Sequence[P.RigidBodyPrototype]], configuration: Optional[Configuration] = None,
                 joints: Optional[Sequence[Sequence[Tuple[int, int, Sequence[float]]]]] = None, device: int = -1,
                 record_contacts: bool = False, exact_all_pairs: bool = False, _prebuilt=None, aabb_mode: int = 0):
        self.configuration = configuration or Configuration()
        # BroadPhaseCollisionDetectionKind.SpatialTree: the candidate ORDER is the persistent tree's (history dependent,
        # Q16); the library keeps one tree per world on the host and feeds the device the ordered list every step
        sc = self.configuration.StepConfig
        if _prebuilt is not None:
            self._worlds = _prebuilt
        else:
            self._worlds = [P.build_world(list(w), sc.GravitationalAccelerationMagnitude, sc.GravityDirection) for w in worlds]
        self._initial = [{k: v.copy() for k, v in w.items()} for w in self._worlds]
        self._record = record_contacts
        self._exact = exact_all_pairs
        self._device = device
        self._joints_spec = joints
        self._native_cfg = to_native_config(self.configuration, record_contacts, exact_all_pairs, aabb_mode)
        self._create()

    # -- construction ---------------------------------------------------------------------------------
    def _create(self):
        bodies, off = P.concat_worlds(self._worlds)
        jl = []
        if self._joints_spec:
            for w, js in enumerate(self._joints_spec):
                for (a, b, anchor) in js:
                    la, lb = P.ball_socket_local_anchors(self._worlds[w], a, b, anchor)
                    jl.append((w, a, b, la, lb))
        self._joints_native = jl
        self._offsets = off
        self._mass = bodies["mass"].copy()
        self.batch = _native.Batch(bodies, off, self._native_cfg, joints=jl or None, device=self._device)

    @property
    def NumWorlds(self) -> int:
        return len(self._worlds)

    # -- the step ---------------------------------------------------------------------------------------
    def Step(self, n: int = 1):
        """n x Simulator.updateSimulation (Simulator.fs:47-62) on every world."""
        self.batch.step(n)

    # -- read-back --------------------------------------------------------------------------------------
    def StepHost(self, state: dict, n: int = 1, out: Optional[dict] = None) -> dict:
        """marshal -> n steps -> unmarshal in one native call (the shape of Simulator.updateSimulation, Simulator.fs:47-62,
        for a caller that keeps the state in host memory): copies overlap the step, chunk of worlds by chunk of worlds."""
        return self.batch.step_host(state, n, out)

    def State(self, w0: int = 0, w1: Optional[int] = None) -> dict:
        return self.batch.get_state(w0, w1)

    def WorldState(self, world: int) -> dict:
        return self.batch.get_state(world, world + 1)

    def Collisions(self, world: int) -> dict:
        return self.batch.contacts(world)

    def Candidates(self, world: int) -> np.ndarray:
        return self.batch.candidates(world)

    def Stats(self) -> dict:
        return self.batch.stats()

    # -- mutations (Simulator.fs:91-121) ------------------------------------------------------------------
    def ApplyImpulse(self, world: int, objectId: int, impulse, offset):
        self.batch.apply_impulse(world, objectId, impulse, offset)

    def AddObject(self, world: int, proto: P.RigidBodyPrototype) -> int:
        """SimulatorStateBuilder.withPrototype (SimulatorState.fs:89-107): id = max id + 1; gravity from the
        hard-coded earth constants (Q21)."""
        one = P.empty_bodies(1)
        P.build_into(one, 0, proto)  # earth gravity, not the configured one
        self.batch.add_body(world, one)
        w = self._worlds[world]
        for k in w:
            w[k] = np.concatenate([w[k], one[k]], axis=0)
        return len(w["mass"]) - 1

    def Reset(self):
        """Simulator.Reset (Simulator.fs:113-121): rebuild from the constructor's prototypes."""
        self._worlds = [{k: v.copy() for k, v in w.items()} for w in self._initial]
        _, off = P.concat_worlds(self._worlds)
        self.batch.reset(off)

    def close(self):
        self.batch.close()


class Simulator:
    """Single-world facade with the reference's member names (Simulator.fs:29-166)."""

    def __init__(self, simulatorObjects: Sequence[P.RigidBodyPrototype], configuration0: Optional[Configuration] = None,
                 device: int = -1, exact_all_pairs: bool = False):
        protos = list(simulatorObjects)
        cfg = configuration0 or Configuration()
        if len(protos) < 3:
            # SimulatorStateBuilder.fromPrototypes builds the hard-coded joint objectMap[1], objectMap[2]
            # (SimulatorState.fs:63-64): fewer than 3 prototypes -> KeyNotFoundException
            raise KeyError("The given key was not present in the dictionary (SimulatorState.fs:63-64 needs ids 1 and 2)")
        joints = [[(1, 2, (0.0, -5.0, 1.25))]] if cfg.RestoreJoints else None
        self._b = BatchSimulator([protos], cfg, joints=joints, device=device, record_contacts=True, exact_all_pairs=exact_all_pairs)
        self._state = "Stopped"
        self._listeners = []

    # events
    def SimulationStateChanged(self, handler):
        self._listeners.append(handler)

    def _trigger(self):
        for h in self._listeners:
            h(self)

    @property
    def State(self):
        return self._state

    @property
    def Configuration(self):
        return self._b.configuration

    @property
    def ObjectsIdentifiers(self) -> set:
        return set(range(len(self._b._worlds[0]["mass"])))

    @property
    def CollisionsIdentifiers(self) -> List[frozenset]:
        return [frozenset((int(i), int(j))) for i, j in self._b.Collisions(0)["pairs"]]

    def PhysicalObject(self, identifier: int) -> PhysicalObjectView:
        n = len(self._b._worlds[0]["mass"])
        if identifier < 0 or identifier >= n:
            raise KeyError(identifier)
        st = self._b.WorldState(0)
        m = float(self._b._worlds[0]["mass"][identifier])
        return PhysicalObjectView(identifier, st["pos"][identifier], st["vel"][identifier], st["rot"][identifier].reshape(3, 3),
                                  st["ang_mom"][identifier], m, math.isinf(m))

    SimulatorObject = PhysicalObject

    @property
    def AllPhysicalObjects(self) -> List[PhysicalObjectView]:
        st = self._b.WorldState(0)
        mass = self._b._worlds[0]["mass"]
        return [PhysicalObjectView(i, st["pos"][i], st["vel"][i], st["rot"][i].reshape(3, 3), st["ang_mom"][i], float(mass[i]),
                                   math.isinf(mass[i])) for i in range(len(mass))]

    def Collision(self, identifier) -> Optional[List[ContactPoint]]:
        key = tuple(sorted(int(x) for x in identifier))
        c = self._b.Collisions(0)
        for k, (i, j) in enumerate(c["pairs"]):
            if (int(i), int(j)) == key:
                a, b = c["first"][k], c["first"][k + 1]
                return [ContactPoint(c["normal"][q], c["position"][q], float(c["penetration"][q]), int(c["feature"][q])) for q in range(a, b)]
        return None

    @property
    def AllCollisions(self) -> List[List[ContactPoint]]:
        c = self._b.Collisions(0)
        out = []
        for k in range(len(c["pairs"])):
            a, b = c["first"][k], c["first"][k + 1]
            out.append([ContactPoint(c["normal"][q], c["position"][q], float(c["penetration"][q]), int(c["feature"][q])) for q in range(a, b)])
        return out

    def AddObject(self, obj: P.RigidBodyPrototype):
        self._b.AddObject(0, obj)
        self._trigger()

    def ApplyImpulse(self, physicalObjectIdentifier: int, impulseValue, impulseOffset):
        self._b.ApplyImpulse(0, physicalObjectIdentifier, impulseValue, impulseOffset)
        self._trigger()

    def Reset(self):
        self._b.Reset()
        self._trigger()

    def Step(self, n: int = 1):
        """n timer ticks of updateSimulation, headless."""
        self._b.Step(n)
        self._trigger()

    # pacing shell (Simulator.fs:123-166); the loop itself is the caller's
    def Start(self):
        if self._state != "Stopped":
            raise RuntimeError("simulator should be in stopped state")
        self._state = "Started"

    def Stop(self):
        if self._state == "Stopped":
            raise RuntimeError("simulator is already stopped")
        self._state = "Stopped"

    def PauseResume(self):
        if self._state == "Paused":
            self._state = "Started"
        elif self._state == "Started":
            self._state = "Paused"
        else:
            raise RuntimeError("simulator must be started or paused")

    def SetSimulationSpeedMultiplier(self, multiplier: float):
        if multiplier <= 0.0:
            raise ValueError("multiplier: must be positive")
        self._b.configuration = self._b.configuration.With(SimulationSpeedMultiplier=multiplier)

    def Dispose(self):
        self._b.close()
