"""Golden-trajectory snapshots, format "PSTJ" v1 (include/ps_snapshot.h).

The reference keeps `SimulatorState` (PhysicsSimulatorCore/SimulatorState.fs:11-19) in memory only; this is the
file that carries a trajectory — configuration, constant body parameters and, per recorded step, the SoA state, the
ordered candidate list, the colliding pairs, the contact lists and the contact-feature ids — between the F#
reference (driver in INTEGRATION.md §6), the CPU oracle and the CUDA path.  Host-side only: numpy + struct, no
device code and no dependency on the oracle (tests hand any stepper in through `record`).
"""
from __future__ import annotations

import ctypes
import io
import struct
from dataclasses import dataclass, field
from typing import BinaryIO, Callable, Dict, List, Optional, Sequence

import numpy as np

from .configuration import ps_step_config

MAGIC = 0x4A545350        # "PSTJ"
FRAME_MAGIC = 0x4D415246  # "FRAM"
VERSION = 1
HAS_CANDIDATES = 1
HAS_CONTACTS = 2
PRODUCER_UNKNOWN, PRODUCER_REFERENCE_DOTNET, PRODUCER_ORACLE, PRODUCER_CUDA = 0, 1, 2, 3
PRODUCER_NAMES = {0: "unknown", 1: "reference (.NET)", 2: "oracle (CPU restatement)", 3: "CUDA (libps_b200)"}

_HEAD = struct.Struct("<10I")
_SCENE_BYTES = 64
_CONST_FIELDS = (("shape", np.int32, 1), ("dims", np.float64, 3), ("mass", np.float64, 1), ("inertia_inv", np.float64, 3),
                 ("force", np.float64, 3), ("torque", np.float64, 3), ("elasticity", np.float64, 1),
                 ("mu_kinetic", np.float64, 1), ("mu_static", np.float64, 1))
_STATE_FIELDS = (("pos", 3), ("vel", 3), ("rot", 9), ("ang_mom", 3))
_EMPTY_CONTACTS = lambda: {"pairs": np.zeros((0, 2), np.int32), "first": np.zeros(1, np.int32), "feature": np.zeros(0, np.uint32),  # noqa: E731
                           "normal": np.zeros((0, 3)), "position": np.zeros((0, 3)), "penetration": np.zeros(0)}


class SnapshotError(ValueError):
    """malformed file (ArgumentException class of the reference's conventions)"""


@dataclass
class Frame:
    step: int
    state: Dict[str, np.ndarray]                      # pos (n,3) vel (n,3) rot (n,9) ang_mom (n,3), all worlds
    candidates: Optional[List[np.ndarray]] = None     # per world (k,2) int32, ordered
    contacts: Optional[List[Dict[str, np.ndarray]]] = None  # per world: pairs, first, feature, normal, position, penetration


@dataclass
class Trajectory:
    config: ps_step_config
    world_offset: np.ndarray
    constants: Dict[str, np.ndarray]
    frames: List[Frame] = field(default_factory=list)
    flags: int = 0
    producer: int = PRODUCER_UNKNOWN
    scene: str = ""

    @property
    def n_worlds(self) -> int:
        return len(self.world_offset) - 1

    @property
    def n_bodies(self) -> int:
        return int(self.world_offset[-1])

    def bodies(self, frame: int = 0) -> dict:
        """SoA dict in the layout of prototypes.build_world / ps_bodies_view, with the state of `frame`"""
        d = {k: v.copy() for k, v in self.constants.items()}
        d.update({k: v.copy() for k, v in self.frames[frame].state.items()})
        return d

    def world_bodies(self, world: int, frame: int = 0) -> dict:
        a, b = int(self.world_offset[world]), int(self.world_offset[world + 1])
        return {k: v[a:b].copy() for k, v in self.bodies(frame).items()}


# ------------------------------------------------------------------------------------------------------------------
def _w(f: BinaryIO, a, dtype):
    f.write(np.ascontiguousarray(a, dtype=dtype).tobytes())


def _r(f: BinaryIO, dtype, count: int, shape=None):
    nbytes = np.dtype(dtype).itemsize * count
    b = f.read(nbytes)
    if len(b) != nbytes:
        raise SnapshotError("truncated snapshot")
    a = np.frombuffer(b, dtype=dtype).copy()
    return a.reshape(shape) if shape is not None else a


def write(f: BinaryIO, t: Trajectory) -> None:
    n, nw = t.n_bodies, t.n_worlds
    cfg_bytes = bytes(t.config)
    header_bytes = _HEAD.size + len(cfg_bytes) + _SCENE_BYTES
    f.write(_HEAD.pack(MAGIC, VERSION, header_bytes, t.flags, t.producer, nw, n, len(t.frames), len(cfg_bytes), 0))
    f.write(cfg_bytes)
    f.write(t.scene.encode("utf-8")[:_SCENE_BYTES - 1].ljust(_SCENE_BYTES, b"\0"))
    _w(f, t.world_offset, np.int32)
    for name, dt, width in _CONST_FIELDS:
        a = np.asarray(t.constants[name])
        if a.size != n * width:
            raise SnapshotError(f"constant {name}: {a.size} values for {n} bodies")
        _w(f, a, dt)
    for fr in t.frames:
        f.write(struct.pack("<Ii", FRAME_MAGIC, fr.step))
        for name, width in _STATE_FIELDS:
            a = np.asarray(fr.state[name])
            if a.size != n * width:
                raise SnapshotError(f"frame {fr.step}, {name}: {a.size} values for {n} bodies")
            _w(f, a, np.float64)
        if t.flags & HAS_CANDIDATES:
            for w in range(nw):
                c = np.zeros((0, 2), np.int32) if fr.candidates is None else np.asarray(fr.candidates[w], dtype=np.int32).reshape(-1, 2)
                f.write(struct.pack("<i", len(c)))
                _w(f, c, np.int32)
        if t.flags & HAS_CONTACTS:
            for w in range(nw):
                c = _EMPTY_CONTACTS() if fr.contacts is None else fr.contacts[w]
                npairs, nc = len(c["pairs"]), len(c["feature"])
                f.write(struct.pack("<ii", npairs, nc))
                _w(f, c["pairs"], np.int32)
                _w(f, c["first"], np.int32)
                _w(f, c["feature"], np.uint32)
                _w(f, c["normal"], np.float64)
                _w(f, c["position"], np.float64)
                _w(f, c["penetration"], np.float64)


def read(f: BinaryIO) -> Trajectory:
    h = f.read(_HEAD.size)
    if len(h) != _HEAD.size:
        raise SnapshotError("truncated snapshot header")
    magic, version, header_bytes, flags, producer, nw, n, nframes, cfg_bytes, _ = _HEAD.unpack(h)
    if magic != MAGIC:
        raise SnapshotError("not a PSTJ snapshot")
    if version != VERSION:
        raise SnapshotError(f"PSTJ version {version}, this reader knows {VERSION}")
    rest = f.read(header_bytes - _HEAD.size)
    if len(rest) != header_bytes - _HEAD.size or cfg_bytes > len(rest):
        raise SnapshotError("truncated snapshot header")
    cfg = ps_step_config()
    ctypes.memmove(ctypes.byref(cfg), rest[:cfg_bytes], min(cfg_bytes, ctypes.sizeof(cfg)))
    scene = rest[cfg_bytes:cfg_bytes + _SCENE_BYTES].split(b"\0", 1)[0].decode("utf-8", "replace")
    off = _r(f, np.int32, nw + 1)
    if int(off[-1]) != n or int(off[0]) != 0 or np.any(np.diff(off) < 0):
        raise SnapshotError("world_offset does not describe n_bodies")
    consts = {}
    for name, dt, width in _CONST_FIELDS:
        consts[name] = _r(f, dt, n * width, (n, width) if width > 1 else None)
    t = Trajectory(cfg, off, consts, [], flags, producer, scene)
    for _ in range(nframes):
        fh = f.read(8)
        if len(fh) != 8:
            raise SnapshotError("truncated frame header")
        fm, step = struct.unpack("<Ii", fh)
        if fm != FRAME_MAGIC:
            raise SnapshotError("frame marker missing: file is corrupt or was written with different flags")
        st = {name: _r(f, np.float64, n * width, (n, width)) for name, width in _STATE_FIELDS}
        cands = conts = None
        if flags & HAS_CANDIDATES:
            cands = []
            for _w_ in range(nw):
                (k,) = struct.unpack("<i", f.read(4))
                cands.append(_r(f, np.int32, 2 * k, (k, 2)))
        if flags & HAS_CONTACTS:
            conts = []
            for _w_ in range(nw):
                npairs, nc = struct.unpack("<ii", f.read(8))
                conts.append({"pairs": _r(f, np.int32, 2 * npairs, (npairs, 2)), "first": _r(f, np.int32, npairs + 1),
                              "feature": _r(f, np.uint32, nc), "normal": _r(f, np.float64, 3 * nc, (nc, 3)),
                              "position": _r(f, np.float64, 3 * nc, (nc, 3)), "penetration": _r(f, np.float64, nc)})
        t.frames.append(Frame(step, st, cands, conts))
    return t


def save(path: str, t: Trajectory) -> None:
    with open(path, "wb") as f:
        write(f, t)


def load(path: str) -> Trajectory:
    with open(path, "rb") as f:
        return read(f)


def dumps(t: Trajectory) -> bytes:
    b = io.BytesIO()
    write(b, t)
    return b.getvalue()


# ------------------------------------------------------------------------------------------------------------------
def record(bodies: dict, world_offset: Sequence[int], config: ps_step_config, marks: Sequence[int],
           step: Callable[[int], None], get_state: Callable[[], dict],
           get_contacts: Optional[Callable[[int], dict]] = None, get_candidates: Optional[Callable[[int], np.ndarray]] = None,
           producer: int = PRODUCER_UNKNOWN, scene: str = "") -> Trajectory:
    """Steps any implementation of the path (`step(n)`, `get_state()`, optional per-world `get_contacts(w)` /
    `get_candidates(w)`) and records a frame at step 0 and at every mark.  `bodies`: SoA dict of all worlds."""
    off = np.asarray(world_offset, dtype=np.int32)
    consts = {name: np.array(bodies[name], dtype=dt) for name, dt, _ in _CONST_FIELDS}
    flags = (HAS_CONTACTS if get_contacts else 0) | (HAS_CANDIDATES if get_candidates else 0)
    t = Trajectory(config, off, consts, [], flags, producer, scene)
    nw = len(off) - 1
    keys = [k for k, _ in _STATE_FIELDS]

    def frame(s, first):
        st = get_state()
        fr = Frame(s, {k: np.array(st[k], dtype=np.float64) for k in keys})
        if get_candidates:
            fr.candidates = [np.zeros((0, 2), np.int32) if first else np.asarray(get_candidates(w), dtype=np.int32) for w in range(nw)]
        if get_contacts:
            fr.contacts = [_EMPTY_CONTACTS() if first else {k: np.asarray(v) for k, v in get_contacts(w).items()} for w in range(nw)]
        return fr

    t.frames.append(frame(0, True))
    done = 0
    for m in marks:
        if m <= done:
            raise SnapshotError("marks must be increasing and positive")
        step(m - done)
        done = m
        t.frames.append(frame(m, False))
    return t


# ------------------------------------------------------------------------------------------------------------------
def _canon_bits(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = a.view(np.uint64).copy()
    b[np.isnan(a)] = np.uint64(0x7FF8000000000000)  # the reference compares NaN with nothing (Appendix B Q20)
    return b


@dataclass
class FrameDiff:
    step: int
    bit_exact: bool
    max_rel: Dict[str, float]       # per state field, over dynamic bodies: |a-b| / max(1, |a|, |b|)
    candidates_equal: Optional[bool]
    pairs_equal: Optional[bool]     # colliding-pair sets (order included)
    features_equal: Optional[bool]  # contact counts and feature ids
    contact_max_abs: Optional[float]


def compare(a: Trajectory, b: Trajectory, static_ang_mom: bool = False) -> List[FrameDiff]:
    """Frame-by-frame comparison of two trajectories of the same scene (frames matched by step number).
    States are compared over DYNAMIC bodies; a static body's angular momentum is a meaningless accumulator in the
    reference (Appendix B Q2) and only compared when `static_ang_mom`."""
    if a.n_bodies != b.n_bodies or not np.array_equal(a.world_offset, b.world_offset):
        raise SnapshotError("trajectories describe different batches")
    for name, _, _ in _CONST_FIELDS:
        if not np.array_equal(_canon_bits(a.constants[name]) if a.constants[name].dtype == np.float64 else a.constants[name],
                              _canon_bits(b.constants[name]) if b.constants[name].dtype == np.float64 else b.constants[name]):
            raise SnapshotError(f"trajectories differ in the body constant {name}")
    dyn = ~np.isinf(a.constants["mass"])
    fb = {f.step: f for f in b.frames}
    out = []
    for fa in a.frames:
        if fa.step not in fb:
            continue
        fbb = fb[fa.step]
        exact, rel = True, {}
        for k, _ in _STATE_FIELDS:
            m = np.ones(len(dyn), bool) if (k == "ang_mom" and static_ang_mom) else dyn
            x, y = fa.state[k][m], fbb.state[k][m]
            exact = exact and np.array_equal(_canon_bits(x), _canon_bits(y))
            both_nan = np.isnan(x) & np.isnan(y)
            with np.errstate(invalid="ignore"):
                d = np.abs(x - y) / np.maximum(1.0, np.maximum(np.abs(x), np.abs(y)))
            d = np.where(both_nan, 0.0, d)
            d = np.where(np.isnan(d), np.inf, d)
            rel[k] = float(d.max()) if d.size else 0.0
        ce = pe = fe = cm = None
        if fa.candidates is not None and fbb.candidates is not None:
            ce = all(np.array_equal(p, q) for p, q in zip(fa.candidates, fbb.candidates))
        if fa.contacts is not None and fbb.contacts is not None:
            pe = all(np.array_equal(p["pairs"], q["pairs"]) for p, q in zip(fa.contacts, fbb.contacts))
            counts = pe and all(np.array_equal(p["first"], q["first"]) for p, q in zip(fa.contacts, fbb.contacts))
            if PRODUCER_REFERENCE_DOTNET in (a.producer, b.producer):
                fe = None if counts else False  # the reference stores no feature ids: contact counts only
            else:
                fe = counts and all(np.array_equal(p["feature"], q["feature"]) for p, q in zip(fa.contacts, fbb.contacts))
            if counts:
                cm = 0.0
                for p, q in zip(fa.contacts, fbb.contacts):
                    for k in ("normal", "position", "penetration"):
                        if p[k].size:
                            cm = max(cm, float(np.nanmax(np.abs(p[k] - q[k]))) if not np.all(np.isnan(p[k] - q[k])) else 0.0)
        out.append(FrameDiff(fa.step, exact, rel, ce, pe, fe, cm))
    return out
