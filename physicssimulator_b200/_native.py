"""ctypes binding of libps_b200.so (include/ps_b200.h).  There is no fallback: if the CUDA library is
missing or no device is present, every call fails loudly."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int32, c_int64, c_uint32, c_void_p

import numpy as np

from .configuration import ps_step_config

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libps_b200.so")

PS_OK = 0
PS_ERR_INVALID_ARGUMENT = -1
PS_ERR_OUT_OF_RANGE = -2
PS_ERR_CUDA = -3
PS_ERR_NAN = -4
PS_ERR_INVALID_OPERATION = -5
PS_ERR_CAPACITY = -6

EXPORTS = (
    "ps_init", "ps_last_error", "ps_version", "ps_batch_create", "ps_batch_destroy", "ps_batch_step",
    "ps_batch_step_async", "ps_batch_synchronize", "ps_batch_get_state", "ps_batch_set_state",
    "ps_batch_get_state_device", "ps_batch_get_contacts", "ps_batch_get_candidates", "ps_batch_apply_impulse",
    "ps_batch_add_body", "ps_batch_reset", "ps_batch_stats", "ps_batch_num_bodies", "ps_broadphase_pairs",
    "ps_batch_debug_cycles", "ps_batch_schedule", "ps_batch_step_host",
)


class ps_bodies_view(ctypes.Structure):
    _fields_ = [("n_worlds", c_int32), ("world_offset", POINTER(c_int32)), ("shape", POINTER(c_int32)),
                ("dims", POINTER(c_double)), ("mass", POINTER(c_double)), ("inertia_inv", POINTER(c_double)),
                ("pos", POINTER(c_double)), ("vel", POINTER(c_double)), ("rot", POINTER(c_double)),
                ("ang_mom", POINTER(c_double)), ("force", POINTER(c_double)), ("torque", POINTER(c_double)),
                ("elasticity", POINTER(c_double)), ("mu_kinetic", POINTER(c_double)), ("mu_static", POINTER(c_double))]


class ps_joints_view(ctypes.Structure):
    _fields_ = [("n_joints", c_int32), ("world", POINTER(c_int32)), ("body_a", POINTER(c_int32)),
                ("body_b", POINTER(c_int32)), ("anchor_a", POINTER(c_double)), ("anchor_b", POINTER(c_double))]


class ps_stats(ctypes.Structure):
    _fields_ = [("steps", c_int64), ("body_steps", c_int64), ("pairs_tested", c_int64), ("pairs_colliding", c_int64),
                ("contacts", c_int64), ("fallback_steps", c_int64), ("nan_worlds", c_int64), ("contact_overflow", c_int64),
                ("reference_exceptions", c_int64), ("max_penetration", c_double),
                ("fallback_margin_steps", c_int64), ("fallback_capacity_steps", c_int64), ("fallback_static_steps", c_int64),
                ("margin_verified_steps", c_int64)]


class NativeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"ps_b200 error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    """Load libps_b200.so; raises if it was not built (run __graft_entry__.build())."""
    global _lib
    if _lib is None:
        path = os.environ.get("PS_B200_LIB", LIB_PATH)  # tuning: another BUILD of the same CUDA library (tools/variants)
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: the CUDA extension was not built and there is no CPU path. "
                               "Run `python -c 'import __graft_entry__ as g; g.build()'`.")
        L = ctypes.CDLL(path)
        L.ps_last_error.restype = c_char_p
        _lib = L
    return _lib


def check(rc: int):
    """Map the native error classes back to the reference's exception types (SURVEY.md §8(b) Errors)."""
    if rc == PS_OK:
        return
    msg = lib().ps_last_error().decode(errors="replace")
    if rc == PS_ERR_INVALID_ARGUMENT:
        raise ValueError(msg)  # ArgumentException
    if rc == PS_ERR_OUT_OF_RANGE:
        raise KeyError(msg)  # KeyNotFoundException
    if rc == PS_ERR_INVALID_OPERATION:
        raise RuntimeError(msg)  # InvalidOperationException
    raise NativeError(rc, msg)


def dptr(a: np.ndarray, ct=c_double):
    return a.ctypes.data_as(POINTER(ct))


def make_view(bodies: dict, world_offset: np.ndarray, keep: list) -> ps_bodies_view:
    """bodies: SoA dict (prototypes.BODY_FIELDS).  `keep` receives the contiguous arrays (lifetime)."""
    def f64(k):
        a = np.ascontiguousarray(bodies[k], dtype=np.float64)
        keep.append(a)
        return dptr(a)
    off = np.ascontiguousarray(world_offset, dtype=np.int32)
    shape = np.ascontiguousarray(bodies["shape"], dtype=np.int32)
    keep += [off, shape]
    return ps_bodies_view(n_worlds=len(off) - 1, world_offset=dptr(off, c_int32), shape=dptr(shape, c_int32), dims=f64("dims"),
                          mass=f64("mass"), inertia_inv=f64("inertia_inv"), pos=f64("pos"), vel=f64("vel"), rot=f64("rot"),
                          ang_mom=f64("ang_mom"), force=f64("force"), torque=f64("torque"), elasticity=f64("elasticity"),
                          mu_kinetic=f64("mu_kinetic"), mu_static=f64("mu_static"))


class Batch:
    """Thin RAII wrapper of a native batch handle."""

    def __init__(self, bodies: dict, world_offset, config: ps_step_config, joints=None, device: int = -1):
        L = lib()
        check(L.ps_init(c_int32(device)))
        keep = []
        view = make_view(bodies, world_offset, keep)
        jv = None
        if joints:
            w = np.array([j[0] for j in joints], dtype=np.int32)
            a = np.array([j[1] for j in joints], dtype=np.int32)
            b = np.array([j[2] for j in joints], dtype=np.int32)
            la = np.ascontiguousarray([j[3] for j in joints], dtype=np.float64)
            lb = np.ascontiguousarray([j[4] for j in joints], dtype=np.float64)
            keep += [w, a, b, la, lb]
            jv = ps_joints_view(len(joints), dptr(w, c_int32), dptr(a, c_int32), dptr(b, c_int32), dptr(la), dptr(lb))
        self.h = c_void_p()
        check(L.ps_batch_create(ctypes.byref(self.h), ctypes.byref(view), ctypes.byref(jv) if jv else None, ctypes.byref(config)))
        self.world_offset = np.array(world_offset, dtype=np.int32)
        self.n_worlds = len(self.world_offset) - 1

    def close(self):
        if getattr(self, "h", None):
            lib().ps_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _refresh_offsets_after_add(self, world):
        self.world_offset[world + 1:] += 1

    def step(self, n=1):
        check(lib().ps_batch_step(self.h, c_int32(n)))

    def step_async(self, n=1, stream=0):
        check(lib().ps_batch_step_async(self.h, c_int32(n), c_void_p(stream)))

    def synchronize(self):
        check(lib().ps_batch_synchronize(self.h))

    def get_state(self, w0=0, w1=None, out=None):
        w1 = self.n_worlds if w1 is None else w1
        if not (0 <= w0 <= w1 <= self.n_worlds):  # let the native range check speak (KeyNotFoundException)
            check(lib().ps_batch_get_state(self.h, c_int32(w0), c_int32(w1), None, None, None, None))
        n = int(self.world_offset[w1] - self.world_offset[w0])
        if out is None:
            out = {"pos": np.empty((n, 3)), "vel": np.empty((n, 3)), "rot": np.empty((n, 9)), "ang_mom": np.empty((n, 3))}
        check(lib().ps_batch_get_state(self.h, c_int32(w0), c_int32(w1), dptr(out["pos"]), dptr(out["vel"]), dptr(out["rot"]), dptr(out["ang_mom"])))
        return out

    def set_state(self, st: dict, w0=0, w1=None):
        w1 = self.n_worlds if w1 is None else w1
        a = [np.ascontiguousarray(st[k], dtype=np.float64) for k in ("pos", "vel", "rot", "ang_mom")]
        check(lib().ps_batch_set_state(self.h, c_int32(w0), c_int32(w1), dptr(a[0]), dptr(a[1]), dptr(a[2]), dptr(a[3])))

    def step_host(self, st: dict, n=1, out=None):
        """set_state(all) + step(n) + get_state(all) in one call, copies pipelined against the step over chunks of
        worlds (ps_batch_step_host).  `st`/`out`: dicts of C-contiguous float64 arrays (pinned memory makes the copies
        asynchronous); out may be st."""
        a = [np.ascontiguousarray(st[k], dtype=np.float64) for k in ("pos", "vel", "rot", "ang_mom")]
        if out is None:
            n_b = int(self.world_offset[-1])
            out = {"pos": np.empty((n_b, 3)), "vel": np.empty((n_b, 3)), "rot": np.empty((n_b, 9)), "ang_mom": np.empty((n_b, 3))}
        check(lib().ps_batch_step_host(self.h, c_int32(n), dptr(a[0]), dptr(a[1]), dptr(a[2]), dptr(a[3]),
                                       dptr(out["pos"]), dptr(out["vel"]), dptr(out["rot"]), dptr(out["ang_mom"])))
        return out

    def get_state_device(self, d_pos, d_vel, d_rot, d_L, w0=0, w1=None, stream=0):
        """raw device pointers (ints), e.g. torch tensors' data_ptr()"""
        w1 = self.n_worlds if w1 is None else w1
        check(lib().ps_batch_get_state_device(self.h, c_int32(w0), c_int32(w1), c_void_p(d_pos), c_void_p(d_vel), c_void_p(d_rot), c_void_p(d_L), c_void_p(stream)))

    def contacts(self, world: int, pair_cap=8192, contact_cap=65536):
        ids = np.zeros((pair_cap, 2), dtype=np.int32)
        first = np.zeros(pair_cap + 1, dtype=np.int32)
        nrm, pos, pen = np.zeros((contact_cap, 3)), np.zeros((contact_cap, 3)), np.zeros(contact_cap)
        feat = np.zeros(contact_cap, dtype=np.uint32)
        n = c_int32(0)
        check(lib().ps_batch_get_contacts(self.h, c_int32(world), ctypes.byref(n), dptr(ids, c_int32), dptr(first, c_int32), dptr(nrm), dptr(pos),
                                          dptr(pen), dptr(feat, c_uint32), c_int32(pair_cap), c_int32(contact_cap)))
        n = n.value
        nc = int(first[n])
        return {"pairs": ids[:n].copy(), "first": first[:n + 1].copy(), "normal": nrm[:nc].copy(), "position": pos[:nc].copy(),
                "penetration": pen[:nc].copy(), "feature": feat[:nc].copy()}

    def candidates(self, world: int):
        n = c_int32(0)
        check(lib().ps_batch_get_candidates(self.h, c_int32(world), ctypes.byref(n), None, c_int32(0)))
        ids = np.zeros((max(n.value, 1), 2), dtype=np.int32)
        check(lib().ps_batch_get_candidates(self.h, c_int32(world), ctypes.byref(n), dptr(ids, c_int32), c_int32(n.value)))
        return ids[:n.value].copy()

    def apply_impulse(self, world, body, impulse, offset):
        J = (c_double * 3)(*[float(x) for x in impulse])
        O = (c_double * 3)(*[float(x) for x in offset])
        check(lib().ps_batch_apply_impulse(self.h, c_int32(world), c_int32(body), J, O))

    def add_body(self, world: int, one: dict):
        keep = []
        view = make_view(one, np.array([0, 1], dtype=np.int32), keep)
        check(lib().ps_batch_add_body(self.h, c_int32(world), ctypes.byref(view)))
        self._refresh_offsets_after_add(world)

    def reset(self, world_offset):
        check(lib().ps_batch_reset(self.h))
        self.world_offset = np.array(world_offset, dtype=np.int32)

    def stats(self) -> dict:
        s = ps_stats()
        check(lib().ps_batch_stats(self.h, ctypes.byref(s)))
        return {k: getattr(s, k) for k, _ in ps_stats._fields_}

    def schedule(self):
        p, l, k = c_int32(0), c_int32(0), c_int64(0)
        check(lib().ps_batch_schedule(self.h, ctypes.byref(p), ctypes.byref(l), ctypes.byref(k)))
        return {"path": "wide" if p.value else "fused", "lanes_per_world": l.value, "kernel_launches": k.value}

    def debug_cycles(self):
        out = np.zeros((self.n_worlds, 8), dtype=np.uint64)
        check(lib().ps_batch_debug_cycles(self.h, out.ctypes.data_as(POINTER(ctypes.c_uint64))))
        return out

    def num_bodies(self):
        a, b = c_int32(0), c_int32(0)
        check(lib().ps_batch_num_bodies(self.h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value


def broadphase_pairs(b: dict, aabb_mode=0, capacity=None):
    """Grid broad phase on one large world (config 5). Returns (pairs[n,2] sorted, kernel_ms)."""
    n = len(b["mass"])
    capacity = capacity or max(32 * n, 4096)
    out = np.zeros((capacity, 2), dtype=np.int32)
    shape = np.ascontiguousarray(b["shape"], dtype=np.int32)
    arrs = [np.ascontiguousarray(b[k], dtype=np.float64) for k in ("dims", "mass", "pos", "rot")]
    npairs = c_int64(0)
    ms = c_double(0)
    check(lib().ps_broadphase_pairs(c_int32(n), dptr(shape, c_int32), dptr(arrs[0]), dptr(arrs[1]), dptr(arrs[2]), dptr(arrs[3]),
                                    c_int32(aabb_mode), ctypes.byref(npairs), dptr(out, c_int32), c_int64(capacity), ctypes.byref(ms)))
    return out[:npairs.value].copy(), ms.value
