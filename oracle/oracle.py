"""ctypes wrapper of oracle/libps_oracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) import this.
The product package never does."""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_char_p, c_double, c_int, c_int32, c_int64, c_uint32, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libps_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("ps_oracle.cpp", "ref_math.hpp", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        L = _LIB
        L.orc_world_create.restype = c_void_p
        L.orc_tree_create.restype = c_void_p
        L.orc_overlap_pairs.restype = c_int64
        for name in ("orc_world_destroy", "orc_world_set_joints", "orc_world_step", "orc_world_get_state",
                     "orc_world_set_state", "orc_world_apply_impulse", "orc_world_stats", "orc_batch_step",
                     "orc_world_integrate_only", "orc_tree_destroy", "orc_tree_set_extent",
                     "orc_tree_set_root_leaf", "orc_tree_set_root_children", "orc_tree_remove", "orc_aabbs"):
            getattr(L, name).restype = None
    return _LIB


def _p(a, ct=c_double):
    return a.ctypes.data_as(POINTER(ct))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class OracleWorld:
    """One world stepped by the CPU restatement.  `bodies` is the SoA dict of
    physicssimulator_b200.prototypes.build_world; `cfg` a ps_step_config ctypes struct."""

    def __init__(self, bodies: dict, cfg, joints=None):
        L = lib()
        self.n = len(bodies["mass"])
        b = {k: np.ascontiguousarray(v) for k, v in bodies.items()}
        self._cfg = cfg
        self.h = c_void_p(L.orc_world_create(
            c_int(self.n), _p(b["shape"].astype(np.int32), c_int32), _p(_f64(b["dims"])), _p(_f64(b["mass"])),
            _p(_f64(b["inertia_inv"])), _p(_f64(b["pos"])), _p(_f64(b["vel"])), _p(_f64(b["rot"])),
            _p(_f64(b["ang_mom"])), _p(_f64(b["force"])), _p(_f64(b["torque"])), _p(_f64(b["elasticity"])),
            _p(_f64(b["mu_kinetic"])), _p(_f64(b["mu_static"])), ctypes.byref(cfg)))
        if joints:
            a = np.array([j[0] for j in joints], dtype=np.int32)
            bb = np.array([j[1] for j in joints], dtype=np.int32)
            la = _f64([j[2] for j in joints])
            lb = _f64([j[3] for j in joints])
            L.orc_world_set_joints(self.h, c_int(len(joints)), _p(a, c_int32), _p(bb, c_int32), _p(la), _p(lb))

    def __del__(self):
        try:
            if self.h:
                lib().orc_world_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def use_spatial_tree(self, leaf_cap, max_depth, mn, mx, aabb_mode=0) -> int:
        return lib().orc_world_use_spatial_tree(self.h, c_int(leaf_cap), c_int(max_depth), _p(_f64(mn)), _p(_f64(mx)),
                                                c_int(aabb_mode))

    def step(self, n=1):
        lib().orc_world_step(self.h, c_int(n))

    def integrate_only(self, n=1):
        lib().orc_world_integrate_only(self.h, c_int(n))

    def num_bodies(self) -> int:
        return lib().orc_world_num_bodies(self.h)

    def get_state(self):
        n = self.num_bodies()
        pos, vel, rot, L = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 9)), np.zeros((n, 3))
        lib().orc_world_get_state(self.h, _p(pos), _p(vel), _p(rot), _p(L))
        return {"pos": pos, "vel": vel, "rot": rot, "ang_mom": L}

    def set_state(self, st):
        lib().orc_world_set_state(self.h, _p(_f64(st["pos"])), _p(_f64(st["vel"])), _p(_f64(st["rot"])), _p(_f64(st["ang_mom"])))

    def apply_impulse(self, body, J, off):
        lib().orc_world_apply_impulse(self.h, c_int(body), _p(_f64(J)), _p(_f64(off)))

    def add_body(self, one: dict) -> int:
        g = lambda k: _f64(one[k]).reshape(-1)
        return lib().orc_world_add_body(
            self.h, c_int(int(one["shape"][0])), _p(g("dims")), c_double(float(one["mass"][0])), _p(g("inertia_inv")),
            _p(g("pos")), _p(g("vel")), _p(g("rot")), _p(g("ang_mom")), _p(g("force")), _p(g("torque")),
            c_double(float(one["elasticity"][0])), c_double(float(one["mu_kinetic"][0])), c_double(float(one["mu_static"][0])))

    def candidates(self):
        cap = self.num_bodies() * self.num_bodies()
        ids = np.zeros((max(cap, 1), 2), dtype=np.int32)
        n = lib().orc_world_get_candidates(self.h, _p(ids, c_int32), c_int(cap))
        return ids[:n].copy()

    def contacts(self, pair_cap=4096, contact_cap=65536):
        ids = np.zeros((pair_cap, 2), dtype=np.int32)
        first = np.zeros(pair_cap + 1, dtype=np.int32)
        nrm, pos = np.zeros((contact_cap, 3)), np.zeros((contact_cap, 3))
        pen = np.zeros(contact_cap)
        feat = np.zeros(contact_cap, dtype=np.uint32)
        npairs = c_int32(0)
        rc = lib().orc_world_get_contacts(self.h, ctypes.byref(npairs), _p(ids, c_int32), _p(first, c_int32), _p(nrm), _p(pos),
                                          _p(pen), _p(feat, c_uint32), c_int(pair_cap), c_int(contact_cap))
        assert rc == 0
        n = npairs.value
        nc = int(first[n])
        return {"pairs": ids[:n].copy(), "first": first[:n + 1].copy(), "normal": nrm[:nc].copy(),
                "position": pos[:nc].copy(), "penetration": pen[:nc].copy(), "feature": feat[:nc].copy()}

    def detect(self, i, j, cap=32):
        nrm, pos, pen = np.zeros((cap, 3)), np.zeros((cap, 3)), np.zeros(cap)
        feat = np.zeros(cap, dtype=np.uint32)
        n = lib().orc_world_detect(self.h, c_int(i), c_int(j), _p(nrm), _p(pos), _p(pen), _p(feat, c_uint32), c_int(cap))
        if n < 0:
            return None
        return {"normal": nrm[:n].copy(), "position": pos[:n].copy(), "penetration": pen[:n].copy(), "feature": feat[:n].copy()}

    def stats(self):
        t, h, c, e = c_int64(0), c_int64(0), c_int64(0), c_int64(0)
        mp, nan = c_double(0), c_int32(0)
        lib().orc_world_stats(self.h, ctypes.byref(t), ctypes.byref(h), ctypes.byref(c), ctypes.byref(e), ctypes.byref(mp), ctypes.byref(nan))
        return {"pairs_tested": t.value, "pairs_colliding": h.value, "contacts": c.value, "errors": e.value,
                "max_penetration": mp.value, "nan": nan.value}


def batch_step(worlds, nsteps: int, nthreads: int):
    arr = (c_void_p * len(worlds))(*[w.h for w in worlds])
    lib().orc_batch_step(arr, c_int(len(worlds)), c_int(nsteps), c_int(nthreads))


def clip(eps, plane_n, plane_d, poly, cap=64):
    out = np.zeros((cap, 3))
    pn, pd, pl = _f64(plane_n), _f64(plane_d), _f64(poly)
    n = lib().orc_clip(c_double(eps), c_int(len(pd)), _p(pn), _p(pd), c_int(len(pl)), _p(pl), _p(out), c_int(cap))
    return out[:n].copy()


def overlap_pairs(b: dict, aabb_mode=0, method=1, cap=None):
    n = len(b["mass"])
    cap = cap or max(16 * n, 1024)
    out = np.zeros((cap, 2), dtype=np.int32)
    npairs = lib().orc_overlap_pairs(c_int(n), _p(b["shape"].astype(np.int32), c_int32), _p(_f64(b["dims"])), _p(_f64(b["mass"])),
                                     _p(_f64(b["pos"])), _p(_f64(b["rot"])), c_int(aabb_mode), c_int(method),
                                     _p(out, c_int32), c_int64(cap))
    assert npairs <= cap, "overlap pair capacity"
    return out[:npairs].copy()


def aabbs(b: dict, aabb_mode=0):
    n = len(b["shape"])
    size, mn = np.zeros((n, 3)), np.zeros((n, 3))
    lib().orc_aabbs(c_int(n), _p(b["shape"].astype(np.int32), c_int32), _p(_f64(b["dims"])), _p(_f64(b["pos"])),
                    _p(_f64(b["rot"])), c_int(aabb_mode), _p(size), _p(mn))
    return size, mn


class OracleTree:
    """SpatialTree restatement, for the reference's known-answer facts."""

    def __init__(self, max_leaf, max_depth, bounds):
        bb = _f64(bounds).reshape(-1)
        self.ndim = len(bb) // 2
        h = lib().orc_tree_create(c_int(max_leaf), c_int(max_depth), c_int(self.ndim), _p(bb))
        if not h:
            raise ValueError("SpatialTree.init: invalid argument")
        self.h = c_void_p(h)

    def __del__(self):
        try:
            if self.h:
                lib().orc_tree_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_extent(self, oid, extents):
        """extents: list of (size, position) per dimension"""
        lib().orc_tree_set_extent(self.h, c_int(oid), _p(_f64(extents).reshape(-1)))

    def set_root_leaf(self, ids):
        a = np.array(list(ids), dtype=np.int32)
        lib().orc_tree_set_root_leaf(self.h, c_int(len(a)), _p(a, c_int32))

    def set_root_children(self, children):
        counts = np.array([len(c) for c in children], dtype=np.int32)
        ids = np.array([i for c in children for i in c], dtype=np.int32)
        lib().orc_tree_set_root_children(self.h, c_int(len(children)), _p(counts, c_int32), _p(ids, c_int32))

    def insert(self, oid) -> int:
        return lib().orc_tree_insert(self.h, c_int(oid))

    def remove(self, oid):
        lib().orc_tree_remove(self.h, c_int(oid))

    def dump(self) -> str:
        buf = ctypes.create_string_buffer(1 << 16)
        n = lib().orc_tree_dump(self.h, buf, c_int(len(buf)))
        assert n >= 0
        return buf.value.decode()

    def num_buckets(self) -> int:
        return lib().orc_tree_num_buckets(self.h)
