"""Shared helpers of the test-suite (not product code)."""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from physicssimulator_b200 import prototypes as P  # noqa: E402
from physicssimulator_b200 import scenes  # noqa: E402
from physicssimulator_b200.configuration import Configuration, StepConfiguration, to_native_config  # noqa: E402


def bits(a):
    """bit pattern with every NaN mapped to one canonical pattern: x86 and the GPU disagree on the sign/payload
    of a freshly generated NaN (0xFFF8.. vs 0x7FF8..), and the reference compares NaN == NaN nowhere (Q20)"""
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = a.view(np.uint64).copy()
    b[np.isnan(a)] = np.uint64(0x7FF8000000000000)
    return b


def same_bits(a, b):
    """bit-exact equality (+0 and -0 differ); NaN == NaN"""
    return np.array_equal(bits(a), bits(b))


def assert_state_bits(a: dict, b: dict, what=""):
    for k in ("pos", "vel", "rot", "ang_mom"):
        if not same_bits(a[k], b[k]):
            bad = np.argwhere(bits(a[k]) != bits(b[k]))
            i = tuple(bad[0])
            raise AssertionError(f"{what}: field {k} differs at {i}: {a[k][i]!r} vs {b[k][i]!r} ({len(bad)} entries)")


def to_planar(bodies: dict):
    """per-world planar layout of ps_step_kernel.cuh: (par[16*N], dyn[18*N])"""
    n = len(bodies["mass"])
    par = np.zeros((16, n))
    with np.errstate(divide="ignore"):
        par[0] = 1.0 / bodies["mass"]
    par[1:4] = bodies["inertia_inv"].T
    par[4:7] = bodies["force"].T
    par[7:10] = bodies["torque"].T
    par[10] = bodies["elasticity"]
    par[11] = bodies["mu_kinetic"]
    par[12:15] = bodies["dims"].T
    par[15] = bodies["shape"]
    dyn = np.zeros((18, n))
    dyn[0:3] = bodies["pos"].T
    dyn[3:6] = bodies["vel"].T
    dyn[6:15] = bodies["rot"].T
    dyn[15:18] = bodies["ang_mom"].T
    return np.ascontiguousarray(par), np.ascontiguousarray(dyn)


def from_planar_dyn(dyn):
    return {"pos": np.ascontiguousarray(dyn[0:3].T), "vel": np.ascontiguousarray(dyn[3:6].T),
            "rot": np.ascontiguousarray(dyn[6:15].T), "ang_mom": np.ascontiguousarray(dyn[15:18].T)}


_HH = None


def host_harness():
    """g++ build of the device physics headers (tests/host_harness) — CPU check of the per-pair routines."""
    global _HH
    if _HH is None:
        d = os.path.join(ROOT, "tests", "host_harness")
        so = os.path.join(d, "libps_host_harness.so")
        src = os.path.join(d, "physics_host.cpp")
        deps = [src] + [os.path.join(ROOT, "physicssimulator_b200", "csrc", f) for f in ("ps_physics.cuh", "ps_math.cuh", "ps_sincos.h", "ps_octree.hpp", "ps_coop.cuh", "ps_team.cuh")]
        if not os.path.exists(so) or any(os.path.getmtime(x) > os.path.getmtime(so) for x in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-pthread", "-shared", "-x", "c++", src, "-o", so])
        _HH = ctypes.CDLL(so)
        _HH.hh_world_step_v.restype = None
        _HH.hh_tree_create.restype = ctypes.c_void_p
        for f in ("hh_tree_destroy", "hh_tree_set_extent", "hh_tree_set_root_leaf", "hh_tree_set_root_children", "hh_tree_remove"):
            getattr(_HH, f).restype = None
    return _HH


class ProductTree:
    """the product's host-side SpatialTree (physicssimulator_b200/csrc/ps_octree.hpp, compiled into the host harness)
    behind the interface of oracle.OracleTree"""

    def __init__(self, max_leaf, max_depth, bounds):
        bb = np.ascontiguousarray(np.array(bounds, dtype=np.float64).reshape(-1))
        self.ndim = len(bb) // 2
        h = host_harness().hh_tree_create(ctypes.c_int(max_leaf), ctypes.c_int(max_depth), ctypes.c_int(self.ndim),
                                          bb.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        if not h:
            raise ValueError("SpatialTree.init: invalid argument")
        self.h = ctypes.c_void_p(h)

    def __del__(self):
        try:
            if self.h:
                host_harness().hh_tree_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_extent(self, oid, extents):
        e = np.ascontiguousarray(np.array(extents, dtype=np.float64).reshape(-1))
        host_harness().hh_tree_set_extent(self.h, ctypes.c_int(oid), e.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))

    def set_root_leaf(self, ids):
        a = np.array(list(ids), dtype=np.int32)
        host_harness().hh_tree_set_root_leaf(self.h, ctypes.c_int(len(a)), a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))

    def set_root_children(self, children):
        counts = np.array([len(c) for c in children], dtype=np.int32)
        ids = np.array([i for c in children for i in c], dtype=np.int32)
        host_harness().hh_tree_set_root_children(self.h, ctypes.c_int(len(children)), counts.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                                 ids.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))

    def insert(self, oid) -> int:
        return host_harness().hh_tree_insert(self.h, ctypes.c_int(oid))

    def remove(self, oid):
        host_harness().hh_tree_remove(self.h, ctypes.c_int(oid))

    def dump(self) -> str:
        buf = ctypes.create_string_buffer(1 << 16)
        n = host_harness().hh_tree_dump(self.h, buf, ctypes.c_int(len(buf)))
        assert n >= 0
        return buf.value.decode()

    def num_buckets(self) -> int:
        return host_harness().hh_tree_num_buckets(self.h)

    def candidates(self, is_static):
        st = np.ascontiguousarray(np.array(is_static, dtype=np.uint8))
        out = np.zeros(1 << 20, dtype=np.uint32)
        n = host_harness().hh_tree_candidates(self.h, st.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)),
                                              out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), ctypes.c_int(len(out)))
        assert 0 <= n <= len(out)
        return [(int(k) & 0xffff, int(k) >> 16) for k in out[:n]]


def harness_step(bodies: dict, cfg_native, nsteps: int, cap_pairs=4096, cap_contacts=32768, fast=False, mode=None):
    hh = host_harness()
    par, dyn = to_planar(bodies)
    n = len(bodies["mass"])
    counters = np.zeros(8, dtype=np.int64)
    pairs = np.zeros((cap_pairs, 3), dtype=np.int32)
    feats = np.zeros(cap_contacts, dtype=np.uint32)
    cont = np.zeros((cap_contacts, 7))
    counts = np.zeros(2, dtype=np.int32)
    P_ = ctypes.POINTER
    hh.hh_world_step_v(ctypes.c_int(n), par.ctypes.data_as(P_(ctypes.c_double)), dyn.ctypes.data_as(P_(ctypes.c_double)),
                       ctypes.byref(cfg_native), ctypes.c_int(nsteps), counters.ctypes.data_as(P_(ctypes.c_int64)),
                       pairs.ctypes.data_as(P_(ctypes.c_int32)), feats.ctypes.data_as(P_(ctypes.c_uint32)),
                       cont.ctypes.data_as(P_(ctypes.c_double)), counts.ctypes.data_as(P_(ctypes.c_int32)),
                       ctypes.c_int(cap_pairs), ctypes.c_int(cap_contacts), ctypes.c_int(mode if mode is not None else (1 if fast else 0)))
    st = from_planar_dyn(dyn)
    npairs, nc = int(counts[0]), int(counts[1])
    first = np.concatenate([[0], np.cumsum(pairs[:npairs, 2])]).astype(np.int32)
    rec = {"pairs": pairs[:npairs, :2].copy(), "first": first, "normal": cont[:nc, 0:3].copy(), "position": cont[:nc, 3:6].copy(),
           "penetration": cont[:nc, 6].copy(), "feature": feats[:nc].copy()}
    return st, rec, {"pairs_tested": int(counters[0]), "pairs_colliding": int(counters[1]), "contacts": int(counters[2]),
                     "errors": int(counters[3]), "overflow": int(counters[4]),
                     "team_runs": int(counters[5]), "team_literal": int(counters[6]), "team_mismatch": int(counters[7])}


def assert_contacts_equal(a: dict, b: dict, what=""):
    assert np.array_equal(a["pairs"], b["pairs"]), f"{what}: colliding pair sets differ\n{a['pairs']}\n{b['pairs']}"
    assert np.array_equal(a["first"], b["first"]), f"{what}: contact counts differ"
    assert np.array_equal(a["feature"], b["feature"]), f"{what}: feature ids differ"
    for k in ("normal", "position", "penetration"):
        assert same_bits(a[k], b[k]), f"{what}: contact {k} differs"
