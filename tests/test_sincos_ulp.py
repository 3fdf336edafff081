"""csrc/ps_sincos.h is SHARED by the device path and the oracle (so that they can agree bit for bit); the reference calls
System.Math.Sin/Cos, i.e. the platform CRT.  This pins the shared routine against libm (numpy -> glibc): at most 1 ulp
apart over the angles a step can produce, and exactly equal on most of them — the claim DESIGN.md section 2 makes."""
import ctypes

import numpy as np

from common import host_harness


def _ulps(a, b):
    ia, ib = a.view(np.int64), b.view(np.int64)
    # map the sign-magnitude patterns onto a monotone integer line
    ia = np.where(ia < 0, np.int64(-2**63) - ia, ia)
    ib = np.where(ib < 0, np.int64(-2**63) - ib, ib)
    return np.abs(ia - ib)


def _sincos(x):
    hh = host_harness()
    hh.hh_sincos.restype = None
    s, c = np.empty_like(x), np.empty_like(x)
    P = ctypes.POINTER(ctypes.c_double)
    hh.hh_sincos(ctypes.c_int(len(x)), x.ctypes.data_as(P), s.ctypes.data_as(P), c.ctypes.data_as(P))
    return s, c


def test_shared_sincos_is_within_one_ulp_of_libm():
    rng = np.random.default_rng(12345)
    parts = [rng.uniform(-np.pi / 4, np.pi / 4, 200_000),      # |w| dt of a step: no range reduction
             rng.uniform(-10.0, 10.0, 200_000),
             rng.uniform(-1e6, 1e6, 200_000),                   # yaw/pitch/roll a caller may hand in
             np.array([0.0, -0.0, 1e-300, -1e-300, 1e-8, np.pi / 4, -np.pi / 4, np.pi / 2, np.pi, 1.0, 100.0]),
             rng.uniform(-1.0, 1.0, 100_000) * 10.0 ** rng.uniform(-12, 0, 100_000)]
    for x in parts:
        x = np.ascontiguousarray(x, dtype=np.float64)
        s, c = _sincos(x)
        us, uc = _ulps(s, np.sin(x)), _ulps(c, np.cos(x))
        assert us.max() <= 1 and uc.max() <= 1, (int(us.max()), int(uc.max()))
        assert (us == 0).mean() > 0.9 and (uc == 0).mean() > 0.9


def test_shared_sincos_special_values():
    x = np.array([np.inf, -np.inf, np.nan, 0.0, -0.0])
    s, c = _sincos(x)
    assert np.isnan(s[:3]).all() and np.isnan(c[:3]).all()
    assert s[3] == 0.0 and not np.signbit(s[3]) and np.signbit(s[4]) and c[3] == 1.0 and c[4] == 1.0
