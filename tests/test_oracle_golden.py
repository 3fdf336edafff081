"""The oracle against the committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the
oracle itself — a regression pin, not reference output), and the device physics headers (host build) against the
same vectors."""
import os

import numpy as np
import pytest

from common import ROOT, P, to_native_config, same_bits, harness_step, bits
from golden.make_golden import CASES, generate


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden(name):
    gold = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    now = generate(name)
    assert sorted(gold.files) == sorted(now.keys())
    for k in gold.files:
        a, b = gold[k], now[k]
        if a.dtype == np.float64:
            assert same_bits(a, b), f"{name}: {k}"
        else:
            assert np.array_equal(a, b), f"{name}: {k}"


@pytest.mark.parametrize("name", ["stack_scene", "c2_world0", "c3_world0_small"])
def test_device_physics_host_build_reproduces_golden(name):
    gold = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    mk, cfg, marks = CASES[name]
    b = P.build_world(mk(), cfg.StepConfig.GravitationalAccelerationMagnitude, cfg.StepConfig.GravityDirection)
    nc = to_native_config(cfg)
    done = 0
    for m in marks:
        st, rec, _ = harness_step(b, nc, m - done)
        done = m
        b.update(st)
        for k in ("pos", "vel", "rot", "ang_mom"):
            assert same_bits(st[k], gold[f"s{m}_{k}"]), f"{name} step {m} {k}"
        assert np.array_equal(rec["pairs"], gold[f"s{m}_pairs"])
        assert np.array_equal(rec["feature"], gold[f"s{m}_feature"])
