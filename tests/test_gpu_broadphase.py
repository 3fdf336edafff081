"""T9: grid broad phase (BASELINE config 5) — sorted pair list identical to the oracle's, both aabb modes,
bodies straddling cells, a huge static ground, and subsamples cross-checked against all-pairs."""
import numpy as np
import pytest

from common import scenes

pytestmark = pytest.mark.gpu


def _check(b, mode, method):
    from oracle import oracle
    from physicssimulator_b200 import _native
    ref = oracle.overlap_pairs(b, aabb_mode=mode, method=method, cap=max(64 * len(b["mass"]), 1 << 16))
    got, ms = _native.broadphase_pairs(b, aabb_mode=mode, capacity=max(64 * len(b["mass"]), 1 << 16))
    assert got.shape == ref.shape, f"pair count {got.shape[0]} vs oracle {ref.shape[0]}"
    assert np.array_equal(got, ref)
    return len(ref), ms


@pytest.mark.parametrize("mode", [0, 1])
def test_c5_subsample_vs_all_pairs(mode):
    b = scenes.c5_big(n_spheres=6000, n_boxes=600)
    # squeeze the slab so that there are plenty of overlaps
    b["pos"][1:, :2] *= 0.2
    b["pos"][1:, 2] = 0.3 + 0.1 * b["pos"][1:, 2]
    n, _ = _check(b, mode, method=0)
    assert n > 500


@pytest.mark.parametrize("mode", [0, 1])
def test_c5_medium_vs_sweep(mode):
    b = scenes.c5_big(n_spheres=180000, n_boxes=20000)
    b["pos"][1:, 2] = 0.2 + 0.05 * b["pos"][1:, 2]  # near the ground: the big static body overlaps many
    n, ms = _check(b, mode, method=1)
    assert n > 1000


def test_degenerate_inputs():
    from physicssimulator_b200 import _native
    b = scenes.c5_big(n_spheres=1, n_boxes=0)
    got, _ = _native.broadphase_pairs(b)
    assert got.shape[0] == 0
    # all bodies at one point
    b = scenes.c5_big(n_spheres=300, n_boxes=30)
    b["pos"][1:] = (1.0, 2.0, 3.0)
    _check(b, 0, 0)
    _check(b, 1, 0)


def test_c5_full_size_properties():
    """full BASELINE size (1M spheres + 100K boxes + ground): size-independent properties + oracle sweep"""
    from physicssimulator_b200 import _native
    b = scenes.c5_big()
    got, ms = _native.broadphase_pairs(b, aabb_mode=1, capacity=1 << 24)
    assert np.all(got[:, 0] < got[:, 1])
    key = got[:, 0].astype(np.int64) << 32 | got[:, 1]
    assert np.all(np.diff(key) > 0), "sorted and unique"
    # idempotence: same input, same list
    again, _ = _native.broadphase_pairs(b, aabb_mode=1, capacity=1 << 24)
    assert np.array_equal(got, again)
    from oracle import oracle
    ref = oracle.overlap_pairs(b, aabb_mode=1, method=1, cap=1 << 24)
    assert np.array_equal(got, ref)
