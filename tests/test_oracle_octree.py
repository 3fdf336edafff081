"""T0: the oracle's SpatialTree restatement AND the product's host-side tree (csrc/ps_octree.hpp, an independent
implementation: flat node pool, sorted vectors) against the ONLY golden vectors the reference holds for this path:
the five known-answer facts of /root/reference/PhysicsSimulator.Tests/UtilitiesTests.fs:159-276, plus the seven
properties of :24-157 as randomized checks.  (Extents below are (Size, Position) like the reference's ObjectExtent.)"""
import random

import pytest

from common import ProductTree
from oracle.oracle import OracleTree


@pytest.fixture(params=["oracle", "product"])
def Tree(request):
    return OracleTree if request.param == "oracle" else ProductTree



def _tree_from_posmap(Tree, bounds, posmap):
    """singleNodeTreeFromPosMap (UtilitiesTests.fs:14-18): capacity = count-1, depth 10, root leaf = first count-1 keys"""
    count = len(posmap) - 1
    t = Tree(count, 10, bounds)
    for k, ext in posmap.items():
        t.set_extent(k, ext)
    t.set_root_leaf(sorted(posmap)[:count])
    return t


def test_fact_1d_straddling_insert(Tree):  # UtilitiesTests.fs:159-179
    posmap = {1: [(1.0, 2.0)], 2: [(1.0, 7.0)], 3: [(1.1, 4.0)]}
    t = _tree_from_posmap(Tree, [(0, 10)], posmap)
    assert t.insert(3) == 0
    assert t.dump() == "N[L{1,3} L{2,3}]"


def test_fact_1d_simple(Tree):  # :181-200
    posmap = {1: [(2.0, 2.0)], 2: [(1.0, 7.0)], 3: [(1.1, 3.0)]}
    t = _tree_from_posmap(Tree, [(0, 10)], posmap)
    assert t.insert(3) == 0
    assert t.dump() == "N[L{1,3} L{2}]"


def test_fact_1d_simple_2(Tree):  # :202-221
    posmap = {1: [(2.0, 12.0)], 2: [(1.0, 17.0)], 3: [(1.1, 18.0)]}
    t = _tree_from_posmap(Tree, [(10, 20)], posmap)
    assert t.insert(3) == 0
    assert t.dump() == "N[L{1} L{2,3}]"


def test_fact_1d_complex_nested_split(Tree):  # :223-250
    posmap = {1: [(2.0, 0.0)], 2: [(1.0, 6.0)], 3: [(1.1, 17.0)], 4: [(1.1, 1.0)]}
    t = Tree(2, 10, [(0, 20)])
    for k, e in posmap.items():
        t.set_extent(k, e)
    t.set_root_children([[1, 2], [3]])
    assert t.insert(4) == 0
    assert t.dump() == "N[N[L{1,4} L{2}] L{3}]"


def test_fact_2d_simple_child_index_to_quadrant(Tree):  # :253-276
    posmap = {1: [(2.0, 2.0), (2.0, 2.0)], 2: [(1.0, 2.0), (2.0, 7.0)], 3: [(1.0, 7.0), (2.0, 2.0)],
              4: [(1.0, 7.0), (2.0, 7.0)], 5: [(1.1, 1.0), (2.0, 2.0)]}
    t = _tree_from_posmap(Tree, [(0, 10), (0, 10)], posmap)
    assert t.insert(5) == 0
    assert t.dump() == "N[L{1,5} L{2} L{3} L{4}]"


# ---- the seven properties (:24-157), randomized --------------------------------------------------------------
def _rand_bounds(rng, ndim):
    out = []
    for _ in range(ndim):
        lo = rng.uniform(-1000, 1000)
        out.append((lo, lo + rng.uniform(0.1, 1000)))
    return out


def test_prop_init_with_empty_boundaries_throws(Tree):  # :24-27
    with pytest.raises(ValueError):
        Tree(3, 3, [])


def test_prop_init_returns_empty_leaf_and_no_buckets(Tree):  # :29-50
    rng = random.Random(1)
    for _ in range(50):
        t = Tree(rng.randint(1, 9), rng.randint(1, 9), _rand_bounds(rng, rng.randint(1, 3)))
        assert t.dump() == "L{}"
        assert t.num_buckets() == 0


def test_prop_single_insert_gives_singleton_bucket(Tree):  # :52-73
    rng = random.Random(2)
    for _ in range(50):
        b = _rand_bounds(rng, rng.randint(1, 3))
        t = Tree(rng.randint(1, 5), rng.randint(1, 5), b)
        ext = []
        for lo, hi in b:
            p = rng.uniform(lo, hi)
            ext.append((rng.uniform(0, hi - p) + 1e-9, p))
        t.set_extent(7, ext)
        assert t.insert(7) == 0
        assert t.dump() == "L{7}" and t.num_buckets() == 1


def test_prop_full_root_splits_when_depth_allows(Tree):  # :75-110
    rng = random.Random(3)
    for _ in range(30):
        cap = rng.randint(1, 4)
        ndim = rng.randint(1, 3)
        t = Tree(cap, rng.randint(2, 6), [(0.0, 100.0)] * ndim)
        for k in range(cap + 1):
            t.set_extent(k, [(1.0, rng.uniform(1, 98))] * ndim)
        t.set_root_leaf(range(cap))
        assert t.insert(cap) == 0
        d = t.dump()
        assert d.startswith("N[") and d.count("L{") == 2 ** ndim


def test_prop_full_root_at_max_depth_just_grows(Tree):  # :112-136
    rng = random.Random(4)
    for _ in range(30):
        cap = rng.randint(1, 4)
        t = Tree(cap, 1, [(0.0, 100.0)])
        for k in range(cap + 1):
            t.set_extent(k, [(1.0, rng.uniform(1, 98))])
        t.set_root_leaf(range(cap))
        assert t.insert(cap) == 0
        assert t.dump() == "L{" + ",".join(str(k) for k in range(cap + 1)) + "}"


def test_prop_out_of_bounds_insert_throws(Tree):  # :138-157
    t = Tree(2, 5, [(10.0, 20.0), (10.0, 20.0)])
    t.set_extent(1, [(0.0, 0.0), (0.0, 0.0)])
    assert t.insert(1) == -1
    t.set_extent(2, [(1.0, 30.0), (1.0, 12.0)])
    assert t.insert(2) == -1
