"""The N>1 path on CPU: two gloo ranks shard a batch into contiguous world ranges and all-reduce the statistics
block — the only collective of the path.  (Compute is per-GPU and needs no exchange)"""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from common import ROOT  # noqa: F401
from physicssimulator_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(101, rank, world)
    local = {k: 0 for k in sharding.SUM_KEYS}
    local["steps"] = (hi - lo) * 10
    local["body_steps"] = (hi - lo) * 10 * 64
    local["pairs_colliding"] = rank + 5
    local["max_penetration"] = 0.25 * (rank + 1)
    out = sharding.allreduce_stats(local)
    q.put((rank, lo, hi, out))
    dist.destroy_process_group()


def test_two_rank_sharding_and_stats_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, s0), (r1, lo1, hi1, s1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 50, 50, 101)
    assert s0 == s1
    assert s0["steps"] == 1010 and s0["body_steps"] == 1010 * 64 and s0["pairs_colliding"] == 11
    assert s0["max_penetration"] == 0.5


def test_single_process_is_identity():
    s = {k: 1 for k in sharding.SUM_KEYS}
    s["max_penetration"] = 0.1
    assert sharding.allreduce_stats(s) == s
