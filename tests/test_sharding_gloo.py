"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from optimalbeziertrajectorygeneration_b200 import sharding


def test_block_ranges_cover_and_balance():
    for total in (0, 1, 7, 523776, 1000003):
        for world in (1, 2, 3, 8):
            rs = [sharding.block_range(total, world, r) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == total
            for a, b in zip(rs[:-1], rs[1:]):
                assert a[1] == b[0]
            sizes = [e - b for b, e in rs]
            assert max(sizes) - min(sizes) <= 1
    assert sharding.pair_range(1024, 8, 7)[1] == 1024 * 1023 // 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        P, B = 45, 3
        full = torch.arange(world * B * P, dtype=torch.float64).reshape(world * B, P) * 0.5 - 7.0
        # batch mode: each rank owns B rows
        got = sharding.gather_pair_minima(full[rank * B:(rank + 1) * B].clone(), mode="batch")
        ok1 = torch.equal(got, full)
        # pair mode: unequal column ranges
        one = full[:B]
        b, e = sharding.block_range(P, world, rank)
        got2 = sharding.gather_pair_minima(one[:, b:e].clone(), mode="pairs", total=P)
        ok2 = torch.equal(got2, one)
        # the overlapped gatherer (synchronous on CPU): buffer rotation over several steps
        g = sharding.PairMinimaGatherer(B, P, "cpu")
        ok3 = True
        for step in range(5):
            loc = g.local_buffer()
            loc.copy_(full[rank * B:(rank + 1) * B] + step)
            ok3 = ok3 and torch.equal(g.gather(), full + step)
        g.finish()
        q.put((rank, bool(ok1), bool(ok2 and ok3)))
    finally:
        dist.destroy_process_group()


def test_allgather_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] for r in res)
