"""Two-GPU check of the fused all-gather (in-kernel NVLink peer stores into symmetric
memory) against a plain NCCL all-gather.  Skipped on boxes with fewer than two GPUs."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from oracle.make_golden import synthetic_swarm_args
    from optimalbeziertrajectorygeneration_b200 import optimization as gopt
    from optimalbeziertrajectorygeneration_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        args, x = synthetic_swarm_args(70)
        b = gopt.BezOptimization(**args)
        eng = b._engine(True)
        B, E = 3, 100
        X = x[None, :] + np.random.default_rng(rank).normal(size=(B, x.size)) * 0.05
        cpts, _ = eng.assemble(eng.upload(X), E)
        P = 70 * 69 // 2
        pm = sharding.PeerMinima(B, P, eng.device)
        ok = True
        for step in range(5):
            local, peers = pm.targets()
            sep = eng.separation(cpts, E, 0.9 + 0.1 * step, pairmin=local, peer_ptrs=peers)
            gathered = pm.complete()
            pm.wait()
            ref = torch.empty((world * B, P), dtype=torch.float64, device=eng.device)
            dist.all_gather_into_tensor(ref, sep.min(dim=2).values.contiguous())
            torch.cuda.synchronize()
            ok = ok and torch.equal(ref, gathered)
        # back-to-back steps without a consumer in between (what bench.py does), checked at the end
        keep = []
        for step in range(7):
            local, peers = pm.targets()
            sep = eng.separation(cpts, E, 1.5 + 0.1 * step, pairmin=local, peer_ptrs=peers)
            gathered = pm.complete()
            keep.append(sep.min(dim=2).values.contiguous())
        pm.wait()
        ref = torch.empty((world * B, P), dtype=torch.float64, device=eng.device)
        dist.all_gather_into_tensor(ref, keep[-1])
        torch.cuda.synchronize()
        ok = ok and torch.equal(ref, gathered)
        # strong-scaling layout: one batch, the pair list cut into `world` ranges, fused gather
        # into one [B, P] matrix (min_pitch = P, pointers offset to the range's first column)
        ps = sharding.PeerMinima(B, P, eng.device, layout="pairs")
        Xs = x[None, :] + np.random.default_rng(99).normal(size=(B, x.size)) * 0.05      # same on every rank
        cs, _ = eng.assemble(eng.upload(Xs), E)
        full = torch.empty((B, P), dtype=torch.float64, device=eng.device)
        eng.separation(cs, E, 0.9, pairmin=full, rows=False)
        for step in range(4):
            local, peers = ps.targets()
            lo, hi = ps.pair_lo, ps.pair_hi
            eng.separation(cs, E, 0.9, pair_begin=lo, npairs=hi - lo, pairmin=local, min_pitch=P,
                           peer_ptrs=peers, rows=False)
            g2 = ps.complete()
        ps.wait()
        torch.cuda.synchronize()
        ok = ok and torch.equal(g2, full)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_fused_allgather_matches_nccl():
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(r[1] for r in res)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_engine_on_a_non_current_device():
    """BezOptimization(device=1) while device 0 is current: the per-device kernel
    attributes, the launch stream and the allocations must all follow the engine's device, and
    the caller's current device must be left alone."""
    from oracle.make_golden import synthetic_swarm_args
    from optimalbeziertrajectorygeneration_b200 import optimization as gopt
    torch.cuda.set_device(0)
    args, x = synthetic_swarm_args(20)
    gopt.DEG_ELEV = 100
    try:
        b0 = gopt.BezOptimization(**args, device=0)
        b1 = gopt.BezOptimization(**args, device=1)
        want = b0.temporalSeparationConstraints(x)          # ~96 KB dynamic shared memory: opt-in on device 0
        got = b1.temporalSeparationConstraints(x)           # ... must also be granted on device 1
        assert torch.cuda.current_device() == 0
        assert np.array_equal(got, want)
        assert np.array_equal(b1.maxSpeedConstraints(x), b0.maxSpeedConstraints(x))
        assert np.array_equal(b1.temporalSeparationConstraints_jac(x), b0.temporalSeparationConstraints_jac(x))
        red = b1.evaluate_sweep_active(np.stack([x, x + 0.01]), elev=100, chunk=2)
        flags, _, _, _ = red.pairs(0)
        assert flags.shape == (2, 190)
        assert b1._engine(True).device.index == 1 and torch.cuda.current_device() == 0
    finally:
        gopt.DEG_ELEV = 0
