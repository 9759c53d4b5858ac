"""N>1 host path on CPU: world_size-2 gloo processes -- batch sharding, the flat gradient all-reduce (mean) and the
rank-0 buffer broadcast give exactly the single-process result on the full batch."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1, bias=False), torch.nn.BatchNorm2d(8), torch.nn.ELU(),
                               torch.nn.Conv2d(8, 1, 1, bias=False))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bts_b200 import dist as D
    m = _model()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(4, 3, 8, 8, generator=g)
    (xs,) = D.shard_batch([x], rank, world)
    red = D.FlatGradReducer(m.parameters())
    # per-rank BN statistics (no SyncBN, as the reference); loss = mean over the LOCAL shard
    m(xs).pow(2).mean().backward()
    red.reduce()
    if rank == 1:                       # make rank 1's buffers differ, then broadcast rank 0's
        m[1].running_mean.add_(1.0)
    D.broadcast_buffers(m, 0)
    q.put((rank, [p.grad.clone() for p in m.parameters()], m[1].running_mean.clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_allreduce_matches_mean_of_shard_gradients():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    import socket
    with socket.socket() as sk:                 # a port that is free right now (a fixed one can clash with another job)
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    # expected: mean over ranks of the gradient each shard produces on its own
    g = torch.Generator().manual_seed(1)
    x = torch.randn(4, 3, 8, 8, generator=g)
    expect, rm0 = None, None
    for r in range(world):
        m = _model()
        m(x[r * 2:(r + 1) * 2]).pow(2).mean().backward()
        gr = [p.grad.clone() for p in m.parameters()]
        expect = gr if expect is None else [a + b for a, b in zip(expect, gr)]
        if r == 0:
            rm0 = m[1].running_mean.clone()
    expect = [e / world for e in expect]
    for rank, grads, rm in res:
        for a, b in zip(grads, expect):
            assert torch.allclose(a, b, rtol=1e-6, atol=1e-7)
        assert torch.equal(rm, rm0)                         # rank-0 running stats everywhere
    assert all(torch.equal(a, b) for a, b in zip(res[0][1], res[1][1]))   # identical on every rank


def test_shard_batch_rejects_indivisible_batches():
    from bts_b200 import dist as D
    with pytest.raises(ValueError):
        D.shard_batch([torch.zeros(5, 3)], 0, 2)
