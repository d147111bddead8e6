"""Host-side sharding / gradient-exchange logic on CPU with the gloo backend, world_size 2
(the data path itself has no collective; see gast_b200/dist.py)."""
import os
import socket
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gast_b200.dist import shard_range, forward_sharded, FlatGradBuffer


def test_shard_range_partitions():
    for n in (0, 1, 5, 8, 4096, 8191):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans[:-1], spans[1:]):
                assert b == c and b >= a
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        # (1) clip sharding: a stand-in "model" (per-clip function) over a ragged global batch
        x = torch.arange(7 * 3 * 2 * 2, dtype=torch.float32).reshape(7, 3, 2, 2)

        def fn(xs):
            return xs.sum(dim=(1, 3), keepdim=False)[:, None, :, None].expand(-1, 1, -1, 3).contiguous()
        y = forward_sharded(fn, x, gather=True)
        ok1 = torch.equal(y, fn(x))
        # (2) gradient exchange: flat buffer, sum then 1/world
        torch.manual_seed(0)
        lin = torch.nn.Linear(4, 3)
        fb = FlatGradBuffer(lin.parameters())
        inp = torch.full((2, 4), float(rank + 1))
        lin(inp).sum().backward()
        g_local = fb.flat.clone()
        fb.all_reduce_mean()
        gathered = [torch.empty_like(g_local) for _ in range(world)]
        dist.all_gather(gathered, g_local)
        ok2 = torch.allclose(fb.flat, sum(gathered) / world)
        ok3 = lin.weight.grad.data_ptr() == fb.flat.data_ptr()
        q.put((rank, bool(ok1), bool(ok2), bool(ok3)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_sharding_and_grad_allreduce():
    world = 2
    ctx = mp.get_context('spawn')
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
    assert all(r[1] and r[2] and r[3] for r in res), res
