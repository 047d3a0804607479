"""Host logic of the data-parallel gradient exchange with the gloo backend on CPU, world_size 2:
averaged gradients == mean of per-rank gradients, bucketing covers every tensor exactly once, rank-0 broadcast."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


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
        from models._ddp import DataParallelContext
        ctx = DataParallelContext(torch.device("cpu"), bucket_elems=50)
        assert DataParallelContext.from_env(torch.device("cpu")) is not None
        torch.manual_seed(100 + rank)
        shapes = [(7,), (3, 5), (40,), (2, 2, 3), (64,), (1,)]
        params = [torch.nn.Parameter(torch.randn(s)) for s in shapes]
        # rank-0 broadcast of parameters and buffers
        lin = torch.nn.Linear(3, 2)
        ctx.broadcast_module(lin)
        w = [torch.zeros_like(lin.weight) for _ in range(world)]
        dist.all_gather(w, lin.weight.data)
        assert torch.equal(w[0], w[1])
        # gradient averaging
        g_local = []
        for i, p in enumerate(params):
            p.grad = None if i == 3 and False else torch.randn_like(p)
            g_local.append(p.grad.clone())
        params[5].grad = None                        # a parameter unused at this depth on every rank
        opt = torch.optim.SGD(params, lr=0.1)
        scale = ctx.allreduce_grads(opt)
        assert abs(scale - 1.0 / world) < 1e-12
        gathered = []
        for i, g in enumerate(g_local):
            if i == 5:
                continue
            buf = [torch.zeros_like(g) for _ in range(world)]
            dist.all_gather(buf, g)
            gathered.append(sum(buf) / world)
        got = [p.grad * scale for i, p in enumerate(params) if i != 5]    # sums; the Adam kernel applies `scale`
        ok = all(torch.allclose(a, b, atol=1e-6) for a, b in zip(got, gathered)) and params[5].grad is None
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res), res
