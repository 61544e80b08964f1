"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: node-range partition, halo exchange, fused gradient
all-reduce.  The row gather is injected (the product default is the CUDA kernel, which has no CPU path)."""
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


def _worker(rank, world, port, fn, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _spawn(fn, world=2):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, fn, ret), nprocs=world, join=True)
    return dict(ret)


def _cpu_gather(table, ids):
    return table[ids]


def _halo_job(rank, world):
    from dgll_b200 import parallel as P
    n, f = 1003, 7                                   # not divisible by world
    full = torch.arange(n * f, dtype=torch.float32).reshape(n, f)
    lo, hi = P.local_range(rank, n, world)
    hx = P.HaloExchange(n, full[lo:hi].clone(), gather_fn=_cpu_gather)
    g = torch.Generator().manual_seed(100 + rank)
    ok = True
    hx.set_bucket_capacity(1003, slack=1.5)
    for m in (0, 1, 257, 1003):
        ids = torch.randperm(n, generator=g)[:m]
        out = hx.fetch(ids)
        ok = ok and out.shape == (m, f) and torch.equal(out, full[ids])
        if m:
            out2 = hx.fetch_padded(ids)
            ok = ok and torch.equal(out2, full[ids]) and not hx.check_overflow()
    # a deliberately tiny capacity must raise the overflow flag (and only zero the rows that did not fit)
    ids = torch.randperm(n, generator=g)[:400]
    out3 = hx.fetch_padded(ids, cap=16)
    bad = (out3 != full[ids]).any(dim=1)
    ok = ok and hx.check_overflow() and bool((out3[bad] == 0).all()) and int((~bad).sum()) >= 16
    # ids all owned by ONE rank (empty buckets elsewhere), duplicates allowed
    ids = torch.randint(0, hi - lo, (50,), generator=g) + (0 if rank else P.part_size(n, world))
    ids = ids.clamp(max=n - 1)
    ok = ok and torch.equal(hx.fetch(ids), full[ids])
    return ok, hx.stats["rows"], hx.stats["remote_rows"]


def test_halo_exchange_world2():
    res = _spawn(_halo_job, 2)
    assert all(v[0] for v in res.values()), res
    assert all(v[2] > 0 for v in res.values())       # something really crossed ranks


def _grad_job(rank, world):
    from dgll_b200 import parallel as P
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.Linear(4, 3))
    x = torch.full((2, 5), float(rank + 1))
    model(x).sum().backward()
    mine = [p.grad.clone() for p in model.parameters()]
    P.allreduce_gradients(model.parameters())
    gathered = []
    for g in mine:
        buf = [torch.zeros_like(g) for _ in range(world)]
        dist.all_gather(buf, g)
        gathered.append(sum(buf) / world)
    return all(torch.allclose(p.grad, e) for p, e in zip(model.parameters(), gathered))


def test_allreduce_gradients_world2():
    assert all(_spawn(_grad_job, 2).values())


def test_partition_helpers():
    from dgll_b200 import parallel as P
    n, w = 111059956, 8
    ids = torch.tensor([0, 13882494, 13882495, 111059955])
    assert P.part_size(n, w) == 13882495
    assert P.owner_of(ids, n, w).tolist() == [0, 0, 1, 7]
    assert P.local_range(7, n, w) == (97177465, 111059956)
    seeds = torch.arange(0, 100)
    assert P.shard_seeds(seeds, 100, 1, 4).tolist() == list(range(25, 50))
    with pytest.raises(ValueError):
        P.HaloExchange(10, torch.zeros(3, 2))
