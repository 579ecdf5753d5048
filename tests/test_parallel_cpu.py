"""world_size-2 gloo test of the object sharding / scatter / gather host logic (the N > 1 path of
bench.py and inference_dpm_latent.py): the gathered result must equal the 1-rank result bit for bit."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gvfdiffusion_b200 import parallel as P


def _fake_process(i, cond):
    # stands in for sample+decode+render: any deterministic function of the object's conditioning
    return (cond["a"] * (i + 1)).cumsum(-1) + cond["b"].sum()


def _worker(rank, world, port, n_obj, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    conds = [{"a": torch.randn(3, 5, generator=g), "b": torch.randn(4, generator=g)} for _ in range(n_obj)]
    mine = P.scatter_conditioning(conds if rank == 0 else None, n_obj, src=0)
    assert sorted(mine) == P.object_shard(n_obj, rank, world)
    res = {i: _fake_process(i, c) for i, c in mine.items()}
    out = P.gather_results(res, n_obj, dst=0)
    if rank == 0:
        q.put([t.clone() for t in out])
    dist.barrier()
    dist.destroy_process_group()


def test_shard_scatter_gather_world2():
    assert P.object_shard(5, 0, 2) == [0, 2, 4] and P.object_shard(5, 1, 2) == [1, 3]
    assert P.object_shard(1, 1, 2) == []
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    n_obj = 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_obj, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(0)
    conds = [{"a": torch.randn(3, 5, generator=g), "b": torch.randn(4, generator=g)} for _ in range(n_obj)]
    for i in range(n_obj):
        assert torch.equal(got[i], _fake_process(i, conds[i]))


def _worker_pipelined(rank, world, port, n_steps, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    specs = {"a": ((3, 5), torch.float32), "b": ((4,), torch.float32)}
    ex = P.PipelinedExchange(specs, (3, 5), "cpu")
    def host(k):                                   # the root's objects of step k, one per rank
        if rank != 0:
            return None
        g = torch.Generator().manual_seed(100 + k)
        return [{"a": torch.randn(3, 5, generator=g), "b": torch.randn(4, generator=g)} for _ in range(world)]
    got = []
    ex.prime(host(0))
    for k in range(n_steps):
        last = k + 1 == n_steps
        ex.post(k, None if last else host(k + 1), scatter=not last)
        if rank == 0 and k > 0:
            got.append([t.clone() for t in ex.host_results])          # frames of step k - 1
        inp, out = ex.inputs(k), ex.output(k)
        out.copy_(_fake_process(rank, inp))
        ex.done(k)
    ex.flush(n_steps)
    if rank == 0:
        got.append([t.clone() for t in ex.host_results])
        q.put(got)
    dist.barrier()
    dist.destroy_process_group()


def test_pipelined_exchange_world2():
    """PipelinedExchange (scatter of step k+1 and gather of step k-1 around compute k): every step's gathered
    frames equal what a single process computes from the same objects."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    n_steps, world = 4, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_pipelined, args=(r, world, port, n_steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert len(got) == n_steps
    for k in range(n_steps):
        g = torch.Generator().manual_seed(100 + k)
        conds = [{"a": torch.randn(3, 5, generator=g), "b": torch.randn(4, generator=g)} for _ in range(world)]
        for r in range(world):
            assert torch.equal(got[k][r], _fake_process(r, conds[r]))
