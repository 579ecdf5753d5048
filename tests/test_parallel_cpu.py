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
