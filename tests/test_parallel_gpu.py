"""The data-parallel path on real hardware (SURVEY.md section 8e, section 4 "N ranks, distinct objects, gather,
compare with 1-rank results bit-for-bit"): two ranks over NCCL, rank 0 owns both objects in pinned host memory,
`PipelinedExchange` scatters the conditioning and gathers the rendered frames around sample -> decode -> render
on the tiny golden models; the gathered frames must EQUAL what one process computes for the same objects.
Needs two GPUs (skipped on a one-GPU box; run with `gpurun --gpus 2`)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STEPS, RES, DPM_STEPS = 3, 64, 4


def _build(dev):
    from gvfdiffusion_b200.model.autoencoder import GSKLTemporalVariationalAutoEncoder as VAE
    from gvfdiffusion_b200.model.dit import DiT
    from gvfdiffusion_b200.pipeline import GVFPipeline
    from oracle import dpm as ODPM
    gd = torch.load(os.path.join(G, "dit_tiny.pt"), weights_only=False)
    gv = torch.load(os.path.join(G, "vae_tiny.pt"), weights_only=False)
    dit, vae = DiT(**gd["cfg"]), VAE(**gv["cfg"])
    dit.load_state_dict(gd["state_dict"])
    vae.load_state_dict(gv["state_dict"])
    pipe = GVFPipeline(dit.to(dev).eval(), vae.to(dev).eval(), torch.from_numpy(ODPM.reference_betas(1000)), device=dev,
                       resolution=RES, num_latents=gd["cfg"]["resolution"], num_static=40)
    return pipe, gd["cfg"], gv["cfg"]


def _object(step, r, dcfg, T):
    """Object of rank r at step `step`: canonical Gaussians + conditioning + noise (host tensors)."""
    from gvfdiffusion_b200 import synthetic as S
    seed = 1000 + 10 * step + r
    canon = S.canonical_gaussians(num_voxels=64, seed=seed)
    g = torch.Generator().manual_seed(seed)
    d = {("canon." + k): v.contiguous() for k, v in canon.items()}
    d["cond_images"] = torch.randn(1, T, 10, dcfg["image_cond_channels"], generator=g)
    d["noise"] = torch.randn(1, T, dcfg["resolution"], 16, generator=g)
    return d


def _compute(pipe, inp, out, T):
    from gvfdiffusion_b200 import synthetic as S
    o = pipe.prepare_object({n[6:]: t for n, t in inp.items() if n.startswith("canon.")})
    lat = pipe.sample(o, inp["cond_images"], inp["noise"], steps=DPM_STEPS)
    pipe.render(o, pipe.decode(lat, o), S.orbit_extrinsics(T), S.intrinsics(), out=out)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from gvfdiffusion_b200.parallel import PipelinedExchange
    pipe, dcfg, vcfg = _build(dev)
    T = vcfg["num_timesteps"]
    proto = _object(0, 0, dcfg, T)
    specs = {k: (tuple(v.shape), v.dtype) for k, v in proto.items()}
    ex = PipelinedExchange(specs, (T, 4, RES, RES), dev)
    host = lambda k: [{n: t.pin_memory() for n, t in _object(k, r, dcfg, T).items()} for r in range(world)] if rank == 0 else None
    got = []
    ex.prime(host(0))
    for k in range(STEPS):
        last = k + 1 == STEPS
        ex.post(k, None if last else host(k + 1), scatter=not last)
        _compute(pipe, ex.inputs(k), ex.output(k), T)
        ex.done(k)
        if rank == 0 and k > 0:
            ex.comm.synchronize()                              # frames of step k - 1 are in host memory
            got.append([t.clone() for t in ex.host_results])
    ex.flush(STEPS)
    torch.cuda.synchronize()
    if rank == 0:
        got.append([t.clone() for t in ex.host_results])
        # the same objects, one after the other, on this rank alone
        worst = 0.0
        for k in range(STEPS):
            for r in range(world):
                inp = {n: t.to(dev) for n, t in _object(k, r, dcfg, T).items()}
                out = torch.empty((T, 4, RES, RES), dtype=torch.float32, device=dev)
                _compute(pipe, inp, out, T)
                same = torch.equal(out.cpu(), got[k][r])
                worst = max(worst, float((out.cpu() - got[k][r]).abs().max()))
                if not same:
                    q.put(("mismatch", k, r, worst))
                    break
        q.put(("ok", worst, ex.bytes_p2p, float(got[0][0].std())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_ranks_scatter_gather_equals_one_rank():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res[0] == "ok", res
    assert res[1] == 0.0                       # bit for bit
    assert res[2] > 0                          # bytes really went over NCCL point-to-point
    assert res[3] > 0                          # and the frames are not blank
