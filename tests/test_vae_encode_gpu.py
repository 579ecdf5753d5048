"""Motion-VAE encoder (reference model/autoencoder.py:502-550) on the device engine against the reference's
own `encode` output (tests/golden/vae_encode_tiny.pt, recorded on the CPU with deterministic stand-ins for the absent
torch_cluster.fps / pytorch3d.knn_points) and against the oracle restatement at the shipped width."""
import os

import pytest
import torch

from oracle import vae as OVAE

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEV = "cuda"


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / b.norm())


def test_encode_golden_fixture():
    from gvfdiffusion_b200.model.autoencoder import GSKLTemporalVariationalAutoEncoder as VAE
    g = torch.load(os.path.join(G, "vae_encode_tiny.pt"), weights_only=False)
    dec = torch.load(os.path.join(G, "vae_tiny.pt"), weights_only=False)
    v = VAE(**g["cfg"])
    v.load_state_dict({**dec["state_dict"], **g["state_dict"]})
    v = v.to(DEV)
    kl, x, post, sgs = v.encode(g["static_pc"].to(DEV), g["delta_pc"].to(DEV), [t.to(DEV) for t in g["static_gs"]],
                                noise=g["noise"])
    ref16, ref32 = g["autocast_fp16"], g["fp32"]
    assert torch.equal(sgs.cpu(), ref32["sampled_static_gs"])                 # same farthest point sample
    errs = {k: (rel(t, ref16[k]), rel(t, ref32[k])) for k, t in (("mean", post["mean"]), ("logvar", post["logvar"]), ("x", x))}
    print("encode vs reference (autocast fp16, fp32):", {k: (f"{a:.2e}", f"{b:.2e}") for k, (a, b) in errs.items()})
    for k, (a, b) in errs.items():
        assert a < 2e-3 and b < 3e-3, (k, a, b)
    assert torch.allclose(kl.cpu(), ref32["kl"], rtol=5e-3, atol=1e-5)
    # decode-only checkpoints still load; their encoder refuses to run
    v2 = VAE(**g["cfg"])
    v2.load_state_dict(dec["state_dict"])
    with pytest.raises(RuntimeError):
        v2.to(DEV).encode(g["static_pc"].to(DEV), g["delta_pc"].to(DEV), [t.to(DEV) for t in g["static_gs"]])
    # forward(): encode -> pad_static_gs -> decode (model/autoencoder.py:620-627)
    out = v(g["static_gs"], g["static_pc"], g["delta_pc"])
    assert out["logits"].shape == (2, g["cfg"]["num_timesteps"], 100, g["cfg"]["output_dim"]) and out["kl"].shape == (6,)


def test_encode_shipped_width_matches_oracle():
    """dim 768, 12 heads of 64, 512 anchors from 3000 Gaussians, 2048 tracked points, T = 4, K = 8."""
    from gvfdiffusion_b200.model.autoencoder import GSKLTemporalVariationalAutoEncoder as VAE
    torch.manual_seed(3)
    cfg = dict(depth=1, dim=768, queries_dim=768, output_dim=14, num_inputs=2048, num_latents=512, latent_dim=16, heads=12,
               dim_head=-1, num_timesteps=4, knn_k=8, beta=7.0)
    v = VAE(**cfg)
    gen = torch.Generator().manual_seed(4)
    for p in v.parameters():
        if p.abs().sum() == 0:
            p.data = torch.randn(p.shape, generator=gen) * 0.05
    sd = {k: t.clone() for k, t in v.state_dict().items()}
    static_pc = torch.rand(1, 2048, 3, generator=gen) - 0.5
    delta_pc = torch.randn(1, 4, 2048, 3, generator=gen) * 0.05
    gs = torch.randn(3000, 14, generator=gen) * 0.3
    gs[:, :3] = torch.rand(3000, 3, generator=gen) - 0.5
    noise = torch.randn(4, 512, 16, generator=gen)
    kl, x, post, sgs = v.to(DEV).encode(static_pc.to(DEV), delta_pc.to(DEV), [gs.to(DEV)], noise=noise)
    o = OVAE.vae_encode(sd, static_pc, delta_pc, [gs], 12, 512, 8, 7.0, "fp16", noise=noise)
    assert torch.equal(sgs.cpu(), o["sampled_static_gs"])
    e = {k: rel(t, o[k]) for k, t in (("mean", post["mean"]), ("logvar", post["logvar"]), ("x", x))}
    print("encode (dim 768) vs oracle fp16:", {k: f"{a:.2e}" for k, a in e.items()})
    assert max(e.values()) < 2e-3, e
    assert torch.allclose(kl.cpu(), o["kl"], rtol=5e-3, atol=1e-5)
