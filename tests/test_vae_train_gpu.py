"""Training step of the motion-VAE decoder (BASELINE configs[2] / [4]; reference train_vae.py:293-353 back-propagates through
model/autoencoder.py:552-609): the hand-written backward (gvfdiffusion_b200/vae_train.py + csrc/attn_bwd.cu + csrc/backward.cu)
against torch autograd of the CPU oracle (oracle/vae.py, pinned to the reference's module by tests/test_oracle_golden.py).

Tolerance: the reference's own autocast backward carries fp16 activation gradients; relative L2 <= 1e-2 per gradient
tensor against fp32 autograd of the oracle (measured values are printed), forward <= 2e-3."""
import os

import pytest
import torch

from oracle import vae as OVAE

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEV = "cuda"


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


def _oracle_grads(sd, z, q, heads, T, R, precision):
    sd = {k: v.clone().float().requires_grad_(True) for k, v in sd.items()}
    z = z.clone().requires_grad_(True)
    q = q.clone().requires_grad_(True)
    out = OVAE.vae_decode(sd, z, q, heads, T, precision)
    (out * R).sum().backward()
    return out.detach(), z.grad, q.grad, {k: v.grad for k, v in sd.items()}


def _check(v, sd, z, q, heads, T, tol, label):
    torch.manual_seed(5)
    v = v.to(DEV)
    zd, qd = z.to(DEV).requires_grad_(True), q.to(DEV).requires_grad_(True)
    out = v.decode(zd, qd)
    assert out.requires_grad
    R = torch.randn(out.shape, generator=torch.Generator().manual_seed(9))
    (out * R.to(DEV)).sum().backward()
    with torch.no_grad():
        inf = v.decode(zd.detach(), qd.detach())
    assert rel(out, inf) < 1e-3, ("training forward vs inference engine", rel(out, inf))
    worst = {}
    for prec in ("fp32", "fp16"):
        o_ref, dz, dq, gw = _oracle_grads(sd, z, q, heads, T, R, prec)
        errs = {"out": rel(out, o_ref), "dz": rel(zd.grad, dz), "dqueries": rel(qd.grad, dq)}
        named = dict(v.named_parameters())
        for n in v._param_names:                                  # every DECODE parameter (the encoder is forward-only)
            p = named[n]
            assert p.grad is not None and p.grad.shape == p.shape, n
            errs[n] = rel(p.grad, gw[n])
        worst[prec] = max(errs.values())
        top = sorted(errs.items(), key=lambda kv: -kv[1])[:4]
        print(f"{label} vs oracle autograd ({prec}): forward {errs['out']:.2e}, dz {errs['dz']:.2e}, dqueries "
              f"{errs['dqueries']:.2e}, worst " + ", ".join(f"{k} {e:.2e}" for k, e in top))
        assert errs["out"] < 4e-3
        assert max(errs.values()) < tol, top
    return worst


def test_decode_backward_tiny_golden():
    """The reference's own tiny module (tests/golden/vae_tiny.pt: dim 96, 3 heads of 32, 2 layers, T = 3)."""
    from gvfdiffusion_b200.model.autoencoder import GSKLTemporalVariationalAutoEncoder as VAE
    g = torch.load(os.path.join(G, "vae_tiny.pt"), weights_only=False)
    v = VAE(**g["cfg"])
    v.load_state_dict(g["state_dict"])
    sd = {k: t.clone() for k, t in v.state_dict().items()}
    _check(v, sd, g["z"], g["queries"], g["cfg"]["heads"], g["cfg"]["num_timesteps"], 1e-2, "tiny golden")


@pytest.mark.parametrize("B,Q", [(1, 300), (2, 257)])
def test_decode_backward_shipped_width(B, Q):
    """Shipped width (dim 768, 12 heads of 64, GEGLU 6144), 2 layers, T = 2, 512 latents; B = 2 exercises the per-object
    decoder attention with shared queries and the stacked wgrads."""
    from gvfdiffusion_b200.model.autoencoder import GSKLTemporalVariationalAutoEncoder as VAE
    torch.manual_seed(1)
    cfg = dict(depth=2, dim=768, queries_dim=768, output_dim=14, num_inputs=8192, num_latents=512, latent_dim=16,
               heads=12, dim_head=-1, num_timesteps=2)
    v = VAE(**cfg)
    gen = torch.Generator().manual_seed(2)
    for p in v.parameters():
        if p.abs().sum() == 0:
            p.data = torch.randn(p.shape, generator=gen) * 0.05
    sd = {k: t.clone() for k, t in v.state_dict().items()}
    z = torch.randn(B * 2, 512, 16, generator=gen)
    q = torch.randn(B, Q, 14, generator=gen) * 0.3
    _check(v, sd, z, q, 12, 2, 1e-2, f"dim 768 B={B}")


def test_decode_train_step_updates_follow_new_weights():
    """An optimiser step changes the parameters in place: the engines must pick the new weights up (version check)."""
    from gvfdiffusion_b200.model.autoencoder import GSKLTemporalVariationalAutoEncoder as VAE
    g = torch.load(os.path.join(G, "vae_tiny.pt"), weights_only=False)
    v = VAE(**g["cfg"])
    v.load_state_dict(g["state_dict"])
    v = v.to(DEV)
    opt = torch.optim.SGD(v.parameters(), lr=0.5)
    z, q = g["z"].to(DEV), g["queries"].to(DEV)
    o0 = v.decode(z, q)
    o0.square().mean().backward()
    opt.step()
    o1 = v.decode(z, q)
    assert float(o1.square().mean()) < float(o0.square().mean())
    with torch.no_grad():
        assert rel(v.decode(z, q), o1) < 1e-3
