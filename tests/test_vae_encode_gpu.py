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


def _oracle_train(sd, g, heads, cfg, R, wkl, noise):
    sd = {k: v.clone().float().requires_grad_(True) for k, v in sd.items()}
    o = OVAE.vae_encode(sd, g["static_pc"], g["delta_pc"], g["static_gs"], heads, cfg["num_latents"], cfg["knn_k"], cfg["beta"],
                        "fp16", noise=noise)                     # fp16 roundings of the forward, straight-through gradients
    ((o["x"] * R).sum() + wkl * o["kl"].sum()).backward()
    return {k: v.grad for k, v in sd.items() if v.grad is not None}


@pytest.mark.parametrize("attn_std", [None, 0.1])
def test_encode_backward_matches_oracle_autograd(attn_std):
    """Gradients of (sampled latent, KL) with respect to every encoder parameter against torch autograd of the oracle
    restatement, tiny golden configuration (golden weights, and to_q / to_kv re-drawn with std 0.1 = logits of order one);
    tolerance 1e-2 rel. L2 (fp16 activation gradients).  One gradient is bounded differently: d to_q.  The keys of this
    attention are embeddings of neighbouring points, i.e. nearly parallel vectors, and every row of dS sums to zero, so
    dQ = dS K cancels to ~0.2 % of its terms (measured: |dQ| 1e-3 for |dO| 0.5) and carries the fp16 rounding of dS
    amplified -- 5.6e-2 relative on dQ with exactly the same kernel that gives 3e-4 on uncorrelated keys
    (tests/test_backward_gpu.py); flash-attn rounds dS to fp16 as well.  It is bounded against the size of the sibling
    gradient d to_kv instead."""
    from gvfdiffusion_b200.model.autoencoder import GSKLTemporalVariationalAutoEncoder as VAE
    g = torch.load(os.path.join(G, "vae_encode_tiny.pt"), weights_only=False)
    dec = torch.load(os.path.join(G, "vae_tiny.pt"), weights_only=False)
    enc_sd = {k: t.clone() for k, t in g["state_dict"].items()}
    if attn_std is not None:
        gen = torch.Generator().manual_seed(8)
        for k in ("cross_attend_blocks.0.fn.to_q.weight", "cross_attend_blocks.0.fn.to_kv.weight"):
            enc_sd[k] = torch.randn(enc_sd[k].shape, generator=gen) * attn_std
    v = VAE(**g["cfg"])
    v.load_state_dict({**dec["state_dict"], **enc_sd})
    v = v.to(DEV)
    R = torch.randn(g["noise"].shape, generator=torch.Generator().manual_seed(2))
    kl, x, post, _ = v.encode(g["static_pc"].to(DEV), g["delta_pc"].to(DEV), [t.to(DEV) for t in g["static_gs"]], noise=g["noise"])
    assert x.requires_grad and kl.requires_grad
    ((x * R.to(DEV)).sum() + 3.0 * kl.sum()).backward()
    ref = _oracle_train(enc_sd, g, g["cfg"]["heads"], g["cfg"], R, 3.0, g["noise"])
    named = dict(v.named_parameters())
    errs = {n: rel(named[n].grad, ref[n]) for n in v._enc_names}
    print(f"encode backward vs oracle autograd (attn std {attn_std}):",
          {k: f"{e:.2e}" for k, e in sorted(errs.items(), key=lambda kv: -kv[1])[:4]})
    tq, tkv = "cross_attend_blocks.0.fn.to_q.weight", "cross_attend_blocks.0.fn.to_kv.weight"
    cancel = float((named[tq].grad.cpu() - ref[tq]).norm() / ref[tkv].norm())
    print(f"  |d to_q| / |d to_kv| = {float(ref[tq].norm() / ref[tkv].norm()):.2e}, d to_q error / |d to_kv| = {cancel:.2e}")
    assert cancel < 2e-3
    errs.pop(tq)
    assert max(errs.values()) < 1e-2, errs
    assert all(named[n].grad is None for n in v._param_names)                # the decoder was not involved


def test_model_forward_is_differentiable_end_to_end():
    """model(static_gs, static_pc, delta_pc) = encode -> pad -> decode (model/autoencoder.py:620-627): one backward reaches
    encoder and decoder parameters (train_vae.py:293-353: `output = self.model(...)`, KL + reconstruction terms)."""
    from gvfdiffusion_b200.model.autoencoder import GSKLTemporalVariationalAutoEncoder as VAE
    g = torch.load(os.path.join(G, "vae_encode_tiny.pt"), weights_only=False)
    dec = torch.load(os.path.join(G, "vae_tiny.pt"), weights_only=False)
    v = VAE(**g["cfg"])
    v.load_state_dict({**dec["state_dict"], **g["state_dict"]})
    v = v.to(DEV)
    torch.manual_seed(0)
    out = v([t.to(DEV) for t in g["static_gs"]], g["static_pc"].to(DEV), g["delta_pc"].to(DEV))
    # a SUM, not a mean: fp16 activation gradients need the magnitude a GradScaler would give them (with a mean over
    # 4 200 outputs the attention's dS underflows fp16 and d to_q comes out exactly zero -- in the reference too)
    loss = out["logits"].square().sum() + out["kl"].sum()
    loss.backward()
    named = dict(v.named_parameters())
    for n in v._enc_names + v._param_names:
        assert named[n].grad is not None and torch.isfinite(named[n].grad).all(), n
    assert float(named["cross_attend_blocks.0.fn.to_q.weight"].grad.abs().max()) > 0


def test_engines_refresh_in_place_after_an_optimiser_step():
    """After optimiser steps the three engines (inference decode, training decode, encode) keep their device buffers and
    take the new values by device-side casts (`refresh`): forward + backward of the stepped model must equal a model built
    fresh from its state dict, and the weight buffers must not have moved."""
    from gvfdiffusion_b200.model.autoencoder import GSKLTemporalVariationalAutoEncoder as VAE
    g = torch.load(os.path.join(G, "vae_encode_tiny.pt"), weights_only=False)
    dec = torch.load(os.path.join(G, "vae_tiny.pt"), weights_only=False)
    v = VAE(**g["cfg"])
    v.load_state_dict({**dec["state_dict"], **g["state_dict"]})
    v = v.to(DEV)
    gs, pc, dpc = [t.to(DEV) for t in g["static_gs"]], g["static_pc"].to(DEV), g["delta_pc"].to(DEV)
    noise = torch.randn(pc.shape[0] * g["cfg"]["num_timesteps"], g["cfg"]["num_latents"], g["cfg"]["latent_dim"],
                        generator=torch.Generator().manual_seed(2)).to(DEV)
    # AdamW(fused=True) updates the parameters WITHOUT bumping their version counters: the training Functions refresh the
    # engines on every forward, the inference engine after train() / eval() switches
    opt = torch.optim.AdamW(v.parameters(), lr=1e-3, fused=True)
    ptrs = None
    for it in range(2):                                    # step twice: the second forward runs on refreshed engines
        out = v(gs, pc, dpc, noise=noise)
        (out["logits"].square().sum() + out["kl"].sum()).backward()
        with torch.no_grad():
            v.decode(out["posterior"]["mean"], torch.stack([t[:64] for t in gs]))      # builds the inference engine too
        now = (v.train_engine().layers[0]["w_qkv"].data_ptr(), v.train_engine().layers[0]["w1"].data_ptr(),
               v.encode_engine().w_kv.data_ptr(), v.engine().w_dkv.data_ptr())
        assert ptrs is None or now == ptrs
        ptrs = now
        opt.step()
        opt.zero_grad()
    fresh = VAE(**g["cfg"])
    fresh.load_state_dict(v.state_dict())
    fresh = fresh.to(DEV)
    v.eval()
    v.train()                                              # mode switch: the inference engine re-reads the weights too
    o1, o2 = v(gs, pc, dpc, noise=noise), fresh(gs, pc, dpc, noise=noise)
    assert torch.equal(o1["logits"], o2["logits"]) and torch.equal(o1["kl"], o2["kl"])
    (o1["logits"].square().sum() + o1["kl"].sum()).backward()
    (o2["logits"].square().sum() + o2["kl"].sum()).backward()
    a, b = dict(v.named_parameters()), dict(fresh.named_parameters())
    for n in v._enc_names + v._param_names:
        assert rel(a[n].grad, b[n].grad) < 1e-5, n         # split-K reduce-adds arrive in any order
    with torch.no_grad():
        q = torch.stack([t[:64] for t in gs])
        assert torch.equal(v.decode(o1["posterior"]["mean"], q), fresh.decode(o2["posterior"]["mean"], q))
