"""GPU parity: windowed sparse self-attention (gather / scatter fused into the kernel) against the CPU oracle
of the reference path sparse/attention/windowed_attn.py:61-135, through the C ABI."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _voxels(n_per_batch, res, batches, seed):
    g = torch.Generator().manual_seed(seed)
    out = []
    for b in range(batches):
        lin = torch.randperm(res ** 3, generator=g)[:n_per_batch]
        xyz = torch.stack([lin // (res * res), (lin // res) % res, lin % res], 1)
        out.append(torch.cat([torch.full((n_per_batch, 1), b), xyz], 1))
    return torch.cat(out).int()


@pytest.mark.parametrize("n,res,window,shift,H", [(1500, 64, 8, (0, 0, 0), 12), (1500, 64, 8, (4, 4, 4), 12),
                                                   (900, 16, 8, (0, 0, 0), 3),      # dense windows: up to 512 voxels, many key chunks
                                                   (70, 8, 8, (0, 0, 0), 2), (5, 32, 8, (4, 4, 4), 1)])
@pytest.mark.parametrize("packed", [False, True, "tma"])
def test_windowed_attention_matches_oracle(n, res, window, shift, H, packed):
    from gvfdiffusion_b200.sparse.attention import sparse_windowed_scaled_dot_product_self_attention
    from oracle import sparse_window as OSW
    coords = _voxels(n, res, 2, seed=n + H)
    g = torch.Generator().manual_seed(7)
    qkv = (torch.randn(coords.shape[0], 3, H, 64, generator=g) * 1.5).half()
    ref = OSW.windowed_attention(qkv, coords, window, shift)
    from gvfdiffusion_b200 import _lib
    _lib.lib().gvf_sparse_attn_set_tma(int(packed == "tma"))        # rows staged by TMA tile::gather4 instead of cp.async
    try:
        out = sparse_windowed_scaled_dot_product_self_attention(qkv.to(DEV), coords.to(DEV), window, shift, packed=bool(packed)).float().cpu()
    finally:
        _lib.lib().gvf_sparse_attn_set_tma(0)
    err = (out - ref).abs().max().item() / ref.abs().max().item()
    assert err < 2e-3, err       # packed: 64 consecutive sorted positions per CTA, rows masked to their own window


def test_windowed_attention_on_reference_partition_fixture():
    """Same voxel sets the reference's calc_window_partition was run on (tests/golden/window_partition.pt): the
    device partition must reproduce its seq_lens, and attention computed through the REFERENCE's fwd order must
    equal attention through ours (order inside a window is irrelevant)."""
    from gvfdiffusion_b200 import _lib
    from gvfdiffusion_b200._lib import check, current_stream, ptr
    from gvfdiffusion_b200.sparse.attention import calc_window_partition, sparse_windowed_scaled_dot_product_self_attention
    cases = torch.load(os.path.join(G, "window_partition.pt"), weights_only=False)
    for c in cases:
        coords = c["coords"].to(DEV)
        fwd, bwd, lens, batch = calc_window_partition(coords, c["window"], c["shift"])
        assert lens.tolist() == c["seq_lens"] and batch.tolist() == c["seq_batch_indices"]
        T, H = coords.shape[0], 4
        g = torch.Generator().manual_seed(T)
        qkv = torch.randn(T, 3, H, 64, generator=g).half().to(DEV)
        ours = sparse_windowed_scaled_dot_product_self_attention(qkv, coords, c["window"], c["shift"])
        cu = torch.zeros(len(c["seq_lens"]) + 1, dtype=torch.int32)
        cu[1:] = torch.cumsum(torch.tensor(c["seq_lens"]), 0)
        out = torch.empty((T, H, 64), dtype=torch.float16, device=DEV)
        fr = c["fwd"].int().to(DEV)
        check(_lib.lib().gvf_sparse_window_attn_f16(ptr(qkv), ptr(out), ptr(fr), ptr(cu.to(DEV)), len(c["seq_lens"]),
                                                    max(c["seq_lens"]), H, 64, 1.0 / math.sqrt(64), current_stream()),
              "gvf_sparse_window_attn_f16")
        assert (out.float() - ours.float()).abs().max().item() < 2e-3


def test_sparse_transformer_blocks_match_oracle():
    """Two swin blocks (unshifted + shifted windows) at the static-VAE width (768 = 12 heads x 64) on ~2 x 700
    voxels: device engine against the torch restatement of SparseTransformerBlock (fp16-emulating and fp32)."""
    from gvfdiffusion_b200.sparse.transformer import SparseTransformerBlocks
    from oracle import sparse_window as OSW
    g = torch.Generator().manual_seed(21)
    C, H, NB = 768, 12, 2
    sd = {}
    for i in range(NB):
        for name, (o, k) in {"attn.to_qkv": (3 * C, C), "attn.to_out": (C, C), "mlp.mlp.0": (4 * C, C), "mlp.mlp.2": (C, 4 * C)}.items():
            sd[f"blocks.{i}.{name}.weight"] = torch.randn(o, k, generator=g) * 0.03
            sd[f"blocks.{i}.{name}.bias"] = torch.randn(o, generator=g) * 0.05
    coords = _voxels(700, 32, 2, seed=3)
    feats = torch.randn(coords.shape[0], C, generator=g)
    eng = SparseTransformerBlocks(sd, "blocks.", NB, H, 8, device=DEV)
    y = eng.forward(feats.to(DEV), coords.to(DEV)).cpu()
    y16 = OSW.transformer_blocks(sd, "blocks.", NB, H, feats, coords, 8, "fp16")
    y32 = OSW.transformer_blocks(sd, "blocks.", NB, H, feats, coords, 8, "fp32")
    rel = lambda a, b: float((a - b).norm() / b.norm())
    assert rel(y, y16) < 1e-3, rel(y, y16)
    assert rel(y, y32) < 3e-3, rel(y, y32)


def test_sparse_vae_decode_matches_oracle():
    """SparseTransformerVAE.decode at the shipped widths (latent 8 -> 768, 12 heads, out 112, norm_output, fp16
    residual stream), 2 blocks, against the torch restatement."""
    from gvfdiffusion_b200.sparse.transformer import SparseTransformerVAE
    from oracle import sparse_window as OSW
    g = torch.Generator().manual_seed(33)
    C, H, NB = 768, 12, 2
    sd = {"from_latent.weight": torch.randn(C, 8, generator=g) * 0.3, "from_latent.bias": torch.randn(C, generator=g) * 0.1,
          "out_layer.weight": torch.randn(112, C, generator=g) * 0.03, "out_layer.bias": torch.randn(112, generator=g) * 0.1}
    for i in range(NB):
        for name, (o, k) in {"attn.to_qkv": (3 * C, C), "attn.to_out": (C, C), "mlp.mlp.0": (4 * C, C), "mlp.mlp.2": (C, 4 * C)}.items():
            sd[f"decoder.{i}.{name}.weight"] = torch.randn(o, k, generator=g) * 0.03
            sd[f"decoder.{i}.{name}.bias"] = torch.randn(o, generator=g) * 0.05
    coords = _voxels(600, 64, 2, seed=9)
    latent = torch.randn(coords.shape[0], 8, generator=g)
    vae = SparseTransformerVAE(sd, NB, H, 8, use_fp16=True, norm_output=True, device=DEV)
    y = vae.decode(latent.to(DEV), coords.to(DEV)).cpu()
    ref = OSW.vae_decode(sd, NB, H, latent, coords, 8, "fp16", use_fp16=True, norm_output=True)
    rel = float((y - ref).norm() / ref.norm())
    assert y.shape == (coords.shape[0], 112) and rel < 2e-3, rel


@pytest.mark.parametrize("n_vox,res,window,shift", [(1500, 64, 8, (0, 0, 0)), (1500, 64, 8, (4, 4, 4)), (300, 16, 4, (2, 2, 2)),
                                                     (4000, 32, 8, (4, 4, 4)), (64, 8, 8, (0, 0, 0)), (3000, 16, 8, (0, 0, 0))])
@pytest.mark.parametrize("packed", [False, True])
def test_windowed_attention_backward_matches_autograd(n_vox, res, window, shift, packed):
    """csrc/sparse_attn_bwd.cu (dq + dk/dv kernels, gather fused) against torch autograd of the reference formulation:
    gather by window, softmax attention inside every window, scatter back.  fp16 gradients: 5e-3 rel. L2."""
    from gvfdiffusion_b200.sparse.attention.windowed_attn import calc_window_partition, sparse_windowed_attention_autograd
    g = torch.Generator().manual_seed(n_vox + res)
    coords = []
    for b in range(2):
        lin = torch.randperm(res ** 3, generator=g)[:n_vox]
        coords.append(torch.cat([torch.full((n_vox, 1), b), torch.stack([lin // (res * res), (lin // res) % res, lin % res], 1)], 1))
    coords = torch.cat(coords).int().to(DEV)
    T, H, C = coords.shape[0], 3, 64
    qkv = (torch.randn(T, 3, H, C, generator=g) * 0.8).half().to(DEV).requires_grad_(True)
    dout = (torch.randn(T, H, C, generator=g) * 0.5).half().to(DEV)
    out = sparse_windowed_attention_autograd(qkv, coords, window, shift, packed=packed)
    out.backward(dout)
    fwd, bwd, seq_lens, _ = calc_window_partition(coords, window, shift)
    qf = qkv.detach().float().requires_grad_(True)
    gq = qf[fwd]
    outs, s0 = [], 0
    for n in seq_lens.tolist():
        q_, k_, v_ = (gq[s0:s0 + n, i].transpose(0, 1)[None] for i in range(3))
        outs.append(F.scaled_dot_product_attention(q_, k_, v_)[0].transpose(0, 1))
        s0 += n
    ref = torch.cat(outs)[bwd]
    ref.backward(dout.float())
    rel = lambda a, b: float((a.float() - b.float()).norm() / b.float().norm())
    e_o, e_g = rel(out, ref), [rel(qkv.grad[:, i], qf.grad[:, i]) for i in range(3)]
    print(f"windowed attn bwd T={T} windows={len(seq_lens)} max={int(seq_lens.max())}: out {e_o:.2e} dq {e_g[0]:.2e} dk {e_g[1]:.2e} dv {e_g[2]:.2e}")
    assert e_o < 2e-3 and max(e_g) < 5e-3, (e_o, e_g)


@pytest.mark.parametrize("use_fp16,norm_output,old_impl", [(True, True, False), (False, False, False), (True, False, True)])
def test_sparse_vae_decode_backward_matches_autograd(use_fp16, norm_output, old_impl):
    """SparseTransformerVAE.decode_train / decode_backward (from_latent + APE -> swin blocks -> LayerNorm -> out_layer, all
    gradients on the library's kernels) against torch autograd of the oracle restatement in its fp16-emulating regime, at
    the shipped widths (768 channels, 12 heads, window 8, out 112), 2 blocks.  fp16 activation gradients: 1e-2 rel. L2."""
    from gvfdiffusion_b200 import ops
    from gvfdiffusion_b200.sparse.transformer import SparseTransformerVAE
    from oracle import sparse_window as OSW
    g = torch.Generator().manual_seed(35)
    C, H, NB = 768, 12, 2
    sd = {"from_latent.weight": torch.randn(C, 8, generator=g) * 0.3, "from_latent.bias": torch.randn(C, generator=g) * 0.1,
          "out_layer.weight": torch.randn(112, C, generator=g) * 0.03, "out_layer.bias": torch.randn(112, generator=g) * 0.1}
    for i in range(NB):
        for name, (o, k) in {"attn.to_qkv": (3 * C, C), "attn.to_out": (C, C), "mlp.mlp.0": (4 * C, C), "mlp.mlp.2": (C, 4 * C)}.items():
            sd[f"decoder.{i}.{name}.weight"] = torch.randn(o, k, generator=g) * 0.03
            sd[f"decoder.{i}.{name}.bias"] = torch.randn(o, generator=g) * 0.05
    sd = {k: v.half().float() for k, v in sd.items()}
    coords = _voxels(700, 64, 2, seed=11)
    T = coords.shape[0]
    latent = torch.randn(T, 8, generator=g)
    dout = torch.randn(T, 112, generator=g)
    vae = SparseTransformerVAE(sd, NB, H, 8, use_fp16=use_fp16, norm_output=norm_output, device=DEV, use_old_attn_impl=old_impl)
    y, saved = vae.decode_train(latent.to(DEV), coords.to(DEV))
    y_inf = vae.decode(latent.to(DEV), coords.to(DEV))
    grads, dlat = vae.decode_backward(saved, dout.to(DEV))
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    lr = latent.clone().requires_grad_(True)
    ref = OSW.vae_decode(sdr, NB, H, lr, coords, 8, "fp16", use_fp16=use_fp16, norm_output=norm_output, old_attn_impl=old_impl)
    (ref * dout).sum().backward()
    rel = lambda a, b: float((a.float().cpu() - b).norm() / b.norm().clamp_min(1e-20))
    assert rel(y, ref.detach()) < 2e-3 and rel(y, y_inf.cpu()) < 2e-3
    assert set(grads) == set(sd)
    errs = {k: rel(grads[k], sdr[k].grad) for k in sd}
    worst = max(errs.items(), key=lambda kv: kv[1])
    assert worst[1] < 1e-2, worst
    assert rel(dlat, lr.grad) < 1e-2, rel(dlat, lr.grad)


def test_gelu_tanh_forward_backward():
    from gvfdiffusion_b200 import ops
    g = torch.Generator().manual_seed(3)
    h = (torch.randn(1000, 512, generator=g) * 2.5).half().to(DEV)
    dy = torch.randn(1000, 512, generator=g).half().to(DEV)
    hr = h.float().requires_grad_(True)
    ref = torch.nn.functional.gelu(hr, approximate="tanh")
    ref.backward(dy.float())
    assert (ops.gelu_tanh(h).float() - ref.detach()).abs().max() < 4e-3          # tanh.approx + fp16 rounding
    assert (ops.gelu_tanh_bwd(h, dy).float() - hr.grad).abs().max() < 8e-3


def test_sparse_vae_full_train_step_gradients_match_autograd():
    """SparseTransformerVAE.forward (encode -> posterior sample -> decode, sparse_transformer_vae.py:206-210) as one autograd
    node: out and kl forward, and every parameter gradient of loss = <out, w> + 0.3 kl, against torch autograd of the oracle
    restatement (fp16-emulating regime) at the shipped widths (in 1024 -> 768, 12 heads, latent 8, out 112), 2 + 2 blocks."""
    from gvfdiffusion_b200.sparse.transformer import SparseTransformerVAE, sparse_vae_forward_autograd
    from oracle import sparse_window as OSW
    from oracle.dit import _P, absolute_position_embedding
    g = torch.Generator().manual_seed(36)
    C, H, NB, CIN = 768, 12, 2, 1024
    sd = {"input_layer.weight": torch.randn(C, CIN, generator=g) * 0.03, "input_layer.bias": torch.randn(C, generator=g) * 0.1,
          "to_latent.weight": torch.randn(16, C, generator=g) * 0.03, "to_latent.bias": torch.randn(16, generator=g) * 0.1,
          "from_latent.weight": torch.randn(C, 8, generator=g) * 0.3, "from_latent.bias": torch.randn(C, generator=g) * 0.1,
          "out_layer.weight": torch.randn(112, C, generator=g) * 0.03, "out_layer.bias": torch.randn(112, generator=g) * 0.1}
    for side in ("encoder", "decoder"):
        for i in range(NB):
            for name, (o, k) in {"attn.to_qkv": (3 * C, C), "attn.to_out": (C, C), "mlp.mlp.0": (4 * C, C), "mlp.mlp.2": (C, 4 * C)}.items():
                sd[f"{side}.{i}.{name}.weight"] = torch.randn(o, k, generator=g) * 0.03
                sd[f"{side}.{i}.{name}.bias"] = torch.randn(o, generator=g) * 0.05
    sd = {k: v.half().float() for k, v in sd.items()}
    coords = _voxels(600, 64, 2, seed=12)
    T = coords.shape[0]
    feats = torch.randn(T, CIN, generator=g)
    noise = torch.randn(T, 8, generator=g)
    dout = torch.randn(T, 112, generator=g)
    vae = SparseTransformerVAE(sd, NB, H, 8, use_fp16=True, norm_output=True, device=DEV)
    out, kl, mean, logvar = sparse_vae_forward_autograd(vae, feats.to(DEV), coords.to(DEV), noise)
    ((out * dout.to(DEV)).sum() + 0.3 * kl).backward()
    grads = vae.grads
    # oracle: the same function on the CPU under torch autograd
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    P = _P("fp16")
    h = P.linear(feats, sdr["input_layer.weight"], sdr["input_layer.bias"])
    h = h + absolute_position_embedding(coords[:, 1:].float()[None], C)[0]
    h = OSW.transformer_blocks(sdr, "encoder.", NB, H, h, coords, 8, "fp16", fp16_residual=True)
    ml = P.linear(F.layer_norm(h, (C,)), sdr["to_latent.weight"], sdr["to_latent.bias"])
    m_r, lv_r = ml.chunk(2, dim=-1)
    z = m_r + torch.exp(0.5 * lv_r) * noise
    ref = OSW.vae_decode(sdr, NB, H, z, coords, 8, "fp16", use_fp16=True, norm_output=True)
    kl_r = 0.5 * torch.mean(m_r.pow(2) + lv_r.exp() - lv_r - 1)
    ((ref * dout).sum() + 0.3 * kl_r).backward()
    rel = lambda a, b: float((a.detach().float().cpu() - b.detach()).norm() / b.detach().norm().clamp_min(1e-20))
    assert rel(mean, m_r) < 3e-3 and rel(logvar, lv_r) < 3e-3 and rel(out, ref) < 3e-3
    assert abs(float(kl.detach()) - float(kl_r.detach())) < 2e-3 * abs(float(kl_r.detach()))
    assert set(grads) == set(sd)
    errs = {k: rel(grads[k], sdr[k].grad) for k in sd}
    worst = max(errs.items(), key=lambda kv: kv[1])
    assert worst[1] < 1.5e-2, sorted(errs.items(), key=lambda kv: -kv[1])[:5]


def test_sparse_transformer_vae_module_trains_and_refreshes_in_place():
    """The nn.Module mirror (reference constructor / parameter names): autograd delivers the same parameter gradients as
    the engine-level step; after an optimiser step the engine's fp16 copies and transposes are refreshed IN PLACE and the
    next forward equals a freshly built engine on the new weights."""
    from gvfdiffusion_b200.model.sparse_voxel_diffusion import SparseTransformerVAE
    from gvfdiffusion_b200.sparse.basic import SparseTensor
    from gvfdiffusion_b200.sparse.transformer import SparseTransformerVAE as Engine
    torch.manual_seed(5)
    m = SparseTransformerVAE(64, 64, 128, 24, 8, 2, window_size=8, use_fp16=True, use_old_attn_impl=False, norm_output=True).to(DEV)
    assert "encoder.1.mlp.mlp.2.weight" in m.state_dict() and "to_latent.bias" in m.state_dict()
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.startswith(("to_latent", "out_layer")) and n.endswith("weight"):
                p.normal_(0, 0.05)
    coords = _voxels(400, 32, 2, seed=3).to(DEV)
    g = torch.Generator().manual_seed(8)
    x = SparseTensor(torch.randn(800, 64, generator=g).to(DEV), coords)
    noise, dout = torch.randn(800, 8, generator=g).to(DEV), torch.randn(800, 24, generator=g).to(DEV)
    out, mean, logvar = m(x, noise=noise)
    ((out.feats * dout).sum() + 0.3 * m.kl).backward()
    eng = Engine({k: v.detach() for k, v in m.state_dict().items()}, 2, 2, 8, use_fp16=True, norm_output=True, device=DEV)
    o2, _, _, kl2, saved = eng.forward_train(x.feats, coords, noise)
    g2 = eng.backward(saved, dout, torch.tensor(0.3, device=DEV))
    assert torch.equal(out.feats, o2)
    rl = lambda a, b: float((a - b).norm() / b.norm().clamp_min(1e-20))
    for n, p in m.named_parameters():          # split-K partial tiles are summed by TMA reduce-add in arrival order
        assert p.grad is not None and rl(p.grad, g2[n]) < 1e-5, n
    ptr0 = m.engine().decoder.blocks[0]["w_qkv"].data_ptr()
    opt = torch.optim.AdamW(m.parameters(), lr=0.01, fused=True)     # fused: no version bump on the parameters
    opt.step()
    out3, _, _ = m(x, noise=noise)                                     # refreshes the engine in place
    assert m.engine().decoder.blocks[0]["w_qkv"].data_ptr() == ptr0
    eng3 = Engine({k: v.detach() for k, v in m.state_dict().items()}, 2, 2, 8, use_fp16=True, norm_output=True, device=DEV)
    o4, _, _, _, sv4 = eng3.forward_train(x.feats, coords, noise)
    assert torch.equal(out3.feats, o4) and not torch.equal(out3.feats, out.feats)
    (out3.feats * dout).sum().backward()                                # transposes refreshed too: same gradients
    g4 = eng3.backward(sv4, dout, None)
    p = dict(m.named_parameters())["decoder.0.attn.to_qkv.weight"]
    assert rl(p.grad - g2["decoder.0.attn.to_qkv.weight"], g4["decoder.0.attn.to_qkv.weight"]) < 1e-4
