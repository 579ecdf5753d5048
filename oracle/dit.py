"""CPU oracle: the temporal DiT denoiser (TEST INFRASTRUCTURE, see oracle/__init__.py).

Functional torch-CPU restatement of reference `model/dit.py:449-480` (DiT._forward),
`:227-278` (ModulatedSparseTransformerCrossBlock._forward), `:287-303` (FinalLayer),
`:16-56` (AbsolutePositionEmbedder), `:59-100` (TimestepEmbedder) and
`model/attention/modules.py:8-15,112-146` (MultiHeadRMSNorm, MultiHeadAttention.forward)
over a reference-keyed state dict.

precision="fp32": what the reference computes on CPU with ATTN_BACKEND=sdpa.
precision="fp16": emulates the fp16 autocast the reference runs under on GPU
(`inference_dpm_latent.py:122-125`): every nn.Linear rounds its inputs, weights and
output to fp16 (fp32 accumulate), LayerNorm / softmax statistics / the residual stream
stay fp32, attention inputs/outputs are fp16.

Pinned against the reference module itself by tests/golden/make_golden.py.
"""
import math

import torch
import torch.nn.functional as F


class _P:
    """Precision policy."""

    def __init__(self, precision):
        assert precision in ("fp32", "fp16")
        self.h = precision == "fp16"

    def r(self, x):  # round-trip through fp16 when emulating autocast
        return x.half().float() if self.h else x

    def linear(self, x, w, b=None):
        if self.h:
            y = F.linear(x.half().float(), w.half().float(), None if b is None else b.half().float())
            return y.half().float()
        return F.linear(x, w, b)


def timestep_embedding(t, dim=256, max_period=10000):
    # model/dit.py:73-95  (cos first, then sin)
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def absolute_position_embedding(xyz, channels):
    # model/dit.py:16-56: freq_dim = C//3//2; per coordinate [sin(freq_dim), cos(freq_dim)];
    # zero-padded up to `channels`.
    B, L, D = xyz.shape
    freq_dim = channels // D // 2
    freqs = torch.arange(freq_dim, dtype=torch.float32) / freq_dim
    freqs = 1.0 / (10000 ** freqs)
    out = torch.outer(xyz.reshape(-1), freqs)
    out = torch.cat([torch.sin(out), torch.cos(out)], dim=-1).reshape(B * L, -1)
    if out.shape[1] < channels:
        out = torch.cat([out, torch.zeros(B * L, channels - out.shape[1])], dim=-1)
    return out.reshape(B, L, -1)


def layer_norm(x, w=None, b=None, eps=1e-6):
    return F.layer_norm(x.float(), (x.shape[-1],), w, b, eps)


def rms_norm_heads(x, gamma, P):
    # model/attention/modules.py:8-15 ; x [N, L, H, d]
    d = x.shape[-1]
    return P.r(F.normalize(x.float(), dim=-1) * gamma * (d ** 0.5))


def sdpa(q, k, v, P):
    # model/attention/full_attn.py:74-140 ; [N, L, H, d] in / out, scale 1/sqrt(d)
    q, k, v = (t.permute(0, 2, 1, 3) for t in (q, k, v))
    o = F.scaled_dot_product_attention(q, k, v)
    return P.r(o.permute(0, 2, 1, 3))


def mha_self(sd, pre, x, H, P, qk_rms_norm=True):
    # model/attention/modules.py:112-130
    B, L, C = x.shape
    qkv = P.linear(x, sd[pre + "to_qkv.weight"], sd[pre + "to_qkv.bias"]).reshape(B, L, 3, H, -1)
    q, k, v = qkv.unbind(dim=2)
    if qk_rms_norm:
        q = rms_norm_heads(q, sd[pre + "q_rms_norm.gamma"], P)
        k = rms_norm_heads(k, sd[pre + "k_rms_norm.gamma"], P)
    h = sdpa(q, k, v, P).reshape(B, L, -1)
    return P.linear(h, sd[pre + "to_out.weight"], sd[pre + "to_out.bias"])


def mha_cross(sd, pre, x, ctx, H, P, qk_rms_norm=False):
    # model/attention/modules.py:131-146
    B, L, C = x.shape
    Lk = ctx.shape[1]
    q = P.linear(x, sd[pre + "to_q.weight"], sd[pre + "to_q.bias"]).reshape(B, L, H, -1)
    kv = P.linear(ctx, sd[pre + "to_kv.weight"], sd[pre + "to_kv.bias"]).reshape(B, Lk, 2, H, -1)
    k, v = kv.unbind(dim=2)
    if qk_rms_norm:
        q = rms_norm_heads(q, sd[pre + "q_rms_norm.gamma"], P)
        k = rms_norm_heads(k, sd[pre + "k_rms_norm.gamma"], P)
    h = sdpa(q, k, v, P).reshape(B, L, -1)
    return P.linear(h, sd[pre + "to_out.weight"], sd[pre + "to_out.bias"])


def _bc(v):  # (B, C) -> (B, 1, 1, C)
    return v.unsqueeze(1).unsqueeze(1)


def block_forward(sd, pre, x, mod, image_emb, static_emb, H, P, qk_rms_norm=True,
                  qk_rms_norm_cross=False):
    """model/dit.py:227-278.  x (B,T,N,C) fp32 residual; mod (B,C); image_emb (B,T,L,C);
    static_emb (B,Ls,C) (the reference repeats it over T, `:465`)."""
    B, T, N, C = x.shape
    smod = P.r(F.silu(mod))
    m6 = P.linear(smod, sd[pre + "adaLN_modulation.1.weight"], sd[pre + "adaLN_modulation.1.bias"])
    sh_s, sc_s, g_s, sh_m, sc_m, g_m = m6.chunk(6, dim=1)
    m3 = P.linear(smod, sd[pre + "adaLN_modulation_temporal.1.weight"],
                  sd[pre + "adaLN_modulation_temporal.1.bias"])
    sh_t, sc_t, g_t = m3.chunk(3, dim=1)

    # spatial self-attention over (B*T) sequences of N tokens
    h = layer_norm(x) * (1 + _bc(sc_s)) + _bc(sh_s)
    h = mha_self(sd, pre + "spatial_self_attn.", h.reshape(B * T, N, C), H, P, qk_rms_norm)
    x = x + P.r(h.reshape(B, T, N, C) * _bc(g_s))

    # temporal self-attention over (B*N) sequences of T tokens
    h = layer_norm(x) * (1 + _bc(sc_t)) + _bc(sh_t)
    h = h.transpose(1, 2).reshape(B * N, T, C)
    h = mha_self(sd, pre + "temporal_self_attn.", h, H, P, qk_rms_norm)
    h = h.reshape(B, N, T, C).transpose(1, 2)
    x = x + P.r(h * _bc(g_t))

    # image cross-attention (no gate, affine LN)
    h = layer_norm(x, sd[pre + "norm3.weight"], sd[pre + "norm3.bias"])
    h = mha_cross(sd, pre + "image_cross_attn.", h.reshape(B * T, N, C),
                  image_emb.reshape(B * T, -1, C), H, P, qk_rms_norm_cross)
    x = x + h.reshape(B, T, N, C)

    # static cross-attention (same context for every frame)
    h = layer_norm(x, sd[pre + "norm4.weight"], sd[pre + "norm4.bias"])
    ctx = static_emb.unsqueeze(1).expand(B, T, static_emb.shape[1], C).reshape(B * T, -1, C)
    h = mha_cross(sd, pre + "static_cross_attn.", h.reshape(B * T, N, C), ctx, H, P,
                  qk_rms_norm_cross)
    x = x + h.reshape(B, T, N, C)

    # MLP
    h = layer_norm(x) * (1 + _bc(sc_m)) + _bc(sh_m)
    h = P.linear(h, sd[pre + "mlp.mlp.0.weight"], sd[pre + "mlp.mlp.0.bias"])
    h = P.r(F.gelu(h, approximate="tanh"))
    h = P.linear(h, sd[pre + "mlp.mlp.2.weight"], sd[pre + "mlp.mlp.2.bias"])
    x = x + P.r(h * _bc(g_m))
    return x


def dit_forward(sd, x, t, cond_images, static_latent, deformation_position_xyz,
                num_heads, precision="fp32", qk_rms_norm=True, qk_rms_norm_cross=False):
    """model/dit.py:449-480.  x (B,T,N,Cin) fp32, t (B,) model time (0..1000),
    cond_images (B,T,L,Ci), static_latent (B,Ls,Cs), xyz (B,N,3) -> (B,T,N,Cout)."""
    P = _P(precision)
    sd = {k: v.float() for k, v in sd.items()}
    C = sd["input_layer.weight"].shape[0]
    num_blocks = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))
    h = P.linear(x, sd["input_layer.weight"], sd["input_layer.bias"])
    te = timestep_embedding(t, sd["t_embedder.mlp.0.weight"].shape[1])
    te = P.linear(te, sd["t_embedder.mlp.0.weight"], sd["t_embedder.mlp.0.bias"])
    te = P.r(F.silu(te))
    t_emb = P.linear(te, sd["t_embedder.mlp.2.weight"], sd["t_embedder.mlp.2.bias"])
    image_emb = P.linear(cond_images, sd["image_cond_proj.weight"], sd["image_cond_proj.bias"])
    static_emb = P.linear(static_latent, sd["static_cond_proj.weight"], sd["static_cond_proj.bias"])
    h = h + absolute_position_embedding(deformation_position_xyz, C).unsqueeze(1)
    for i in range(num_blocks):
        h = block_forward(sd, f"blocks.{i}.", h, t_emb, image_emb, static_emb, num_heads, P,
                          qk_rms_norm, qk_rms_norm_cross)
    # FinalLayer, model/dit.py:298-303
    smod = P.r(F.silu(t_emb))
    m2 = P.linear(smod, sd["final_layer.adaLN_modulation.1.weight"],
                  sd["final_layer.adaLN_modulation.1.bias"])
    shift, scale = m2.chunk(2, dim=1)
    h = layer_norm(h) * (1 + _bc(scale)) + _bc(shift)
    return P.linear(h, sd["final_layer.linear.weight"], sd["final_layer.linear.bias"])
