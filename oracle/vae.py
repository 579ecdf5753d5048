"""CPU oracle: motion ("variation field") VAE decode (TEST INFRASTRUCTURE, see
oracle/__init__.py).

Restates reference `model/autoencoder.py:579-609` (decode), `:552-577` (process_chunk),
`:109-163` (Attention), `:90-107` (GEGLU FeedForward), `:73-88` (PreNorm), `:250-301`
(PointEmbed) over a reference-keyed state dict, torch CPU.

precision="fp32": the module as it runs without autocast (PointEmbed's outer product is
promoted to float64 by the float64 `omega` buffer, then cast back).
precision="fp16": emulation of the fp16 autocast the reference decodes under
(`inference_dpm_latent.py:256-257`): Linear in/out fp16, LayerNorm fp32, the latent
residual stream is fp16 (Linear output + fp16 residual), GEGLU in fp16, and PointEmbed's
einsum runs in fp16 (einsum is on CUDA autocast's fp16 list), i.e. coordinates, omega
and their product are rounded to fp16 before sin/cos.

Pinned against the reference module by tests/golden/make_golden.py.
"""
import numpy as np
import torch
import torch.nn.functional as F

from .dit import _P


def _ln(x, eps):
    return F.layer_norm(x.float(), (x.shape[-1],), None, None, eps)


def point_embed(xyz, dim, P):
    # model/autoencoder.py:250-301 ; xyz (..., 3) fp32 -> (..., dim)
    e = dim // 3 // 2
    omega = np.arange(e, dtype=np.float64)
    omega /= e / 2.0
    omega = torch.from_numpy(1.0 / 10000 ** omega)
    outs = []
    for c in range(3):
        if P.h:
            arg = (xyz[..., c].half()[..., None] * omega.half()).float().half().float()
            outs += [torch.sin(arg).half().float(), torch.cos(arg).half().float()]
        else:
            arg = xyz[..., c].double()[..., None] * omega
            outs += [torch.sin(arg), torch.cos(arg)]
    return torch.cat(outs, dim=-1).to(torch.float32)


def attention(sd, pre, x, ctx, heads, P):
    # model/autoencoder.py:109-163 ; to_q / to_kv without bias, to_out with bias
    B, N, C = x.shape
    q = P.linear(x, sd[pre + "to_q.weight"])
    k, v = P.linear(ctx, sd[pre + "to_kv.weight"]).chunk(2, dim=-1)
    sp = lambda t: t.reshape(t.shape[0], t.shape[1], heads, -1).permute(0, 2, 1, 3)
    o = F.scaled_dot_product_attention(sp(q), sp(k), sp(v))      # scale = dim_head ** -0.5
    o = P.r(o.permute(0, 2, 1, 3).reshape(B, N, -1))
    return P.linear(o, sd[pre + "to_out.weight"], sd[pre + "to_out.bias"])


def feed_forward(sd, pre, x, P):
    h = P.linear(x, sd[pre + "net.0.weight"], sd[pre + "net.0.bias"])
    a, g = h.chunk(2, dim=-1)
    h = P.r(a * P.r(F.gelu(g)))
    return P.linear(h, sd[pre + "net.2.weight"], sd[pre + "net.2.bias"])


def latent_layers(sd, z, heads, P):
    """proj + depth x (PreNorm self-attn + PreNorm GEGLU FF); z ((B*T), L, latent_dim)."""
    depth = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("layers."))
    x = P.linear(z, sd["proj.weight"], sd["proj.bias"])
    for i in range(depth):
        n = _ln(x, 1e-6)
        x = P.r(attention(sd, f"layers.{i}.0.fn.", n, n, heads, P) + x)
        x = P.r(feed_forward(sd, f"layers.{i}.1.fn.", _ln(x, 1e-6), P) + x)
    return x


def query_embed(sd, queries, P):
    """gs_embedding(q) + position_encoding(q[..., :3]) -- model/autoencoder.py:389-391,560.
    Both LayerNorms are nn.LayerNorm defaults (eps 1e-5, no affine)."""
    dim = sd["gs_embedding.0.weight"].shape[0]
    g = _ln(P.linear(queries, sd["gs_embedding.0.weight"], sd["gs_embedding.0.bias"]), 1e-5)
    p = _ln(point_embed(queries[..., :3], dim, P), 1e-5)
    return g + p


def vae_decode(sd, z, queries, heads, num_timesteps, precision="fp16", chunk_size=8192):
    """z ((B*T), L, latent_dim), queries (B, Q, 14) -> (B, T, Q, out_dim) fp32."""
    P = _P(precision)
    sd = {k: v.float() for k, v in sd.items()}
    B, Q = queries.shape[:2]
    T = num_timesteps
    x = latent_layers(sd, z, heads, P)
    ctx = _ln(x, 1e-6)                                           # PreNorm.norm_context
    outs = []
    for s in range(0, Q, chunk_size):
        qe = query_embed(sd, queries[:, s:s + chunk_size], P)    # (B, q, dim), frame independent
        qe = qe.unsqueeze(1).expand(B, T, qe.shape[1], qe.shape[2]).reshape(B * T, -1, qe.shape[2])
        lat = attention(sd, "decoder_cross_attn.fn.", _ln(qe, 1e-6), ctx, heads, P)
        outs.append(P.linear(lat, sd["to_outputs.weight"], sd["to_outputs.bias"]))
    return torch.cat(outs, dim=1).reshape(B, T, Q, -1)


# ---------------------------------------------------------------------------------------------- encode
def fps_indices(points, K, start=0):
    """Greedy farthest point sampling from `start`, squared distances (dx*dx + dy*dy) + dz*dz in fp32, ties -> lowest
    index (the deterministic rule of gvf_fps; torch_cluster.fps, model/autoencoder.py:525, starts at a random point)."""
    p = points.numpy().astype(np.float32)
    mind = np.full(p.shape[0], 3.0e38, np.float32)
    cur, idx = start, [start]
    for _ in range(1, K):
        d = p - p[cur]
        dist = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        mind = np.minimum(mind, dist.astype(np.float32))
        cur = int(np.argmax(mind))
        idx.append(cur)
    return torch.from_numpy(np.array(idx, dtype=np.int64))


def token_embed(sd, disp, xyz, P):
    """input_embedding(disp) + position_encoding(xyz): LN_1e-5(Linear(3 -> dim)(disp)) + LN_1e-5(PointEmbed(xyz)), fp32
    (LayerNorm outputs are fp32 under autocast) -- model/autoencoder.py:386-391,529-533."""
    dim = sd["input_embedding.0.weight"].shape[0]
    a = _ln(P.linear(disp, sd["input_embedding.0.weight"], sd["input_embedding.0.bias"]), 1e-5)
    return a + _ln(point_embed(xyz, dim, P), 1e-5)


def vae_encode(sd, static_pc, delta_pc, static_gs_list, heads, num_latents, knn_k, beta, precision="fp16", noise=None):
    """Restates GSKLTemporalVariationalAutoEncoder.encode (model/autoencoder.py:502-550) and compute_delta_interp
    (:451-500).  -> dict(kl [(B T)], x [(B T), L, latent], mean, logvar, sampled_static_gs [B, L, 14])."""
    from . import losses as OL
    P = _P(precision)
    sd = {k: v.float() for k, v in sd.items()}
    B, N, _ = static_pc.shape
    T = delta_pc.shape[1]
    L = num_latents
    sampled = torch.stack([g[fps_indices(g[:, :3], L)] for g in static_gs_list])            # [B, L, 14]
    gs_xyz = sampled[:, :, :3].contiguous()
    moving = delta_pc + static_pc.unsqueeze(1)
    d, idx = OL.knn_points(gs_xyz, static_pc, K=knn_k)[:2]
    est = OL.interp_deltas(d, idx, static_pc, moving, torch.full((B,), L, dtype=torch.int64), True, beta)   # [B, T, L, 3]
    rep = lambda t: t.unsqueeze(1).expand(B, T, t.shape[1], t.shape[2])
    a = token_embed(sd, est, rep(gs_xyz), P).reshape(B * T, L, -1)                            # latent tokens
    ctx = token_embed(sd, delta_pc, rep(static_pc), P).reshape(B * T, N, -1)                  # point tokens
    pre0, pre1 = "cross_attend_blocks.0.fn.", "cross_attend_blocks.1.fn."
    x = attention(sd, pre0, _ln(a, 1e-6), _ln(ctx, 1e-6), heads, P) + a                       # fp32 residual stream
    x = feed_forward(sd, pre1, _ln(x, 1e-6), P) + x
    mean = P.linear(x, sd["mean_fc.weight"], sd["mean_fc.bias"])
    logvar = P.linear(x, sd["logvar_fc.weight"], sd["logvar_fc.bias"]).clamp(-30.0, 20.0)
    std, var = P.r(torch.exp(0.5 * logvar)), P.r(torch.exp(logvar))
    if noise is None:
        noise = torch.randn(mean.shape)
    sample = mean + std * noise
    kl = 0.5 * torch.mean(P.r(P.r(P.r(mean.pow(2)) + var - 1.0) - logvar), dim=[1, 2])
    return {"kl": kl, "x": sample, "mean": mean, "logvar": logvar, "sampled_static_gs": sampled}
