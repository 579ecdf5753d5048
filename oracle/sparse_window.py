"""ORACLE (test infrastructure, never on the product path): CPU restatement of the reference's windowed
sparse self-attention -- calc_window_partition (sparse/attention/windowed_attn.py:20-58) and the
gather -> per-window softmax attention -> scatter of :92-129 (flash_attn_varlen_qkvpacked_func is plain
scaled-dot-product attention inside each cu_seqlens segment).  Pinned by tests/golden/window_partition.pt,
produced by the reference's own calc_window_partition, and -- blocks, encode and decode trunks, both qkv channel
layouts -- by tests/golden/sparse_vae_tiny.pt, produced by the reference's own SparseTransformerVAE on the CPU
(tests/golden/make_golden.py gen_sparse_vae; tests/test_sparse_vae_cpu.py)."""
import math

import torch


def calc_window_partition(coords, window_size, shift_window=0):
    """coords [T, 1+DIM] int; returns (window id per voxel [T] int64, seq_lens list, seq_batch_indices list)
    -- the quantities of :39-56 that do not depend on argsort's tie order."""
    DIM = coords.shape[1] - 1
    shift = (shift_window,) * DIM if isinstance(shift_window, int) else tuple(shift_window)
    win = (window_size,) * DIM if isinstance(window_size, int) else tuple(window_size)
    sc = coords.clone().long()
    sc[:, 1:] += torch.tensor(shift)[None]
    max_coords = sc[:, 1:].max(dim=0).values.tolist()
    nwin = [math.ceil((mc + 1) / ws) for mc, ws in zip(max_coords, win)]
    offset = torch.cumprod(torch.tensor([1] + nwin[::-1]), dim=0).tolist()[::-1]
    sc[:, 1:] //= torch.tensor(win)[None]
    ids = (sc * torch.tensor(offset)[None]).sum(dim=1)
    counts = torch.bincount(ids)
    batch = torch.arange(counts.shape[0]) // offset[0]
    mask = counts != 0
    return ids, counts[mask].tolist(), batch[mask].tolist()


def windowed_attention(qkv_feats, coords, window_size, shift_window=0):
    """qkv_feats [T,3,H,C] -> [T,H,C] fp32: softmax(q k^T / sqrt(C)) v inside every window."""
    ids, _, _ = calc_window_partition(coords, window_size, shift_window)
    q, k, v = qkv_feats.float().unbind(1)
    out = torch.zeros_like(q)
    scale = 1.0 / math.sqrt(q.shape[-1])
    for wid in torch.unique(ids).tolist():
        sel = (ids == wid).nonzero().squeeze(1)
        s = torch.einsum("qhc,khc->hqk", q[sel], k[sel]) * scale
        out[sel] = torch.einsum("hqk,khc->qhc", s.softmax(-1), v[sel])
    return out


def transformer_blocks(sd, prefix, num_blocks, num_heads, feats, coords, window_size, precision="fp32",
                       fp16_residual=False, old_attn_impl=False):
    """Stack of un-modulated SparseTransformerBlock (reference model/sparse_voxel_diffusion/sparse_transformer.py
    :126-192 with modulated=False, attn_mode "swin": block i uses shift_window = window_size // 2 * (i % 2),
    :24-25) over a reference-keyed state dict: `{prefix}{i}.attn.to_qkv|to_out`, `{prefix}{i}.mlp.mlp.0|2`.
    feats [T, C] fp32, coords [T, 4] int -> [T, C] fp32.  precision="fp16" emulates the autocast regime
    (Linear / attention in fp16, LayerNorm and the residual stream in fp32), as oracle/dit.py does."""
    import torch.nn.functional as F
    from .dit import _P
    P = _P(precision)
    rr = (lambda t: t.half().float()) if fp16_residual else (lambda t: t)     # x.type(fp16) residual stream
    x = rr(feats.float())
    C = x.shape[1]
    for i in range(num_blocks):
        p = f"{prefix}{i}."
        shift = window_size // 2 * (i % 2)
        h = F.layer_norm(x, (C,), None, None, 1e-6)
        qkv = P.linear(h, sd[p + "attn.to_qkv.weight"], sd[p + "attn.to_qkv.bias"])
        if old_attn_impl:     # sparse/attention/modules.py:161-164: channels [H][3][d] (use_old_attn_impl=True)
            qkv = qkv.reshape(-1, num_heads, 3, C // num_heads).permute(0, 2, 1, 3)
        else:                 # :166: channels [3][H][d] (the shipped configs)
            qkv = qkv.reshape(-1, 3, num_heads, C // num_heads)
        a = P.r(windowed_attention(qkv, coords, window_size, shift)).reshape(-1, C)
        x = rr(x + P.linear(a, sd[p + "attn.to_out.weight"], sd[p + "attn.to_out.bias"]))
        h = F.layer_norm(x, (C,), None, None, 1e-6)
        h = P.linear(h, sd[p + "mlp.mlp.0.weight"], sd[p + "mlp.mlp.0.bias"])
        h = P.r(F.gelu(h, approximate="tanh"))
        x = rr(x + P.linear(h, sd[p + "mlp.mlp.2.weight"], sd[p + "mlp.mlp.2.bias"]))
    return x


def vae_decode(sd, num_blocks, num_heads, latent, coords, window_size=8, precision="fp16", use_fp16=True,
               norm_output=False, old_attn_impl=False):
    """SparseTransformerVAE.decode (sparse_transformer_vae.py:178-188): from_latent + APE of the voxel
    coordinates (sparse_transformer.py:62-109) -> decoder blocks -> optional layer_norm (eps 1e-5) -> out_layer."""
    import torch.nn.functional as F
    from .dit import _P, absolute_position_embedding
    P = _P(precision)
    C = sd["from_latent.weight"].shape[0]
    h = P.linear(latent.float(), sd["from_latent.weight"], sd["from_latent.bias"])
    h = h + absolute_position_embedding(coords[:, 1:].float()[None], C)[0]
    h = transformer_blocks(sd, "decoder.", num_blocks, num_heads, h, coords, window_size, precision, fp16_residual=use_fp16,
                           old_attn_impl=old_attn_impl)
    if norm_output:
        h = F.layer_norm(h, (C,))
    return P.linear(h, sd["out_layer.weight"], sd["out_layer.bias"])
