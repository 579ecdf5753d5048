"""TEST INFRASTRUCTURE -- CPU restatement (plain torch, fp32) of the TRELLIS structured-latent flow model, the denoiser of
the stage in front of the GVF path.  Only tests/ may import this; the product never does.

Follows /root/reference, function by function:
  SLatFlowModel.forward                      trellis/models/structured_latent_flow.py:231-262
  SparseResBlock3d.forward                   trellis/models/structured_latent_flow.py:53-66
  TimestepEmbedder                           trellis/models/sparse_structure_flow.py:11-52
  AbsolutePositionEmbedder                   trellis/modules/transformer/blocks.py:8-46
  ModulatedSparseTransformerCrossBlock       trellis/modules/sparse/transformer/modulated.py:147-166
  SparseMultiHeadAttention (self / cross)    trellis/modules/sparse/attention/modules.py:105-139
  SparseMultiHeadRMSNorm                     trellis/modules/sparse/attention/modules.py:12-25
  SparseFeedForwardNet                       trellis/modules/sparse/transformer/blocks.py:11-21
  SparseDownsample / SparseUpsample          trellis/modules/sparse/spatial.py:13-80
  SparseConv3d (spconv.SubMConv3d)           trellis/modules/sparse/conv/conv_spconv.py:6-20 -> oracle.sparse_vae.subm_conv3d

Pinned: tests/golden/slat_flow_tiny.pt holds the output of the reference's own SLatFlowModel class on seeded inputs
(tests/golden/make_golden.py::gen_slat_flow; spconv's convolution, absent here, is the dense cross-correlation at the
active sites in that run as in this file -- that one operator's arithmetic is "parity unpinned").
"""
import math

import torch
import torch.nn.functional as F

from . import sparse_vae as OSV


def timestep_embedding(t, dim=256, max_period=10000):
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def t_embedder(sd, t):
    h = F.linear(timestep_embedding(t), sd["t_embedder.mlp.0.weight"], sd["t_embedder.mlp.0.bias"])
    return F.linear(F.silu(h), sd["t_embedder.mlp.2.weight"], sd["t_embedder.mlp.2.bias"])


def ape(xyz, channels):
    """xyz [N, 3] (integer voxel coordinates) -> [N, channels]: per coordinate [sin | cos] of freq_dim = channels // 6
    frequencies, zero padded."""
    fd = channels // 3 // 2
    freqs = 1.0 / (10000 ** (torch.arange(fd, dtype=torch.float32) / fd))
    out = torch.outer(xyz.reshape(-1).float(), freqs)
    out = torch.cat([torch.sin(out), torch.cos(out)], dim=-1).reshape(xyz.shape[0], -1)
    if out.shape[1] < channels:
        out = torch.cat([out, torch.zeros(xyz.shape[0], channels - out.shape[1])], dim=-1)
    return out


def downsample(feats, coords, factor=2):
    """-> (coarse feats, coarse coords [cells, 4] sorted lexicographically, idx [N] cell of every fine row).  The mean is
    sum / (count + 1): scatter_reduce(zeros, 'mean') counts its initial zero (include_self default, spatial.py:33-39)."""
    c = coords.long().clone()
    c[:, 1:] //= factor
    M = int(c[:, 1:].max()) + 1
    code = ((c[:, 0] * M + c[:, 1]) * M + c[:, 2]) * M + c[:, 3]
    ucode, idx = code.unique(return_inverse=True)
    summed = torch.zeros(ucode.shape[0], feats.shape[1], dtype=feats.dtype).index_add_(0, idx, feats)
    cnt = torch.bincount(idx, minlength=ucode.shape[0]).to(feats.dtype)
    new_coords = torch.stack([ucode // M ** 3, (ucode // M ** 2) % M, (ucode // M) % M, ucode % M], -1)
    return summed / (cnt + 1)[:, None], new_coords.int(), idx


def _conv(sd, prefix, feats, coords, batch_size):
    grid = int(coords[:, 1:].max()) + 1
    return OSV.subm_conv3d(feats, coords, sd[prefix + "conv.weight"], sd.get(prefix + "conv.bias"), batch_size, grid)


def res_block(sd, prefix, feats, coords, emb, batch_size):
    """SparseResBlock3d.forward after `_updown` (the caller resamples): feats [N, C] at `coords`, emb [B, Cm]."""
    b = coords[:, 0].long()
    emb_out = F.linear(F.silu(emb), sd[prefix + "emb_layers.1.weight"], sd[prefix + "emb_layers.1.bias"])
    scale, shift = torch.chunk(emb_out, 2, dim=1)
    C = feats.shape[1]
    h = F.layer_norm(feats, (C,), sd[prefix + "norm1.weight"], sd[prefix + "norm1.bias"], 1e-6)
    h = _conv(sd, prefix + "conv1.", F.silu(h), coords, batch_size)
    h = F.layer_norm(h, (h.shape[1],), None, None, 1e-6) * (1 + scale[b]) + shift[b]
    h = _conv(sd, prefix + "conv2.", F.silu(h), coords, batch_size)
    if prefix + "skip_connection.weight" in sd:
        feats = F.linear(feats, sd[prefix + "skip_connection.weight"], sd[prefix + "skip_connection.bias"])
    return h + feats


def _rms(x, gamma):
    return F.normalize(x.float(), dim=-1) * gamma * x.shape[-1] ** 0.5


def _attend(q, k, v):
    """q [Lq, H, d], k / v [Lk, H, d] -> [Lq, H, d]."""
    return F.scaled_dot_product_attention(q.transpose(0, 1), k.transpose(0, 1), v.transpose(0, 1)).transpose(0, 1)


def cross_block(sd, prefix, x, layout, emb, cond, num_heads, qk_rms_norm, qk_rms_norm_cross):
    """x [N, C] rows grouped by batch entry (`layout` slices), emb [B, C], cond [B, L, Cc]."""
    C = x.shape[1]
    d = C // num_heads
    mod = F.linear(F.silu(emb), sd[prefix + "adaLN_modulation.1.weight"], sd[prefix + "adaLN_modulation.1.bias"])
    out = torch.empty_like(x)
    for bi, s in enumerate(layout):
        xb = x[s]
        shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = mod[bi].chunk(6)
        h = F.layer_norm(xb, (C,), None, None, 1e-6) * (1 + scale_msa) + shift_msa
        qkv = F.linear(h, sd[prefix + "self_attn.to_qkv.weight"], sd[prefix + "self_attn.to_qkv.bias"]).reshape(-1, 3, num_heads, d)
        q, k, v = qkv.unbind(1)
        if qk_rms_norm:
            q, k = _rms(q, sd[prefix + "self_attn.q_rms_norm.gamma"]), _rms(k, sd[prefix + "self_attn.k_rms_norm.gamma"])
        h = _attend(q, k, v).reshape(-1, C)
        h = F.linear(h, sd[prefix + "self_attn.to_out.weight"], sd[prefix + "self_attn.to_out.bias"])
        xb = xb + h * gate_msa
        h = F.layer_norm(xb, (C,), sd[prefix + "norm2.weight"], sd[prefix + "norm2.bias"], 1e-6)
        q = F.linear(h, sd[prefix + "cross_attn.to_q.weight"], sd[prefix + "cross_attn.to_q.bias"]).reshape(-1, num_heads, d)
        kv = F.linear(cond[bi], sd[prefix + "cross_attn.to_kv.weight"], sd[prefix + "cross_attn.to_kv.bias"]).reshape(-1, 2, num_heads, d)
        k, v = kv.unbind(1)
        if qk_rms_norm_cross:
            q, k = _rms(q, sd[prefix + "cross_attn.q_rms_norm.gamma"]), _rms(k, sd[prefix + "cross_attn.k_rms_norm.gamma"])
        h = _attend(q, k, v).reshape(-1, C)
        xb = xb + F.linear(h, sd[prefix + "cross_attn.to_out.weight"], sd[prefix + "cross_attn.to_out.bias"])
        h = F.layer_norm(xb, (C,), None, None, 1e-6) * (1 + scale_mlp) + shift_mlp
        h = F.gelu(F.linear(h, sd[prefix + "mlp.mlp.0.weight"], sd[prefix + "mlp.mlp.0.bias"]), approximate="tanh")
        h = F.linear(h, sd[prefix + "mlp.mlp.2.weight"], sd[prefix + "mlp.mlp.2.bias"])
        out[s] = xb + h * gate_mlp
    return out


def _layout(coords, batch_size):
    counts = torch.bincount(coords[:, 0].long(), minlength=batch_size).tolist()
    off, out = 0, []
    for c in counts:
        out.append(slice(off, off + c))
        off += c
    return out


def slat_flow_forward(sd, cfg, x, coords, t, cond):
    """x [N, in_channels] at coords int [N, 4] (rows grouped by batch entry), t [B], cond [B, L, cond_channels] ->
    [N, out_channels].  cfg: the constructor arguments of the reference class."""
    sd = {k: v.float() for k, v in sd.items()}
    if cfg.get("share_mod") or cfg.get("pe_mode", "ape") != "ape":
        raise NotImplementedError("share_mod / rope are not used by the shipped structured-latent flow checkpoints")
    B = int(coords[:, 0].max()) + 1
    C = cfg["model_channels"]
    heads = cfg.get("num_heads") or C // cfg.get("num_head_channels", 64)
    io = list(cfg["io_block_channels"])
    nres = cfg.get("num_io_res_blocks", 2)
    h = F.linear(x.float(), sd["input_layer.weight"], sd["input_layer.bias"])
    emb = t_embedder(sd, t)
    skips, levels, cur = [], [], coords
    bi = 0
    for _ in io:
        for _ in range(nres - 1):
            h = res_block(sd, f"input_blocks.{bi}.", h, cur, emb, B)
            skips.append(h)
            bi += 1
        h, coarse, idx = downsample(h, cur, 2)               # SparseResBlock3d(downsample=True): _updown first (:56)
        levels.append((cur, idx))
        cur = coarse
        h = res_block(sd, f"input_blocks.{bi}.", h, cur, emb, B)
        skips.append(h)
        bi += 1
    h = h + ape(cur[:, 1:], C)
    layout = _layout(cur, B)
    for i in range(cfg["num_blocks"]):
        h = cross_block(sd, f"blocks.{i}.", h, layout, emb, cond.float(), heads, cfg.get("qk_rms_norm", False),
                        cfg.get("qk_rms_norm_cross", False))
    use_skip = cfg.get("use_skip_connection", True)
    bo = 0
    for _ in io:
        if use_skip:
            h = torch.cat([h, skips.pop()], dim=1)
        fine, idx = levels.pop()
        h, cur = h[idx], fine                                # SparseUpsample: nearest neighbour through the cached index
        h = res_block(sd, f"out_blocks.{bo}.", h, cur, emb, B)
        bo += 1
        for _ in range(nres - 1):
            if use_skip:
                h = torch.cat([h, skips.pop()], dim=1)
            h = res_block(sd, f"out_blocks.{bo}.", h, cur, emb, B)
            bo += 1
    h = F.layer_norm(h, h.shape[-1:])
    return F.linear(h, sd["out_layer.weight"], sd["out_layer.bias"])


def flow_euler_sample(model_fn, noise, steps, rescale_t=1.0, sigma_min=1e-5):
    """FlowEulerSampler.sample (trellis/pipelines/samplers/flow_euler.py:78-117) for a velocity model on feature rows."""
    import numpy as np
    t_seq = np.linspace(1, 0, steps + 1)
    t_seq = rescale_t * t_seq / (1 + (rescale_t - 1) * t_seq)
    x = noise
    for i in range(steps):
        t, tp = float(t_seq[i]), float(t_seq[i + 1])
        v = model_fn(x, 1000.0 * t)
        x = x - (t - tp) * v
    return x


def patchify(x, ps):
    """trellis/modules/spatial.py:16-31 for [N, C, D, D, D]."""
    N, C, D = x.shape[0], x.shape[1], x.shape[2] // ps
    x = x.reshape(N, C, D, ps, D, ps, D, ps).permute(0, 1, 3, 5, 7, 2, 4, 6)
    return x.reshape(N, C * ps ** 3, D, D, D)


def unpatchify(x, ps):
    """trellis/modules/spatial.py:34-48."""
    N, C, D = x.shape[0], x.shape[1] // ps ** 3, x.shape[2]
    x = x.reshape(N, C, ps, ps, ps, D, D, D).permute(0, 1, 5, 2, 6, 3, 7, 4)
    return x.reshape(N, C, D * ps, D * ps, D * ps)


def sparse_structure_flow_forward(sd, cfg, x, t, cond):
    """SparseStructureFlowModel.forward (trellis/models/sparse_structure_flow.py:174-200): the dense DiT over the occupancy
    latent x [B, C, R, R, R] -- patchify, input_layer + APE of the patch grid, ModulatedTransformerCrossBlocks
    (trellis/modules/transformer/modulated.py:132-150, the same arithmetic as the sparse blocks above on one full sequence per
    batch entry), layer_norm, out_layer, unpatchify.  Pinned by tests/golden/sparse_structure_flow_tiny.pt."""
    sd = {k: v.float() for k, v in sd.items()}
    ps, C, B = cfg["patch_size"], cfg["model_channels"], x.shape[0]
    heads = cfg.get("num_heads") or C // cfg.get("num_head_channels", 64)
    h = patchify(x.float(), ps)
    R = h.shape[2]
    h = h.reshape(B, h.shape[1], -1).permute(0, 2, 1)
    h = F.linear(h, sd["input_layer.weight"], sd["input_layer.bias"])
    grid = torch.stack(torch.meshgrid(*[torch.arange(R)] * 3, indexing="ij"), dim=-1).reshape(-1, 3)
    h = (h + ape(grid, C)[None]).reshape(B * R ** 3, C)
    emb = t_embedder(sd, t)
    layout = [slice(b * R ** 3, (b + 1) * R ** 3) for b in range(B)]
    for i in range(cfg["num_blocks"]):
        h = cross_block(sd, f"blocks.{i}.", h, layout, emb, cond.float(), heads, cfg.get("qk_rms_norm", False),
                        cfg.get("qk_rms_norm_cross", False))
    h = F.linear(F.layer_norm(h, h.shape[-1:]), sd["out_layer.weight"], sd["out_layer.bias"])
    h = h.reshape(B, R ** 3, -1).permute(0, 2, 1).reshape(B, -1, R, R, R)
    return unpatchify(h, ps).contiguous()


def pixel_shuffle_3d(x, s=2):
    """trellis/modules/spatial.py:4-13."""
    B, C, H, W, D = x.shape
    c = C // s ** 3
    return x.reshape(B, c, s, s, s, H, W, D).permute(0, 1, 5, 2, 6, 3, 7, 4).reshape(B, c, H * s, W * s, D * s)


def _channel_ln(x, w, b):
    """ChannelLayerNorm32 (trellis/modules/norm.py:19-25): LayerNorm over the channel axis of [B, C, H, W, D], eps 1e-5."""
    return F.layer_norm(x.permute(0, 2, 3, 4, 1), (x.shape[1],), w, b, 1e-5).permute(0, 4, 1, 2, 3)


def _res_block3d(sd, p, x):
    """ResBlock3d.forward (trellis/models/sparse_structure_vae.py:37-45)."""
    h = F.conv3d(F.silu(_channel_ln(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"])), sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=1)
    h = F.conv3d(F.silu(_channel_ln(h, sd[p + "norm2.weight"], sd[p + "norm2.bias"])), sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1)
    if p + "skip_connection.weight" in sd:
        x = F.conv3d(x, sd[p + "skip_connection.weight"], sd[p + "skip_connection.bias"])
    return h + x


def sparse_structure_decoder_forward(sd, cfg, z):
    """SparseStructureDecoder.forward (sparse_structure_vae.py:296-306; norm_type 'layer'): z [B, C, R, R, R] -> occupancy
    logits [B, out, R 2^(levels - 1), ...].  Pinned by tests/golden/sparse_structure_decoder_tiny.pt."""
    sd = {k: v.float() for k, v in sd.items()}
    h = F.conv3d(z.float(), sd["input_layer.weight"], sd["input_layer.bias"], padding=1)
    for i in range(cfg["num_res_blocks_middle"]):
        h = _res_block3d(sd, f"middle_block.{i}.", h)
    bi = 0
    for lvl in range(len(cfg["channels"])):
        for _ in range(cfg["num_res_blocks"]):
            h = _res_block3d(sd, f"blocks.{bi}.", h)
            bi += 1
        if lvl < len(cfg["channels"]) - 1:                       # UpsampleBlock3d, mode 'conv' (:97-101)
            h = pixel_shuffle_3d(F.conv3d(h, sd[f"blocks.{bi}.conv.weight"], sd[f"blocks.{bi}.conv.bias"], padding=1), 2)
            bi += 1
    h = F.silu(_channel_ln(h, sd["out_layer.0.weight"], sd["out_layer.0.bias"]))
    return F.conv3d(h, sd["out_layer.2.weight"], sd["out_layer.2.bias"], padding=1)
