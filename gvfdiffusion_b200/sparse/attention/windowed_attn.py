"""Windowed self-attention over sparse voxels (reference sparse/attention/windowed_attn.py).

Same partition as `calc_window_partition` (:20-58): window id = floor((coord + shift) / window) linearised
with the batch index as the slowest digit, voxels ordered by window id.  The reference then gathers
`qkv.feats[fwd_indices]`, calls flash-attn varlen and scatters back with `bwd_indices` (:92-129); here the
kernel reads and writes through `fwd_indices` itself (gvf_sparse_window_attn_f16), so only the index arrays
are materialised.  Differences that do not change results: the sort is stable (the reference's argsort is
not, so its order inside a window is unspecified -- attention does not depend on it); sequence lengths stay
on the device (the reference pulls them to the host with .tolist(); we read one scalar, the longest window,
and cache the partition per coords tensor like the reference's spatial cache).
"""
import math

import torch

from .. import _partition_cache
from ... import _lib
from ..._lib import check, current_stream, ptr


def calc_window_partition(coords, window_size, shift_window=0):
    """coords [T, 1 + DIM] int32 (batch, x, y, z) on the device.
    -> fwd_indices [T] int64, bwd_indices [T] int64, seq_lens [W] int32, seq_batch_indices [W] int32
    (empty windows removed, as the reference does at :52-55)."""
    DIM = coords.shape[1] - 1
    shift = (shift_window,) * DIM if isinstance(shift_window, int) else tuple(shift_window)
    win = (window_size,) * DIM if isinstance(window_size, int) else tuple(window_size)
    dev = coords.device
    sc = coords.clone().long()
    sc[:, 1:] += torch.tensor(shift, device=dev, dtype=torch.long)[None]
    max_coords = sc[:, 1:].max(dim=0).values.tolist()
    num_windows = [math.ceil((mc + 1) / ws) for mc, ws in zip(max_coords, win)]
    offset = [1]
    for n in num_windows[::-1]:
        offset.append(offset[-1] * n)
    offset = offset[::-1]                                   # [prod(all), ..., nz, 1]
    sc[:, 1:] //= torch.tensor(win, device=dev, dtype=torch.long)[None]
    ids = (sc * torch.tensor(offset, device=dev, dtype=torch.long)[None]).sum(dim=1)
    fwd = torch.argsort(ids, stable=True)
    bwd = torch.empty_like(fwd)
    bwd[fwd] = torch.arange(fwd.shape[0], device=dev)
    counts = torch.bincount(ids)
    batch = torch.arange(counts.shape[0], device=dev, dtype=torch.int32) // offset[0]
    mask = counts != 0
    return fwd, bwd, counts[mask].int(), batch[mask]


def _partition(coords, window_size, shift_window):
    key = (coords.data_ptr(), coords._version, tuple(coords.shape), window_size, tuple(shift_window))
    hit = _partition_cache.get(key)
    if hit is None:
        fwd, bwd, seq_lens, seq_batch = calc_window_partition(coords, window_size, shift_window)
        cu = torch.zeros(seq_lens.shape[0] + 1, dtype=torch.int32, device=coords.device)
        cu[1:] = torch.cumsum(seq_lens, 0)
        # the entry holds `coords` itself: the key contains its address, which must not be recycled by another
        # coordinate tensor while the entry lives (the reference hangs the partition on the SparseTensor's
        # spatial cache, :79-85, which has the same lifetime)
        # window of every sorted position (the packed tiling of csrc/sparse_attn.cu: 64 consecutive positions per CTA)
        seq_of_pos = torch.repeat_interleave(torch.arange(seq_lens.shape[0], device=coords.device, dtype=torch.int32),
                                             seq_lens.long(), output_size=int(fwd.shape[0])).contiguous()
        hit = (fwd.int().contiguous(), bwd, cu, int(seq_lens.max()) if seq_lens.numel() else 0, coords, seq_of_pos)
        if len(_partition_cache) > 64:
            _partition_cache.clear()
        _partition_cache[key] = hit
    return hit[:4]


def _seq_of_pos(coords, window_size, shift_window):
    """seq_of_pos [T] int32 of the (cached) partition."""
    _partition(coords, window_size, shift_window)
    return _partition_cache[(coords.data_ptr(), coords._version, tuple(coords.shape), window_size, tuple(shift_window))][5]


def sparse_windowed_scaled_dot_product_self_attention(qkv_feats, coords, window_size, shift_window=(0, 0, 0), packed=None):
    """qkv_feats [T, 3, H, C] fp16 CUDA (SparseTensor.feats of the fused to_qkv output), coords [T, 4] int32.
    -> [T, H, C] fp16 in the ORIGINAL voxel order (what `qkv.replace(out)` holds in the reference).
    packed: None = choose by the longest window; True / False force the packed / per-window tiling (same results)."""
    if not (qkv_feats.is_cuda and qkv_feats.dtype == torch.float16 and qkv_feats.dim() == 4 and qkv_feats.shape[1] == 3):
        raise ValueError(f"qkv_feats: expected a CUDA fp16 [T, 3, H, C] tensor, got {tuple(qkv_feats.shape)} {qkv_feats.dtype}")
    T, _, H, C = qkv_feats.shape
    shift = tuple(shift_window) if not isinstance(shift_window, int) else (shift_window,) * (coords.shape[1] - 1)
    fwd, _bwd, cu, max_len = _partition(coords, window_size, shift)
    q = qkv_feats.contiguous()
    out = torch.empty((T, H, C), dtype=torch.float16, device=q.device)
    if packed is None:
        packed = max_len < 256            # short windows: several per 64-row tile; long ones: one CTA per window tile
    if packed:
        sop = _seq_of_pos(coords, window_size, shift)
        st = _lib.lib().gvf_sparse_packed_attn_f16(ptr(q), ptr(out), None, ptr(fwd), None, ptr(cu), ptr(sop), T, H, C,
                                                   1.0 / math.sqrt(C), current_stream())
        check(st, "gvf_sparse_packed_attn_f16")
        return out
    st = _lib.lib().gvf_sparse_window_attn_f16(ptr(q), ptr(out), ptr(fwd), ptr(cu), cu.shape[0] - 1, max_len, H, C,
                                               1.0 / math.sqrt(C), current_stream())
    check(st, "gvf_sparse_window_attn_f16")
    return out


def windowed_attention_fwd_lse(qkv_feats, coords, window_size, shift_window, packed=None):
    """Forward that also leaves LSE2 [T, H] (log2-domain log-sum-exp of the scaled scores) for the backward.
    -> (out [T, H, C] fp16, lse2, partition = (fwd_indices, cu_seqlens, max_len, seq_of_pos or None))."""
    T, _, H, C = qkv_feats.shape
    fwd, _bwd, cu, max_len = _partition(coords, window_size, shift_window)
    q = qkv_feats.contiguous()
    out = torch.empty((T, H, C), dtype=torch.float16, device=q.device)
    lse = torch.empty((T, H), dtype=torch.float32, device=q.device)
    if packed is None:
        packed = max_len < 256
    if packed:
        sop = _seq_of_pos(coords, window_size, shift_window)
        check(_lib.lib().gvf_sparse_packed_attn_f16(ptr(q), ptr(out), ptr(lse), ptr(fwd), None, ptr(cu), ptr(sop), T, H, C,
                                                    1.0 / math.sqrt(C), current_stream()), "gvf_sparse_packed_attn_f16")
        return out, lse, (fwd, cu, max_len, sop)
    check(_lib.lib().gvf_sparse_varlen_attn_lse_f16(ptr(q), ptr(out), ptr(lse), ptr(fwd), None, ptr(cu), cu.shape[0] - 1,
                                                    max_len, H, C, 1.0 / math.sqrt(C), current_stream()),
          "gvf_sparse_varlen_attn_lse_f16")
    return out, lse, (fwd, cu, max_len, None)


def windowed_attention_bwd(qkv_feats, out, dout, lse, partition, dqkv=None):
    """dqkv [T, 3, H, C] fp16 of the windowed attention (csrc/sparse_attn_bwd.cu, window gather fused)."""
    fwd, cu, max_len, sop = partition
    T, _, H, C = qkv_feats.shape
    dout = dout.to(torch.float16).contiguous()
    if dqkv is None:
        dqkv = torch.zeros_like(qkv_feats)               # rows outside every window (none for a partition) stay zero
    dsum = torch.empty_like(lse)
    if sop is not None:
        check(_lib.lib().gvf_sparse_packed_attn_bwd_f16(ptr(qkv_feats), ptr(out), ptr(dout), ptr(lse), ptr(dsum), ptr(dqkv), ptr(fwd),
                                                        ptr(cu), ptr(sop), T, T, H, C, 1.0 / math.sqrt(C), current_stream()),
              "gvf_sparse_packed_attn_bwd_f16")
        return dqkv
    check(_lib.lib().gvf_sparse_varlen_attn_bwd_f16(ptr(qkv_feats), ptr(out), ptr(dout), ptr(lse), ptr(dsum), ptr(dqkv), ptr(fwd),
                                                    ptr(cu), cu.shape[0] - 1, max_len, T, H, C, 1.0 / math.sqrt(C),
                                                    current_stream()), "gvf_sparse_varlen_attn_bwd_f16")
    return dqkv


class _WindowedAttnFn(torch.autograd.Function):
    """Windowed attention under autograd (training step of the static VAE): forward that also leaves
    LSE2, backward on csrc/sparse_attn_bwd.cu."""

    @staticmethod
    def forward(ctx, qkv_feats, coords, window_size, shift_window, packed):
        q = qkv_feats.detach().contiguous()
        out, lse, (fwd, cu, max_len, sop) = windowed_attention_fwd_lse(q, coords, window_size, shift_window, packed)
        ctx.save_for_backward(q, out, lse, fwd, cu)
        ctx.max_len, ctx.sop = max_len, sop
        return out

    @staticmethod
    def backward(ctx, dout):
        q, out, lse, fwd, cu = ctx.saved_tensors
        return windowed_attention_bwd(q, out, dout, lse, (fwd, cu, ctx.max_len, ctx.sop)), None, None, None, None


def sparse_windowed_attention_autograd(qkv_feats, coords, window_size, shift_window=(0, 0, 0), packed=None):
    """Differentiable form of `sparse_windowed_scaled_dot_product_self_attention` (gradient with respect to qkv_feats)."""
    shift = tuple(shift_window) if not isinstance(shift_window, int) else (shift_window,) * (coords.shape[1] - 1)
    return _WindowedAttnFn.apply(qkv_feats, coords, window_size, shift, packed)
