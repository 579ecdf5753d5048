"""Serialized (space-filling-curve) windowed self-attention over sparse voxels -- reference
sparse/attention/serialized_attn.py:38-192 (`attn_mode: serialized` of SparseMultiHeadAttention, modules.py:197-201).

Voxels of a batch entry are ordered along a z-order / Hilbert curve (vox2seq codes, csrc/vox2seq.cu) and cut into
ceil(n / window_size) windows of EXACTLY window_size tokens: each window is centred on its share of the sequence and
padded on both sides with its neighbours (wrapping around the ends), and only the results of the un-padded middle are
kept (`out[bwd_indices]`).  The reference gathers `qkv.feats[fwd_indices]`, runs flash-attn and gathers again; here the
sequences are read through `fwd_indices` and every kept row is written straight to its voxel by
gvf_sparse_varlen_attn_f16 (scatter list = voxel row for the valid positions, -1 for the padding), so neither copy
exists.  The partition itself is index plumbing on the device with the reference's arithmetic (window bounds are
computed on the host in Python floats exactly like :84-101).
"""
import math
from enum import Enum

import torch

from .. import _partition_cache
from ... import _lib
from ... import vox2seq
from ..._lib import check, current_stream, ptr


class SerializeMode(Enum):
    Z_ORDER = 0
    Z_ORDER_TRANSPOSED = 1
    HILBERT = 2
    HILBERT_TRANSPOSED = 3


SerializeModes = [SerializeMode.Z_ORDER, SerializeMode.Z_ORDER_TRANSPOSED, SerializeMode.HILBERT,
                  SerializeMode.HILBERT_TRANSPOSED]

_CURVE = {SerializeMode.Z_ORDER: ("z_order", (0, 1, 2)), SerializeMode.Z_ORDER_TRANSPOSED: ("z_order", (1, 0, 2)),
          SerializeMode.HILBERT: ("hilbert", (0, 1, 2)), SerializeMode.HILBERT_TRANSPOSED: ("hilbert", (1, 0, 2))}


def calc_serialization(tensor, window_size, serialize_mode=SerializeMode.Z_ORDER, shift_sequence=0,
                       shift_window=(0, 0, 0), encode=None):
    """tensor: anything with `.coords` [T, 4] int32 (batch, x, y, z), `.layout` (one slice of rows per batch entry) and
    `.device`.  -> fwd_indices [M] int64, bwd_indices [T] int64, seq_lens (list), seq_batch_indices (list), like the
    reference.  `encode` overrides the curve encoder (tests run the partition on CPU tensors with the oracle's)."""
    if serialize_mode not in _CURVE:
        raise ValueError(f"Unknown serialize mode: {serialize_mode}")
    mode, perm = _CURVE[serialize_mode]
    dev = tensor.coords.device
    sc = tensor.coords[:, 1:].clone()
    sc += torch.tensor(shift_window, dtype=sc.dtype, device=dev).reshape(1, 3)
    code = (encode or vox2seq.encode)(sc, permute=list(perm), mode=mode)
    fwd, bwd, seq_lens, seq_batch, base = [], [], [], [], 0
    for bi, s in enumerate(tensor.layout):
        n = s.stop - s.start
        nw = (n + window_size - 1) // window_size
        order = torch.argsort(code[s.start:s.stop])
        if nw == 1:
            inv = torch.empty_like(order)
            inv[order] = torch.arange(n, device=dev)
            fwd.append(order + s.start)
            bwd.append(inv + base)
            seq_lens.append(n)
            seq_batch.append(bi)
            base += n
            continue
        share = n / nw                                           # tokens a window is responsible for (float)
        inv = torch.zeros(n, dtype=torch.int64, device=dev)
        pos = 0                                                  # running position inside this entry's padded list
        for i in range(nw):
            lo, hi = math.floor(i * share + shift_sequence), math.floor((i + 1) * share + shift_sequence)
            p0 = math.floor((i + 0.5) * share + shift_sequence - 0.5 * window_size)
            win = order[torch.arange(p0, p0 + window_size, device=dev) % n]
            pos += lo - p0
            inv.scatter_(0, win[lo - p0:hi - p0], torch.arange(pos, pos + hi - lo, device=dev))
            pos += p0 + window_size - lo
            fwd.append(win + s.start)
        seq_lens += [window_size] * nw
        seq_batch += [bi] * nw
        bwd.append(inv + base)
        base += nw * window_size
    return torch.cat(fwd), torch.cat(bwd), seq_lens, seq_batch


def _serialization(qkv, window_size, serialize_mode, shift_sequence, shift_window):
    coords = qkv.coords
    key = ("ser", coords.data_ptr(), coords._version, tuple(coords.shape), window_size, serialize_mode, shift_sequence,
           tuple(shift_window))
    hit = _partition_cache.get(key)
    if hit is None:
        fwd, bwd, seq_lens, _ = calc_serialization(qkv, window_size, serialize_mode, shift_sequence, shift_window)
        T, M = coords.shape[0], fwd.shape[0]
        scatter = torch.full((M,), -1, dtype=torch.int32, device=coords.device)
        scatter[bwd] = torch.arange(T, dtype=torch.int32, device=coords.device)
        cu = torch.zeros(len(seq_lens) + 1, dtype=torch.int32)
        cu[1:] = torch.cumsum(torch.tensor(seq_lens, dtype=torch.int32), 0)
        hit = (fwd.int().contiguous(), scatter, cu.to(coords.device), max(seq_lens) if seq_lens else 0, coords)
        if len(_partition_cache) > 64:
            _partition_cache.clear()
        _partition_cache[key] = hit
    return hit[:4]


def sparse_serialized_scaled_dot_product_self_attention(qkv, window_size, serialize_mode=SerializeMode.Z_ORDER,
                                                        shift_sequence=0, shift_window=(0, 0, 0)):
    """qkv: SparseTensor with feats [T, 3, H, 64] fp16 on CUDA -> SparseTensor with feats [T, H, 64] (original order)."""
    f = qkv.feats
    if not (f.is_cuda and f.dtype == torch.float16 and f.dim() == 4 and f.shape[1] == 3):
        raise ValueError(f"qkv.feats: expected a CUDA fp16 [T, 3, H, C] tensor, got {tuple(f.shape)} {f.dtype}")
    T, _, H, C = f.shape
    fwd, scatter, cu, max_len = _serialization(qkv, window_size, serialize_mode, shift_sequence, tuple(shift_window))
    q = f.contiguous()
    out = torch.empty((T, H, C), dtype=torch.float16, device=q.device)
    check(_lib.lib().gvf_sparse_varlen_attn_f16(ptr(q), ptr(out), ptr(fwd), ptr(scatter), ptr(cu), cu.shape[0] - 1, max_len,
                                                H, C, 1.0 / math.sqrt(C), current_stream()), "gvf_sparse_varlen_attn_f16")
    return qkv.replace(out)
