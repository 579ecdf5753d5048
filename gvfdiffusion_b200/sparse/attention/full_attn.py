"""Full (per batch entry) attention over sparse voxels -- reference sparse/attention/full_attn.py:90-215
(`attn_mode: full` of SparseMultiHeadAttention, modules.py:189-192; the TRELLIS structured-latent flow blocks,
trellis/modules/sparse/attention/).  The reference turns the batch layout into cu_seqlens and calls
flash_attn_varlen_{qkvpacked,kvpacked,}_func on the packed rows:

  * self-attention, `sparse_scaled_dot_product_attention(qkv)` with qkv.feats [T, 3, H, 64]: one launch of
    gvf_sparse_varlen_attn_f16 over the rows in place (no gather list: a batch entry's voxels are contiguous);
  * cross-attention of voxels against a DENSE context, `(q, kv)` / `(q, k, v)` with q sparse [T, H, C] and the context
    [N, L, 2, H, C] / 2 x [N, L, H, C]: per batch entry one call of the dense tcgen05 kernel (gvf_attn_fwd_f16) on that
    entry's row range -- head dims 32 / 64.
Other argument mixes of the reference's overload set (dense q against sparse kv, sparse against sparse cross) are not
used by the GVF / TRELLIS models and raise.
"""
import math

import torch

from ... import _lib, ops
from ..basic import SparseTensor
from ..._lib import check, current_stream, ptr


def _self_attention(qkv):
    f = qkv.feats
    if not (f.is_cuda and f.dtype == torch.float16 and f.dim() == 4 and f.shape[1] == 3):
        raise ValueError(f"qkv.feats: expected a CUDA fp16 [T, 3, H, C] tensor, got {tuple(f.shape)} {f.dtype}")
    T, _, H, C = f.shape
    lens = [s.stop - s.start for s in qkv.layout]
    cu = torch.zeros(len(lens) + 1, dtype=torch.int32)
    cu[1:] = torch.cumsum(torch.tensor(lens, dtype=torch.int32), 0)
    cu = cu.to(f.device)
    q = f.contiguous()
    out = torch.empty((T, H, C), dtype=torch.float16, device=f.device)
    check(_lib.lib().gvf_sparse_varlen_attn_f16(ptr(q), ptr(out), None, None, ptr(cu), len(lens), max(lens) if lens else 0,
                                                H, C, 1.0 / math.sqrt(C), current_stream()), "gvf_sparse_varlen_attn_f16")
    return qkv.replace(out)


def _cross_dense_context(q, k, v):
    """q SparseTensor [T, H, C]; k, v dense [N, L, H, C] views."""
    f = q.feats
    if not (f.is_cuda and f.dtype == torch.float16 and f.dim() == 3):
        raise ValueError(f"q.feats: expected a CUDA fp16 [T, H, C] tensor, got {tuple(f.shape)} {f.dtype}")
    T, H, C = f.shape
    out = torch.empty((T, H, C), dtype=torch.float16, device=f.device)
    fc = f.contiguous()
    for b, s in enumerate(q.layout):
        if s.stop > s.start:
            ops.attention(fc[s.start:s.stop][None], k[b:b + 1], v[b:b + 1], 1.0 / math.sqrt(C), out=out[s.start:s.stop][None])
    return q.replace(out)


def sparse_scaled_dot_product_attention(*args, **kwargs):
    names = {1: ["qkv"], 2: ["q", "kv"], 3: ["q", "k", "v"]}
    n = len(args) + len(kwargs)
    if n not in names:
        raise ValueError(f"Invalid number of arguments, got {n}, expected 1, 2, or 3")
    vals = list(args) + [kwargs[k] for k in names[n][len(args):]]
    if n == 1:
        (qkv,) = vals
        if not isinstance(qkv, SparseTensor):
            raise TypeError(f"qkv must be a SparseTensor, got {type(qkv)}")
        return _self_attention(qkv)
    if n == 2:
        q, kv = vals
        if isinstance(q, SparseTensor) and isinstance(kv, torch.Tensor):
            if kv.dim() != 5 or kv.shape[2] != 2:
                raise ValueError(f"Invalid shape for kv, got {tuple(kv.shape)}, expected [N, L, 2, H, C]")
            return _cross_dense_context(q, kv[:, :, 0], kv[:, :, 1])
    else:
        q, k, v = vals
        if isinstance(q, SparseTensor) and isinstance(k, torch.Tensor) and isinstance(v, torch.Tensor):
            return _cross_dense_context(q, k, v)
    raise NotImplementedError("only sparse self-attention and sparse-query x dense-context cross-attention are on the "
                              "GVF / TRELLIS paths")
