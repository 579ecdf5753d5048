"""`scaled_dot_product_attention`: the reference's dense-attention dispatch point
(model/attention/full_attn.py:74-140) backed by the sm_100a kernels.  Accepts the same three
call forms -- (qkv [N,L,3,H,C]) | (q [N,L,H,C], kv [N,L,2,H,C]) | (q, k, v [N,L,H,C]) -- and returns
[N,L,H,C].  fp16 tensors are consumed in place (views, no unbind copies); other dtypes are cast."""
import math

import torch

from .. import ops


def scaled_dot_product_attention(*args, **kwargs):
    names = {1: ["qkv"], 2: ["q", "kv"], 3: ["q", "k", "v"]}
    n = len(args) + len(kwargs)
    assert n in names, f"Invalid number of arguments, got {n}, expected 1, 2, or 3"
    vals = list(args) + [kwargs[k] for k in names[n][len(args):]]
    if n == 1:
        qkv = vals[0]
        assert qkv.dim() == 5 and qkv.shape[2] == 3, f"Invalid shape for qkv, got {qkv.shape}, expected [N, L, 3, H, C]"
        q, k, v = qkv.unbind(2)
    elif n == 2:
        q, kv = vals
        assert q.shape[0] == kv.shape[0] and q.dim() == 4 and kv.dim() == 5
        k, v = kv.unbind(2)
    else:
        q, k, v = vals
        assert q.shape[0] == k.shape[0] == v.shape[0] and q.dim() == k.dim() == v.dim() == 4
    dt = q.dtype
    h = lambda t: t if t.dtype == torch.float16 else t.to(torch.float16)
    out = ops.attention(h(q), h(k), h(v), 1.0 / math.sqrt(q.shape[-1]))
    return out if dt == torch.float16 else out.to(dt)
