"""ncu target: the attention backward kernels at the motion-VAE training shapes (self: 24 x 512 x 512, 12 heads of 64;
decoder cross: 16384 shared queries x 24 frames x 512 keys).

    ncu --set full --clock-control none --import-source on -k regex:attn_bwd -c 4 python tools/profile_attn_bwd.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gvfdiffusion_b200 import ops  # noqa: E402

dev = "cuda"
g = torch.Generator().manual_seed(0)
r = lambda *s: (torch.randn(*s, generator=g) * 0.7).half().to(dev)
# self-attention
qkv = r(24, 512, 3, 12, 64)
do = r(24, 512, 12, 64)
o, lse = ops.attention_fwd_lse(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], 0.125)
d = torch.empty_like(qkv)
ops.attention_bwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], o, do, lse, 0.125, d[:, :, 0], d[:, :, 1], d[:, :, 2])
# decoder cross-attention
q = r(16384, 12, 64)
kv = r(24, 512, 2, 12, 64)
do2 = r(24, 16384, 12, 64)
o2, lse2 = ops.attention_fwd_lse(q, kv[:, :, 0], kv[:, :, 1], 0.125, q_shared=True)
dq, dkv = torch.empty_like(q), torch.empty_like(kv)
ops.attention_bwd(q, kv[:, :, 0], kv[:, :, 1], o2, do2, lse2, 0.125, dq, dkv[:, :, 0], dkv[:, :, 1], q_shared=True)
torch.cuda.synchronize()
if "--time" in sys.argv:
    for name, fn in (("self", lambda: ops.attention_bwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], o, do, lse, 0.125, d[:, :, 0],
                                                        d[:, :, 1], d[:, :, 2])),
                     ("decoder", lambda: ops.attention_bwd(q, kv[:, :, 0], kv[:, :, 1], o2, do2, lse2, 0.125, dq, dkv[:, :, 0],
                                                           dkv[:, :, 1], q_shared=True))):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(name, "bwd (prep + dkdv + dq)", e0.elapsed_time(e1) / 10, "ms")
