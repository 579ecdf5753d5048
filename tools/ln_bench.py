"""LayerNorm + modulate pass at the DiT's shape (12288 x 512 fp32 -> fp16, 37.7 MB): one row per warp vs two rows in flight."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gvfdiffusion_b200 import _lib, ops
L = _lib.lib()
g = torch.Generator().manual_seed(0)
for M, C in ((12288, 512), (12288, 768), (3656, 1024)):
    x = torch.randn(M, C, generator=g).cuda()
    mod = (torch.randn(1, 2 * C, generator=g) * 0.3).half().cuda()
    out = torch.empty(M, C, dtype=torch.float16, device="cuda")
    for two, name in ((0, "one row / warp"), (1, "two rows in flight")):
        L.gvf_ln_set_two_rows(two)
        gr = torch.cuda.CUDAGraph()
        for _ in range(3):
            ops.ln_mod(x, out=out, shift=mod[:, :C], scale=mod[:, C:], mod_stride=2 * C, rows_per_batch=M)
        torch.cuda.synchronize()
        with torch.cuda.graph(gr):                     # 50 back-to-back launches: device time without host launch gaps
            for _ in range(50):
                ops.ln_mod(x, out=out, shift=mod[:, :C], scale=mod[:, C:], mod_stride=2 * C, rows_per_batch=M)
        gr.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 50
        print(f"M={M} C={C} {name:20s} {us:6.2f} us  ({M * C * 6 / us / 1e6:5.2f} TB/s)")
L.gvf_ln_set_two_rows(0)
