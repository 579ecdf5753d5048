import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gvfdiffusion_b200 import ops
g = torch.Generator().manual_seed(0)
x = torch.randn(12288, 16, generator=g).cuda(); W = (torch.randn(512, 16, generator=g) * 0.2).half().cuda()
b = torch.randn(512, generator=g).cuda(); pos = torch.randn(512, 512, generator=g).cuda()
out = torch.empty(12288, 512, device="cuda")
X = torch.randn(12288, 512, generator=g).cuda(); mod = (torch.randn(1, 1024, generator=g) * 0.3).half().cuda()
Wf = (torch.randn(16, 512, generator=g) * 0.05).half().cuda(); bf = torch.randn(16, generator=g).cuda(); vo = torch.empty(12288, 16, device="cuda")
for name, fn in (("small_linear 12288x16->512 + add", lambda: ops.small_linear(x, W, b, out_f16=False, add=pos, add_rows=512, out=out)),
                 ("final_layer 12288x512->16", lambda: ops.dit_final_layer(X, mod[:, :512], mod[:, 512:], 1024, 12288, Wf, bf, out=vo))):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(50): fn()
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) * 1e3 / 50:.2f} us")
