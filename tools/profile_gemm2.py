"""A few launches of the generation-2 GEMM at the DiT shapes for `ncu --set full`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gvfdiffusion_b200 import ops
g = torch.Generator().manual_seed(0)
rn = lambda *s: torch.randn(*s, generator=g).cuda().half()
M = 12288
a, wo, w2, h1 = rn(M, 512), rn(512, 512), rn(512, 2048), rn(M, 2048)
b = torch.randn(512, generator=g).cuda()
x = torch.randn(M, 512, generator=g).cuda()
gate = rn(1, 512)
q = torch.empty(M, 512, dtype=torch.float16, device="cuda")
for _ in range(3):
    ops.gemm(a, wo, b, ops.EPI_RESID_F32, out=x, gate=gate, gate_stride=512, rows_per_batch=M)   # out(resid)
    ops.gemm(a, wo, b, ops.EPI_F16, out=q)                                                       # q
    ops.gemm(h1, w2, b, ops.EPI_RESID_F32, out=x, gate=gate, gate_stride=512, rows_per_batch=M)  # fc2
torch.cuda.synchronize()
print("done")
