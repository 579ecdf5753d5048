"""fc2 (12288 x 512 x 2048, gated fp32 residual) under each generation-2 variant, for `ncu --set full`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gvfdiffusion_b200 import _lib, ops
L = _lib.lib()
g = torch.Generator().manual_seed(0)
rn = lambda *s: torch.randn(*s, generator=g).cuda().half()
M = 12288
h1, w2 = rn(M, 2048), rn(512, 2048)
b = torch.randn(512, generator=g).cuda()
x = torch.randn(M, 512, generator=g).cuda()
gate = rn(1, 512)
for v in (4, 5, 7):
    L.gvf_gemm_set_variant(v)
    for _ in range(3):
        ops.gemm(h1, w2, b, ops.EPI_RESID_F32, out=x, gate=gate, gate_stride=512, rows_per_batch=M)
L.gvf_gemm_set_variant(-1)
torch.cuda.synchronize()
print("done")
