"""The two fp32-residual GEMM shapes of a DiT block (out-proj 12288x512x512, fc2 12288x512x2048) under the
default heuristic, three launches each, for `ncu --set full -k regex:gemm_ws`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gvfdiffusion_b200 import ops
g = torch.Generator().manual_seed(0)
rn = lambda *s: torch.randn(*s, generator=g).cuda().half()
M = 12288
x = torch.randn(M, 512, generator=g).cuda()
gate = rn(1, 512)
b = torch.randn(512, generator=g).cuda()
for K in (512, 2048):
    a, w = rn(M, K), rn(512, K)
    for _ in range(3):
        ops.gemm(a, w, b, ops.EPI_RESID_F32, out=x, gate=gate, gate_stride=512, rows_per_batch=M)
torch.cuda.synchronize()
print("done")
