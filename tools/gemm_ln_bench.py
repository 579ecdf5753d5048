"""Fused residual GEMM + LayerNorm vs the two-kernel path at the DiT shapes (warm, back to back)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gvfdiffusion_b200 import ops
g = torch.Generator().manual_seed(0)
rn = lambda *s: torch.randn(*s, generator=g).cuda().half()
M, N = 12288, 512
for K in (512, 2048):
    a, w = rn(M, K), rn(N, K) * 0.05
    b = torch.randn(N, generator=g).cuda()
    x = torch.randn(M, N, generator=g).cuda()
    gate, mod = rn(1, N), rn(1, 2 * N)
    y = torch.empty(M, N, dtype=torch.float16, device="cuda")
    def fused(): ops.gemm_resid_ln(a, w, b, x, y, gate=gate, gate_stride=N, rows_per_batch=M, shift=mod[:, :N], scale=mod[:, N:], mod_stride=2 * N)
    def split():
        ops.gemm(a, w, b, ops.EPI_RESID_F32, out=x, gate=gate, gate_stride=N, rows_per_batch=M)
        ops.ln_mod(x, out=y, shift=mod[:, :N], scale=mod[:, N:], mod_stride=2 * N, rows_per_batch=M)
    for name, fn in (("two kernels", split), ("fused", fused)):
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        print(f"K={K:5d} {name:12s} {e0.elapsed_time(e1) / 20 * 1e3:7.1f} us")
