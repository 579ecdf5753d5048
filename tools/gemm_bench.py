"""GEMM shapes of one DiT block, each scheduling variant, CUDA-event timing (L2 flushed between runs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gvfdiffusion_b200 import _lib, ops
L = _lib.lib()
dev = "cuda"
g = torch.Generator().manual_seed(0)
rn = lambda *s: torch.randn(*s, generator=g).to(dev).half()
M = 12288
shapes = [("qkv(rms)", 1536, 512, "qkv"), ("out(resid)", 512, 512, "resid"), ("q(f16)", 512, 512, "f16"),
          ("fc1(gelu)", 2048, 512, "gelu"), ("fc2(resid)", 512, 2048, "resid"), ("vae ff1", 6144, 768, "f16"),
          ("vae ff2", 768, 3072, "r16"), ("vae qkv", 2304, 768, "f16"), ("img proj", 512, 1024, "f16", 32880),
          ("img kv", 1024, 512, "f16", 32880)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, N, K, kind, *mm in shapes:
    M = mm[0] if mm else 12288
    a, w = rn(M, K), rn(N, K)
    b = torch.randn(N, generator=g).to(dev)
    x = torch.randn(M, N, generator=g).to(dev)
    x16 = x.half()
    o16 = torch.empty(M, N, dtype=torch.float16, device=dev)
    gq = torch.ones(N // 96, 32, device=dev) if kind == "qkv" else None
    res = []
    for variant in (0, 1, 2, 3, 4, 5, 6, 7):
        L.gvf_gemm_set_variant(variant)
        def run():
            if kind == "qkv": ops.gemm_qkv_rmsnorm(a, w, b, gq, gq, o16)
            elif kind == "resid": ops.gemm(a, w, b, ops.EPI_RESID_F32, out=x)
            elif kind == "r16": ops.gemm(a, w, b, ops.EPI_RESID_F16, out=x16)
            elif kind == "gelu": ops.gemm(a, w, b, ops.EPI_GELU_F16, out=o16)
            else: ops.gemm(a, w, b, ops.EPI_F16, out=o16)
        for _ in range(3): run()
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        # warm (no flush, back to back)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): run()
        e1.record(); torch.cuda.synchronize()
        res.append((ts[len(ts) // 2] * 1e3, e0.elapsed_time(e1) / 20 * 1e3))
    # cuBLAS reference point (library GEMM + bias only, no fused epilogue): what a stock fp16 nn.Linear costs
    bh = b.half()
    for _ in range(3): torch.nn.functional.linear(a, w, bh)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): torch.nn.functional.linear(a, w, bh)
    e1.record(); torch.cuda.synchronize()
    cublas_us = e0.elapsed_time(e1) / 20 * 1e3
    fl = 2 * M * N * K
    print(f"{name:12s} N={N:5d} K={K:5d} " + "  ".join(f"v{v}: cold {c:6.1f}us warm {wm:6.1f}us ({fl / wm / 1e6:6.0f} TF/s)" for v, (c, wm) in enumerate(res)) + f"  cuBLAS linear warm {cublas_us:6.1f}us ({fl / cublas_us / 1e6:6.0f} TF/s)")
L.gvf_gemm_set_variant(-1)
