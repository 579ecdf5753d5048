"""Pipeline-depth / single-round probe of the residual GEMMs: time per variant at the DiT shapes and at
M chosen so that every CTA (pair) owns exactly one tile (no wave quantisation)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gvfdiffusion_b200 import _lib, ops
L = _lib.lib()
dev = "cuda"
g = torch.Generator().manual_seed(0)
rn = lambda *s: torch.randn(*s, generator=g).to(dev).half()


def bench(M, N, K, variant, kind="resid"):
    a, w = rn(M, K), rn(N, K)
    b = torch.randn(N, generator=g).to(dev)
    x = torch.randn(M, N, generator=g).to(dev)
    o16 = torch.empty(M, N, dtype=torch.float16, device=dev)
    L.gvf_gemm_set_variant(variant)
    run = (lambda: ops.gemm(a, w, b, ops.EPI_RESID_F32, out=x)) if kind == "resid" else (lambda: ops.gemm(a, w, b, ops.EPI_F16, out=o16))
    for _ in range(3): run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize()
    L.gvf_gemm_set_variant(-1)
    return e0.elapsed_time(e1) / 20 * 1e3


for name, M, N, K in [("fc2 full", 12288, 512, 2048), ("out full", 12288, 512, 512),
                      ("fc2 one round 128x128", 148 * 128, 128, 2048), ("fc2 one round 128x256", 148 * 128, 256, 2048),
                      ("fc2 one round pair 256x256", 74 * 256, 256, 2048), ("fc2 one round pair 256x128", 74 * 256, 128, 2048),
                      ("K=512 one round 128x128", 148 * 128, 128, 512), ("K=8192 one round 128x128", 148 * 128, 128, 8192)]:
    for kind in ("resid", "f16"):
        out = []
        for v in (4, 8, 5, 6, 9, 7):
            try:
                out.append(f"v{v} {bench(M, N, K, v, kind):6.1f}")
            except Exception as e:
                out.append(f"v{v} ERR")
        print(f"{name:28s} {kind:5s} M={M:6d} N={N:4d} K={K:5d}  " + "  ".join(out), flush=True)
