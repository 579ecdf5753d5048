"""Kernel-variant A/B on the three long DiT attention shapes (static / image cross, spatial self):
    python tools/attn_variants.py 0 0x10000 0x20000 ...       (gvf_attn_set_debug values)
Median of 30 launches (CUDA events), max error against an fp32 torch reference of two batch entries."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gvfdiffusion_b200 import _lib, ops
L = _lib.lib()
dev = "cuda"
g = torch.Generator().manual_seed(0)
rn = lambda *s: torch.randn(*s, generator=g).to(dev).half()
T, N, H, D = 24, 512, 16, 32
sc = 1 / math.sqrt(D)
q = rn(T, N, H, D)
kv_s, kv_i, qkv = rn(4096, 2, H, D), rn(T, 1370, 2, H, D), rn(T, N, 3, H, D)
o = torch.empty(T, N, H, D, dtype=torch.float16, device=dev)
shapes = {
    "static": (lambda: ops.attention(q, kv_s[:, 0], kv_s[:, 1], sc, out=o, kv_shared=True), 4096,
               lambda: (q[:2], kv_s[None, :, 0].expand(2, -1, -1, -1), kv_s[None, :, 1].expand(2, -1, -1, -1))),
    "image": (lambda: ops.attention(q, kv_i[:, :, 0], kv_i[:, :, 1], sc, out=o), 1370,
              lambda: (q[:2], kv_i[:2, :, 0], kv_i[:2, :, 1])),
    "spatial": (lambda: ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], sc, out=o), 512,
                lambda: (qkv[:2, :, 0], qkv[:2, :, 1], qkv[:2, :, 2])),
}
refs = {}
for name, (_, _, mk) in shapes.items():
    qq, kk, vv = (t.float() for t in mk())
    s_ = torch.einsum("tnhd,tkhd->thnk", qq, kk) * sc
    refs[name] = torch.einsum("thnk,tkhd->tnhd", s_.softmax(-1), vv)
vals = [int(v, 0) for v in sys.argv[1:]] or [0]
for dbg in vals:
    L.gvf_attn_set_debug(dbg)
    row = [f"dbg {dbg:#9x}"]
    for name, (run, Lk, _) in shapes.items():
        for _ in range(3): run()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(30)]
        for a, b in ev:
            a.record(); run(); b.record()
        torch.cuda.synchronize()
        us = sorted(a.elapsed_time(b) for a, b in ev)[15] * 1e3
        err = (o[:2].float() - refs[name]).abs().max().item() / refs[name].abs().max().item()
        row.append(f"{name} {us:6.1f} us {4 * T * H * N * Lk * D / us / 1e6:6.1f} TF err {err:.1e}")
    print(" | ".join(row), flush=True)
L.gvf_attn_set_debug(0)
