"""Timing experiments on the static cross-attention shape (24 x 512 queries, 4096 shared keys, 16 heads, d 32):
kernel generations / exponential mix, clock64 trace of CTA 0, per-CTA schedule."""
import collections, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gvfdiffusion_b200 import _lib, ops
L = _lib.lib()
dev = "cuda"
g = torch.Generator().manual_seed(0)
rn = lambda *s: torch.randn(*s, generator=g).to(dev).half()
T, N, H, D = 24, 512, 16, 32
q, kv = rn(T, N, H, D), rn(4096, 2, H, D)
o = torch.empty(T, N, H, D, dtype=torch.float16, device=dev)
def run(): ops.attention(q, kv[:, 0], kv[:, 1], 1 / math.sqrt(D), out=o, kv_shared=True)
# fp32 reference of the first two frames
sc = torch.einsum("tnhd,khd->thnk", q[:2].float(), kv[:, 0].float()) / math.sqrt(D)
ref = torch.einsum("thnk,khd->tnhd", sc.softmax(-1), kv[:, 1].float())
for dbg, name in [(0x30, "v4 pingpong"), (0x81, "v6 mufu"), (0x84, "v6 1/4 poly"), (0x88, "v6 1/8 poly"), (0, "default (1/8 poly)")]:
    L.gvf_attn_set_debug(dbg)
    for _ in range(3): run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    err = (o[:2].float() - ref).abs().max().item() / ref.abs().max().item()
    print(f"{name:14s} {ms * 1e3:7.1f} us ({4 * T * H * N * 4096 * D / ms / 1e9:6.1f} TFLOP/s)  max err / max |ref| {err:.2e}")
L.gvf_attn_set_debug(int(os.environ.get("ATTN_DBG", "0"), 0))
tr = torch.zeros(256 + 3 * 1024, dtype=torch.int64, device=dev)
L.gvf_attn_set_trace(_lib.ptr(tr))
run(); torch.cuda.synchronize()
L.gvf_attn_set_trace(None)
L.gvf_attn_set_debug(0)
t = tr.cpu()[:256].view(16, 16)
cta = tr.cpu()[256:].view(1024, 3)[:768]
t0 = int(t[0, 0])
print("v6 trace of CTA 0: block top of tiles 0..3 | tile 0: S loaded, exps done, P stored")
for i in range(16):
    print(f"{i:3d} " + " ".join(f"{int(t[i, k]) - t0:8d}" for k in range(7)))
print("loop end", int(t[0, 13]) - t0)
t_min = int(cta[:, 1].min())
dur = (cta[:, 2] - cta[:, 1]).float() / 1e3
print(f"CTA duration us: min {dur.min():.1f} mean {dur.mean():.1f} max {dur.max():.1f}; kernel span {(int(cta[:, 2].max()) - t_min) / 1e3:.1f} us")
per_sm = collections.Counter(cta[:, 0].tolist())
print("CTAs per SM histogram:", sorted(collections.Counter(per_sm.values()).items()), "SMs used", len(per_sm))
