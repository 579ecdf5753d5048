"""Windowed sparse attention at a static-VAE-like shape: fused gather/scatter kernel vs the reference's data
flow restated in torch (index_select gather -> per-window SDPA via a padded batch -> scatter)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gvfdiffusion_b200.sparse.attention import sparse_windowed_scaled_dot_product_self_attention
from gvfdiffusion_b200.sparse.attention.windowed_attn import _partition
g = torch.Generator().manual_seed(0)
B, res, H, D = 8, 64, 12, 64
coords = []
for b in range(B):                                     # a radius-0.35 shell of ~2000 voxels per object, as in synthetic.py
    p = torch.randn(60000, 3, generator=g); p = p / p.norm(dim=1, keepdim=True) * 0.35 * res + res / 2
    v = torch.unique(p.long().clamp(0, res - 1), dim=0)[:2048]
    coords.append(torch.cat([torch.full((v.shape[0], 1), b), v], 1))
coords = torch.cat(coords).int().cuda()
T = coords.shape[0]
qkv = torch.randn(T, 3, H, D, generator=g).half().cuda()
fwd, bwd, cu, mx = _partition(coords, 8, (0, 0, 0))
W = cu.shape[0] - 1
lens = (cu[1:] - cu[:-1]).long()
print(f"T={T} windows={W} mean len {lens.float().mean():.1f} max {mx}")
def ours(): return sparse_windowed_scaled_dot_product_self_attention(qkv, coords, 8, (0, 0, 0))
def torch_flow():
    x = qkv[fwd.long()]                                                     # gather copy
    pad = torch.zeros(W, mx, 3, H, D, dtype=torch.float16, device="cuda")
    pos = torch.arange(T, device="cuda") - cu[:-1].long().repeat_interleave(lens)
    wid = torch.arange(W, device="cuda").repeat_interleave(lens)
    pad[wid, pos] = x
    mask = (torch.arange(mx, device="cuda")[None] < lens[:, None])[:, None, None, :]
    q, k, v = (pad[:, :, i].transpose(1, 2) for i in range(3))
    o = torch.nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=mask).transpose(1, 2)
    return o[wid, pos][bwd]                                                 # un-pad + scatter copy
a, b = ours().float(), torch_flow().float()
print("max abs diff vs torch flow", (a - b).abs().max().item())
for name, fn in (("fused kernel", ours), ("torch flow", torch_flow)):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    fl = 4 * H * D * float((lens.double() ** 2).sum())
    print(f"{name:14s} {us:8.1f} us   ({fl / us / 1e6:7.1f} TFLOP/s algorithmic, {T * H * D * 2 * 4 / us / 1e3:7.1f} GB/s of qkv+out)")
