"""Timing experiments on the static cross-attention shape: which part of the softmax loop costs what."""
import math, os, sys
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
for dbg, name in [(0, "full")]:
    L.gvf_attn_set_debug(dbg)
    for _ in range(3): run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"dbg {dbg:2d} {name:22s} {ms * 1e3:7.1f} us  ({4 * T * H * N * 4096 * D / ms / 1e9:6.1f} TFLOP/s)")
L.gvf_attn_set_debug(0)

# clock64 trace of CTA (0,0,0): softmax warp 4 and MMA warp 1, first 16 blocks
tr = torch.zeros(256 + 3 * 1024, dtype=torch.int64, device=dev)
L.gvf_attn_set_trace(_lib.ptr(tr))
run(); torch.cuda.synchronize()
L.gvf_attn_set_trace(None)
t = tr.cpu()[:256].view(16, 16)
cta = tr.cpu()[256:].view(1024, 3)[:768]
t0 = int(t[0, 0])
names = ["loop_top", "s_full_seen", "S_loaded", "half_exp", "o_ld_issued", "fold_done", "P_stored", "-", "MMA:p_full_seen", "MMA:pv_issued", "MMA:qk_issued"]
print("block " + " ".join(f"{n:>16s}" for n in names))
for i in range(12):
    print(f"{i:5d} " + " ".join(f"{(int(t[i, k]) - t0) if int(t[i, k]) else 0:16d}" for k in range(11)))

print('loop_top of blocks 16..31:', [int(t[i, 12]) - t0 for i in range(16)], 'loop end', int(t[0, 13]) - t0)
# per-CTA schedule
import collections
t_min = int(cta[:, 1].min())
dur = (cta[:, 2] - cta[:, 1]).float() / 1e3
print(f"CTA duration us: min {dur.min():.1f} mean {dur.mean():.1f} max {dur.max():.1f}; kernel span {(int(cta[:, 2].max()) - t_min) / 1e3:.1f} us")
per_sm = collections.Counter(cta[:, 0].tolist())
print("CTAs per SM histogram:", sorted(collections.Counter(per_sm.values()).items()), "SMs used", len(per_sm))
start = ((cta[:, 1] - t_min).float() / 1e3)
for w in range(0, 768, 96):
    print(f"cta {w:4d}: start {start[w]:7.1f} us dur {dur[w]:6.1f} us sm {int(cta[w, 0])}")
