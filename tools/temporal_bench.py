"""DiT temporal self-attention at the benchmark shape: warp-level MMA kernel vs the CUDA-core kernel."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gvfdiffusion_b200 import _lib, ops
L = _lib.lib()
T, N, H, D = 24, 512, 16, 32
g = torch.Generator().manual_seed(0)
qkv = torch.randn(T, N, 3, H, D, generator=g).cuda().half()
ao = torch.empty(T, N, H, D, dtype=torch.float16, device="cuda")
tv = qkv.permute(1, 0, 2, 3, 4)
to = ao.permute(1, 0, 2, 3)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
outs = {}
for dbg, name in ((0x40, "cuda cores"), (0, "mma.sync"), (0x10000, "mma.sync pipe")):
    L.gvf_attn_set_debug(dbg)
    for _ in range(3): ops.attention(tv[:, :, 0], tv[:, :, 1], tv[:, :, 2], 1 / math.sqrt(D), out=to)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.attention(tv[:, :, 0], tv[:, :, 1], tv[:, :, 2], 1 / math.sqrt(D), out=to); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    outs[name] = ao.clone()
    tw = []
    for _ in range(10):                       # warm: the previous launch left the operands in L2, as the QKV GEMM does
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.attention(tv[:, :, 0], tv[:, :, 1], tv[:, :, 2], 1 / math.sqrt(D), out=to); e1.record()
        torch.cuda.synchronize(); tw.append(e0.elapsed_time(e1) * 1e3)
    tw.sort()
    print(f"{name:14s} median {ts[5]:6.1f} us cold L2 ({50.3e6 / ts[5] / 1e3:6.0f} GB/s), {tw[5]:6.1f} us warm")
print("pipelined == one-sequence-per-CTA bits:", torch.equal(outs["mma.sync"], outs["mma.sync pipe"]))
L.gvf_attn_set_debug(0)
