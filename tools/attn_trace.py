"""clock64 trace of CTA 0 of the static cross-attention launch (v6 kernel, TRACE build):
    python tools/attn_trace.py [debug word ...]
columns per softmax block i of tile 0 (clocks relative to its first block top):
  top | s_full seen | S in regs | exps done | o_full seen | P stored+arrived || issuer: kv ready | s_free seen (QK i+1) | p_full seen (PV i)
then the block tops of tiles 1..3."""
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
qkv = rn(T, N, 3, H, D)
SHAPE = os.environ.get("ATTN_SHAPE", "static")
def run():
    if SHAPE == "spatial": ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], 1 / math.sqrt(D), out=o)
    else: ops.attention(q, kv[:, 0], kv[:, 1], 1 / math.sqrt(D), out=o, kv_shared=True)
for dbg in [int(v, 0) for v in sys.argv[1:]] or [0x81]:
    L.gvf_attn_set_debug(dbg)
    run(); torch.cuda.synchronize()
    tr = torch.zeros(256 + 3 * 1024, dtype=torch.int64, device=dev)
    L.gvf_attn_set_trace(_lib.ptr(tr))
    run(); torch.cuda.synchronize()
    L.gvf_attn_set_trace(None)
    t = tr.cpu()[:256].view(16, 16)
    t0 = int(t[0, 0])
    print(f"--- dbg {dbg:#x} {SHAPE}: loop end {int(t[0, 13]) - t0}")
    e = tr.cpu()[240:245] - t0
    print(f"    CTA 0 (clocks rel. to the first block top): entry {int(e[0])}, barriers + TMEM ready {int(e[1])}, Q and K/V tile 0 in smem {int(e[2])}, "
          f"output stored {int(e[3])}, TMEM released {int(e[4])}")
    print("  i      top   s_full   S_regs  max_done     exps   o_full  P_store ||   kv_rdy   s_free   p_full || tops of tiles 1..3")
    for i in range(16):
        c = lambda k: int(t[i, k]) - t0
        print(f"{i:3d} {c(0):8d} {c(10):8d} {c(4):8d} {c(12):8d} {c(5):8d} {c(11):8d} {c(6):8d} || {c(9):8d} {c(7):8d} {c(8):8d} || {c(1):8d} {c(2):8d} {c(3):8d}")
L.gvf_attn_set_debug(0)
