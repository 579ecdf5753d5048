#!/bin/bash
python - <<'PY'
import math, torch
from gvfdiffusion_b200 import _lib, ops
L = _lib.lib()
g = torch.Generator().manual_seed(1)
rn = lambda *s: torch.randn(*s, generator=g).cuda().half()
for dbg in (0x80, 0x84):
    L.gvf_attn_set_debug(dbg)
    for (Nb, Lq, Lk, H) in ((2, 512, 512, 16), (3, 512, 1374, 4), (2, 300, 200, 2), (1, 1024, 70, 3), (2, 130, 4096, 2)):
        q, k, v = rn(Nb, Lq, H, 32) * 2, rn(Nb, Lk, H, 32) * 2, rn(Nb, Lk, H, 32)
        o = ops.attention(q, k, v, 1 / math.sqrt(32))
        sc = torch.einsum("bqhd,bkhd->bhqk", q.float(), k.float()) / math.sqrt(32)
        ref = torch.einsum("bhqk,bkhd->bqhd", sc.softmax(-1), v.float())
        err = (o.float() - ref).abs().max().item() / ref.abs().max().item()
        print(hex(dbg), (Nb, Lq, Lk, H), f"rel err {err:.2e}", "OK" if err < 2e-3 else "FAIL")
L.gvf_attn_set_debug(0)
PY
ATTN_DBG=0x84 timeout 100 python tools/attn_experiments.py 2>&1 | tail -34
