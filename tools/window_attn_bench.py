"""Window attention forward at the static VAE's shape (2 x 2048 surface voxels, 12 heads): us per launch for the per-window
tiling, the packed tiling with cp.async staging and the packed tiling with TMA tile::gather4 staging.
    python tools/window_attn_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gvfdiffusion_b200 import _lib  # noqa: E402
from gvfdiffusion_b200.sparse.attention import sparse_windowed_scaled_dot_product_self_attention as attn  # noqa: E402
from tools import static_vae_step_bench as SB  # noqa: E402


def timed(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    dev = torch.device("cuda", 0)
    coords = torch.cat([torch.cat([torch.full((SB.NVOX, 1), b), SB.surface_voxels(1 + b)], 1) for b in range(2)]).int().to(dev)
    qkv = (torch.randn(coords.shape[0], 3, 12, 64, device=dev) * 0.8).half()
    for shift in ((0, 0, 0), (4, 4, 4)):
        ref = attn(qkv, coords, 8, shift, packed=False)
        row = [f"per-window {timed(lambda: attn(qkv, coords, 8, shift, packed=False)):.1f} us"]
        for tma in (0, 1):
            _lib.lib().gvf_sparse_attn_set_tma(tma)
            out = attn(qkv, coords, 8, shift, packed=True)
            err = float((out.float() - ref.float()).abs().max())
            row.append(f"packed {'tma' if tma else 'cp.async'} {timed(lambda: attn(qkv, coords, 8, shift, packed=True)):.1f} us (max diff {err:.1e})")
        print(f"shift {shift}: " + " | ".join(row))


if __name__ == "__main__":
    main()
