"""Sweep the split-K factor of gvf_gemm_tn_f16 (weight gradients) per shape: us per launch for ksplit = 1..N and what the
library's automatic choice picks.   python tools/gemm_tn_ksplit_sweep.py > profiles/rNN_gemm_tn_ksplit.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gvfdiffusion_b200 import _lib, ops  # noqa: E402

SHAPES = [(768, 3072, 4096), (3072, 768, 4096), (768, 768, 4096), (2304, 768, 4096), (768, 1024, 4096),
          (768, 3072, 12288), (6144, 768, 12288), (768, 768, 12288), (2304, 768, 12288), (768, 768, 393216), (1536, 768, 12288)]


def timed(fn, reps=20):
    for _ in range(3):
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
    L = _lib.lib()
    for M, N, R in SHAPES:
        a = (torch.randn(R, M, device=dev) * 0.1).half()
        w = (torch.randn(R, N, device=dev) * 0.1).half()
        out = torch.empty(M, N, device=dev)
        row = []
        for ks in (1, 2, 3, 4, 5, 6, 8, 10, 12, 16, 24, 32):
            if ks > max(1, (R // 64) // 2):
                continue
            L.gvf_gemm_set_ksplit(ks)
            row.append((ks, timed(lambda: ops.gemm_tn(a, w, out=out))))
        L.gvf_gemm_set_ksplit(0)
        auto = timed(lambda: ops.gemm_tn(a, w, out=out))
        best = min(row, key=lambda kv: kv[1])
        gf = 2.0 * M * N * R / 1e9
        print(f"M={M} N={N} R={R}: auto {auto:.1f} us ({gf / auto:.0f} TF/s) best ks={best[0]} {best[1]:.1f} us | " +
              " ".join(f"{k}:{t:.1f}" for k, t in row), flush=True)


if __name__ == "__main__":
    main()
