"""Timing of the training-step loss kernels (csrc/losses.cu) on one B200 against the torch formulation the
reference runs (five depthwise conv2d + elementwise + autograd; cdist + topk for the neighbour search).
    python tools/loss_bench.py
"""
import math
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gvfdiffusion_b200 import train_vae as TV          # noqa: E402
from gvfdiffusion_b200.utils import loss_util as LU   # noqa: E402


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def torch_ssim(a, b):
    C = a.size(-3)
    g = torch.tensor([math.exp(-(x - 5) ** 2 / 4.5) for x in range(11)], device=a.device)
    g = (g / g.sum()).unsqueeze(1)
    w = g.mm(g.t()).expand(C, 1, 11, 11).contiguous()
    conv = lambda x: F.conv2d(x, w, padding=5, groups=C)
    mu1, mu2 = conv(a), conv(b)
    s1, s2, s12 = conv(a * a) - mu1 ** 2, conv(b * b) - mu2 ** 2, conv(a * b) - mu1 * mu2
    return (((2 * mu1 * mu2 + 1e-4) * (2 * s12 + 9e-4)) / ((mu1 ** 2 + mu2 ** 2 + 1e-4) * (s1 + s2 + 9e-4))).mean()


def main():
    torch.manual_seed(0)
    shape = (16, 3, 512, 512)                           # bs 2 x 8 cameras (BASELINE configs[4])
    gt = torch.rand(shape, device="cuda")
    pred = (gt + 0.1 * torch.randn(shape, device="cuda")).requires_grad_(True)
    n = gt.numel()

    def ours():
        s, l1 = LU.ssim_l1(pred, gt)
        ((1 - s) * 0.2 + l1).backward()
        pred.grad = None

    def ref():
        ((1 - torch_ssim(pred, gt)) * 0.2 + torch.abs(pred - gt).mean()).backward()
        pred.grad = None

    def ours_fwd():
        with torch.no_grad():
            LU.ssim_l1(pred, gt)

    t_o, t_r, t_f = timeit(ours), timeit(ref), timeit(ours_fwd)
    print(f"ssim+l1 fwd+bwd {shape}: fused {t_o:.1f} us (fwd only {t_f:.1f} us = {n * 8 / t_f / 1e3:.0f} GB/s of the 8 B/px "
          f"algorithmic read; fwd+bwd moves {n * 44 / 1e6:.0f} MB -> {n * 44 / t_o / 1e3:.0f} GB/s), torch conv2d formulation {t_r:.1f} us "
          f"({t_r / t_o:.1f}x)")
    p2 = torch.rand(2, 8192, 3, device="cuda") - 0.5
    p1 = torch.rand(2, 16384, 3, device="cuda") - 0.5
    t_k = timeit(lambda: TV.knn_points(p1, p2, K=4))
    t_c = timeit(lambda: torch.cdist(p1, p2).topk(4, largest=False))
    print(f"knn 2 x 16384 queries x 8192 refs, K=4: gvf_knn {t_k:.1f} us, torch cdist+topk {t_c:.1f} us ({t_c / t_k:.1f}x)")


if __name__ == "__main__":
    main()
