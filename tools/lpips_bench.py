"""LPIPS-VGG16 forward + backward on the joint train step's image batch (50 renders of 512^2 against 50 targets): ms per
call for a few execution choices.   python tools/lpips_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gvfdiffusion_b200.utils.lpips import LPIPS  # noqa: E402


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    dev = torch.device("cuda", 0)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    x = (torch.rand(n, 3, 512, 512, device=dev) * 2 - 1).requires_grad_(True)
    y = torch.rand(n, 3, 512, 512, device=dev) * 2 - 1
    m = LPIPS().to(dev).eval()

    def step():
        x.grad = None
        m(x, y).backward()

    def fwd_only():
        with torch.no_grad():
            m(x, y)

    for bench in (False, True):
        torch.backends.cudnn.benchmark = bench
        print(f"cudnn.benchmark={bench}: fwd+bwd {timed(step):.1f} ms, fwd only (no grad) {timed(fwd_only):.1f} ms", flush=True)
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:16]:
        print(f"{e.device_time_total / 1e3:8.2f} ms  x{e.count:<4d} {e.key[:110]}")


if __name__ == "__main__":
    main()
