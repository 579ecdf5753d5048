"""Farthest point sampling at the benchmark size (16384 points -> 4096 samples): cluster kernel vs single CTA."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
if len(sys.argv) > 1:
    from gvfdiffusion_b200 import ops
    g = torch.Generator().manual_seed(0)
    outs = []
    for P in (16384, 5000, 16001):
        pts = torch.randn(P, 14, generator=g).cuda()
        K = min(4096, P)
        for _ in range(2): idx = ops.fps(pts, K)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): idx = ops.fps(pts, K)
        e1.record(); torch.cuda.synchronize()
        print(f"{sys.argv[1]:8s} P={P:6d} K={K}: {e0.elapsed_time(e1) / 5:7.3f} ms  checksum {int(idx.long().sum())} first {idx[:6].tolist()}")
else:
    for mode in ("cluster", "smem", "v2"):   # v2 = default
        subprocess.run([sys.executable, __file__, mode], env=dict(os.environ, GVF_FPS=mode))
