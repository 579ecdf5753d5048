"""Farthest point sampling at the benchmark size (16384 points -> 4096 samples): cluster kernel vs single CTA."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
if len(sys.argv) > 1:
    from gvfdiffusion_b200 import ops
    g = torch.Generator().manual_seed(0)
    outs = []
    from gvfdiffusion_b200 import synthetic
    shell = synthetic.canonical_gaussians()["_xyz"]
    clouds = [("shell voxel-major", shell), ("shell shuffled", shell[torch.randperm(16384, generator=g)])]
    clouds += [(f"gaussian blob", torch.randn(P, 14, generator=g)) for P in (16384, 5000, 16001)]
    for name, pts in clouds:
        P = pts.shape[0]
        pts = pts.contiguous().cuda()
        K = min(4096, P)
        for _ in range(2): idx = ops.fps(pts, K)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): idx = ops.fps(pts, K)
        e1.record(); torch.cuda.synchronize()
        print(f"{sys.argv[1]:8s} {name:18s} P={P:6d} K={K}: {e0.elapsed_time(e1) / 5:7.3f} ms  checksum {int(idx.long().sum())} first {idx[:6].tolist()}")
else:
    for mode in ("cluster", "smem", "v2", "pruned"):   # v2 = gvf_fps default, pruned = what gvf_fps_ordered runs
        subprocess.run([sys.executable, __file__, mode], env=dict(os.environ, GVF_FPS=mode))
