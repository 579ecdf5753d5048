"""Times the rasteriser alone: F frames x P Gaussians, CUDA events on the current stream."""
import argparse, json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gvfdiffusion_b200 import raster as R, synthetic as S

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=24)
ap.add_argument("--voxels", type=int, default=2048)
ap.add_argument("--res", type=int, default=512)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--backward", action="store_true", help="also time gvf_raster_backward (gradients to raw parameters and delta)")
ap.add_argument("--sort", type=int, default=-1, help="-1 env, 0 bitonic network, 1 bucket sort")
a = ap.parse_args()
dev = "cuda"
from gvfdiffusion_b200 import _lib
_lib.lib().gvf_raster_set_sort(a.sort)
canon = S.canonical_gaussians(num_voxels=a.voxels)
P = canon["_xyz"].shape[0]
delta = S.raster_delta(a.frames, P).to(dev)
cams, tfx, tfy = R.pack_cameras(S.orbit_extrinsics(a.frames), S.intrinsics(), 0.8, 1.6)
cams = cams.to(dev)
prm = R.make_params(a.res, a.res, tfx, tfy, S.gaussian_constants())
rz = R.Rasterizer(dev)
arrays = R.canon_arrays(canon, dev)
out = torch.empty((a.frames, 4, a.res, a.res), device=dev)
for _ in range(3):
    rz.forward(prm, arrays, delta, cams, out=out, want_radii=False)
Rn, ovf, longest = rz.status()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for _ in range(a.iters):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rz.forward(prm, arrays, delta, cams, out=out, want_radii=False, check_overflow=False)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
ms = ts[len(ts) // 2]
alg = a.frames * (112 * P + 16 * a.res * a.res) + 64 * Rn
bw = {}
if a.backward:
    g = torch.randn(a.frames, 4, a.res, a.res, device=dev)
    rz.forward(prm, arrays, delta, cams, out=out, want_radii=False, check_overflow=False)
    for _ in range(2):
        rz.backward(prm, arrays, delta, cams, g)
    tb = []
    for _ in range(a.iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rz.backward(prm, arrays, delta, cams, g)
        e1.record()
        torch.cuda.synchronize()
        tb.append(e0.elapsed_time(e1))
    tb.sort()
    alg_b = alg + a.frames * (16 * a.res * a.res + 112 * P) + 40 * Rn      # A_bwd
    bw = {"bwd_ms_median": tb[len(tb) // 2], "bwd_ms_min": tb[0], "bwd_alg_bytes": alg_b,
          "bwd_GBps": alg_b / tb[len(tb) // 2] / 1e6}
print(json.dumps({**bw, "sort": a.sort, "frames": a.frames, "P": P, "res": a.res, "num_rendered": Rn, "overflow": ovf,
                  "ms_median": ms, "ms_min": ts[0], "alg_bytes": alg, "GBps": alg / ms / 1e6,
                  "frames_per_s": a.frames / ms * 1e3}))
