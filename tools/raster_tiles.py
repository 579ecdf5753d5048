"""Per-tile segment length statistics of the benchmark raster scene (what bounds sort_blend)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gvfdiffusion_b200 import raster as R, synthetic as S
dev = "cuda"
F = 24
canon = S.canonical_gaussians(num_voxels=2048)
P = canon["_xyz"].shape[0]
delta = S.raster_delta(F, P).to(dev)
cams, tfx, tfy = R.pack_cameras(S.orbit_extrinsics(F), S.intrinsics(), 0.8, 1.6)
prm = R.make_params(512, 512, tfx, tfy, S.gaussian_constants())
rz = R.Rasterizer(dev)
rgba, radii = rz.forward(prm, R.canon_arrays(canon, dev), delta, cams.to(dev))
torch.cuda.synchronize()
ts = rz.buffer("tile_start", torch.int32, F * 1024 + 1).cpu().long()
n = ts[1:] - ts[:-1]
print("status", rz.status(), "tiles", n.numel(), "mean", n.float().mean().item(), "max", n.max().item())
for q in (0.5, 0.75, 0.9, 0.95, 0.99, 0.999):
    print(f"  q{q}: {n.float().quantile(q).item():.0f}")
print("empty tiles", int((n == 0).sum()), " tiles > 256:", int((n > 256).sum()), " > 1024:", int((n > 1024).sum()), " > 2048:", int((n > 2048).sum()))
print("sum n*1 (blend evals / 256):", int(n.sum()), " sum over tiles of n*log2(n)^2:", float((n.float() * torch.log2(n.float().clamp(min=2)) ** 2).sum()))
r = radii.cpu()
print("radii: mean", r[r > 0].float().mean().item(), "max", r.max().item(), "tiles/gaussian mean", n.sum().item() / max(1, int((r > 0).sum())))
nc = rz.buffer("n_contrib", torch.int32, F * 512 * 512).cpu().long()
print("n_contrib (last contributor index per pixel): mean", nc.float().mean().item(), "max", nc.max().item(), "sum", int(nc.sum()))
