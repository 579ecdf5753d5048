"""Directional finite-difference check of the rasteriser backward at several sizes (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gvfdiffusion_b200 import raster as R, synthetic as S
DEV = "cuda"
for (F, vox, H, shift) in [(2, 128, 128, 0.0), (2, 128, 128, 3.0), (24, 2048, 512, 0.0), (24, 2048, 512, 3.0), (4, 2048, 512, 3.0)]:
    W = H
    canon = S.canonical_gaussians(num_voxels=vox, seed=0)
    canon["_features_dc"] = canon["_features_dc"] + shift          # shift > 0: no colour reaches the clamp at 0
    P = canon["_xyz"].shape[0]
    delta = S.raster_delta(F, P).to(DEV)
    cams, tfx, tfy = R.pack_cameras(S.orbit_extrinsics(F), S.intrinsics(), 0.8, 1.6)
    prm = R.make_params(H, W, tfx, tfy, S.gaussian_constants())
    rz = R.Rasterizer(DEV)
    arrays = R.canon_arrays(canon, DEV)
    cams = cams.to(DEV)
    ys, xs = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
    w1 = torch.stack([torch.sin(3 * xs + f) + 1.5 for f in range(4)])[None].expand(F, -1, -1, -1).contiguous().to(DEV)
    rz.forward(prm, arrays, delta, cams, want_radii=False)
    outs, g1, _ = rz.backward(prm, arrays, delta, cams, w1)
    g1 = g1.clone()
    def loss(d):
        rgba, _ = rz.forward(prm, arrays, d.contiguous(), cams, want_radii=False)
        return float((rgba.double() * w1.double()).sum())
    gen = torch.Generator().manual_seed(4)
    for name, sl, eps in (("rgb", slice(10, 13), 2e-2), ("xyz", slice(0, 3), 2e-4), ("opacity", slice(13, 14), 2e-2), ("scale", slice(3, 6), 2e-3)):
        d = torch.zeros_like(delta)
        d[..., sl] = torch.randn(delta[..., sl].shape, generator=gen).to(DEV)
        fd = (loss(delta + eps * d) - loss(delta - eps * d)) / (2 * eps)
        an = float((g1.double() * d.double()).sum())
        # same-sign direction: d = sign(g) removes the random cancellation
        d2 = torch.zeros_like(delta); d2[..., sl] = torch.sign(g1[..., sl])
        fd2 = (loss(delta + eps * d2) - loss(delta - eps * d2)) / (2 * eps)
        an2 = float((g1.double() * d2.double()).sum())
        print(f"F={F} P={P} H={H} shift={shift} {name:8s} random: fd {fd:12.4f} an {an:12.4f} rel {abs(fd-an)/abs(an):.3e} | sign: fd {fd2:12.4f} an {an2:12.4f} rel {abs(fd2-an2)/abs(an2):.3e}")
