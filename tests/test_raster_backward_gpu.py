"""GPU parity of the rasteriser BACKWARD (cfg 3: VAE decode + render fwd+bwd) against torch.autograd
over the differentiable CPU restatement (oracle/raster_torch.py, itself checked against
oracle/raster.c).  Tolerance: relative L2 <= 2e-3 per gradient tensor (fp32 atomics reorder sums)."""
import numpy as np
import pytest
import torch

from oracle import gaussian as OG
from oracle import raster_torch as RT
from tests import _scenes

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-12))


@pytest.mark.parametrize("voxels,F,H,W,boost", [(24, 2, 48, 64, 0.0), (16, 1, 64, 64, 2.0)])
def test_backward_matches_autograd(voxels, F, H, W, boost):
    from gvfdiffusion_b200 import raster as R
    canon, delta, ext, intr, const = _scenes.scene(voxels, F, H, W, seed=3, scale_boost=boost)
    gen = torch.Generator().manual_seed(9)
    wts = torch.randn(F, 4, H, W, generator=gen)
    # ---- oracle: autograd over the torch restatement
    oc = {k: v.clone().requires_grad_(True) for k, v in canon.items()}
    od = delta.clone().requires_grad_(True)
    loss = 0
    imgs = []
    for f in range(F):
        vt, pt, _, tfx, tfy = OG.camera_matrices(ext[f], intr, 0.8, 1.6)
        img = RT.render(oc, od[f], const, vt, pt, H, W, tfx, tfy)
        imgs.append(img.detach())
        loss = loss + (img * wts[f]).sum()
    loss.backward()
    # ---- ours
    cams, tfx, tfy = R.pack_cameras(ext, intr, 0.8, 1.6)
    prm = R.make_params(H, W, tfx, tfy, const)
    rz = R.Rasterizer(DEV)
    P = canon["_xyz"].shape[0]
    t = {k: v.to(DEV).reshape(P, -1).clone().requires_grad_(True) for k, v in canon.items()}
    dd = delta.to(DEV).clone().requires_grad_(True)
    rgba, radii = R.RasterizeFrames.apply(rz, prm, cams.to(DEV), t["_xyz"], t["_features_dc"], t["_scaling"],
                                          t["_rotation"], t["_opacity"].reshape(P), dd)
    assert np.abs(rgba.detach().cpu().numpy() - torch.stack(imgs).numpy()).max() < 1e-3
    (rgba * wts.to(DEV)).sum().backward()
    for k in ("_xyz", "_features_dc", "_scaling", "_rotation", "_opacity"):
        ref = oc[k].grad.reshape(P, -1)
        assert _rel(t[k].grad.reshape(P, -1), ref) < 2e-3, (k, _rel(t[k].grad.reshape(P, -1), ref))
    assert _rel(dd.grad, od.grad) < 2e-3, _rel(dd.grad, od.grad)
    for sl, name in ((slice(0, 3), "xyz"), (slice(3, 6), "scale"), (slice(6, 10), "rot"), (slice(10, 13), "rgb"),
                     (slice(13, 14), "opacity")):
        assert _rel(dd.grad[..., sl], od.grad[..., sl]) < 3e-3, (name, _rel(dd.grad[..., sl], od.grad[..., sl]))


def test_renderer_autograd_path_and_detach_static():
    from gvfdiffusion_b200.renderers import GaussianRenderer
    from tests.test_api_gpu import _model
    canon, delta, ext, intr, const = _scenes.scene(16, 1, 48, 48, seed=5)
    r = GaussianRenderer({"near": 0.8, "far": 1.6, "bg_color": (1.0, 1.0, 1.0)})
    r.pipe.use_mip_gaussian = True
    r.rendering_options.resolution = 48
    gm = _model(canon)
    for k in ("_xyz", "_features_dc", "_scaling", "_rotation", "_opacity"):
        getattr(gm, k).requires_grad_(True)
    d = delta[0].to(DEV).requires_grad_(True)
    res = r.render(gm, ext[0].to(DEV), intr.to(DEV), delta_pc=d, detach_static=True)
    res["rgb"].sum().backward()
    assert d.grad is not None and d.grad.abs().sum() > 0
    assert gm._xyz.grad is None                           # detach_static=True: only delta receives gradients
    res = r.render(gm, ext[0].to(DEV), intr.to(DEV), delta_pc=d, detach_static=False)
    (res["rgb"].sum() + res["alpha"].sum()).backward()
    assert gm._xyz.grad is not None and gm._scaling.grad.abs().sum() > 0
