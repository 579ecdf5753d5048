"""Alignment pre-step (reference utils/inference_utils.py:37-177): the plain-3DGS rasteriser mode
(`pipe.use_mip_gaussian = False`, gvf_raster_params.mip_filter = 0) against the C oracle, and the batched 360-view azimuth
search + in-place rotation on a synthetic object whose "conditioning image" is one of its own views."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _model(num_voxels=256, seed=0):
    from gvfdiffusion_b200 import synthetic as S
    from gvfdiffusion_b200.representations.gaussian import GaussianModel
    canon = S.canonical_gaussians(num_voxels=num_voxels, seed=seed)
    c = S.gaussian_constants()
    gm = GaussianModel(sh_degree=0, aabb=c["aabb"], mininum_kernel_size=c["min_kernel"], scaling_bias=0.004, opacity_bias=0.1,
                       scaling_activation="softplus", device=DEV)
    for k, v in canon.items():
        setattr(gm, k, v.to(DEV))
    return gm, canon, c


def test_plain_3dgs_mode_matches_oracle():
    from gvfdiffusion_b200 import raster as R, synthetic as S
    from oracle import gaussian as OG, raster as OR
    gm, canon, const = _model(128, 2)
    ext, intr = S.orbit_extrinsics(2), S.intrinsics()
    cams, tfx, tfy = R.pack_cameras(ext, intr, 0.8, 1.6)
    for mip, ks in ((False, 0.3), (True, 0.1)):
        prm = R.make_params(96, 96, tfx, tfy, const, ks, 1.0, (1.0, 1.0, 1.0), mip_filter=mip)
        rgba, radii = R.Rasterizer(DEV).forward(prm, R.canon_arrays(canon, DEV), None, cams.to(DEV))
        vt, pt = [], []
        for f in range(2):
            v, p, _, a, b = OG.camera_matrices(ext[f], intr, 0.8, 1.6)
            vt.append(v.numpy())
            pt.append(p.numpy())
        oprm = OR.make_params(96, 96, a, b, const, kernel_size=ks, mip_filter=mip)
        ref, nr, oradii = OR.render_frames(oprm, {k: v.numpy() for k, v in canon.items()}, None, np.stack(vt), np.stack(pt),
                                           want_radii=True)
        assert np.array_equal(radii.cpu().numpy(), oradii)
        assert np.abs(rgba.cpu().numpy() - ref).max() < 2e-3
    # the two modes differ (opacity compensation, 0.3 vs 0.1 dilation)
    p0 = R.make_params(96, 96, tfx, tfy, const, 0.3, 1.0, (1.0, 1.0, 1.0), mip_filter=False)
    p1 = R.make_params(96, 96, tfx, tfy, const, 0.3, 1.0, (1.0, 1.0, 1.0), mip_filter=True)
    a0, _ = R.Rasterizer(DEV).forward(p0, R.canon_arrays(canon, DEV), None, cams.to(DEV))
    a1, _ = R.Rasterizer(DEV).forward(p1, R.canon_arrays(canon, DEV), None, cams.to(DEV))
    assert (a0 - a1).abs().max() > 1e-3


def test_azimuth_search_and_rotation_recover_a_known_view():
    from gvfdiffusion_b200 import synthetic as S
    from gvfdiffusion_b200.renderers.gaussian_render import GaussianRenderer
    from gvfdiffusion_b200.utils import inference_utils as IU
    gm, canon, _ = _model(512, 4)
    rd = GaussianRenderer({"resolution": 512, "near": 0.8, "far": 1.6, "ssaa": 1, "bg_color": (1.0, 1.0, 1.0)})
    rd.pipe.use_mip_gaussian = True
    intr = S.intrinsics()
    true_azi = 37
    rd.pipe.use_mip_gaussian = False
    view = rd.render(gm, IU.orbit_extrinsics([true_azi])[0].to(DEV), intr.to(DEV))
    rd.pipe.use_mip_gaussian = True
    image, alpha = view["rgb"].clamp(0, 1), view["alpha"]
    azi, scale, table = IU.find_best_azimuth(gm, rd, image, alpha, intr, in_the_wild=True)
    assert rd.pipe.use_mip_gaussian is True                      # restored
    assert len(table) == 360 and azi == true_azi and abs(scale - 1.0) < 1e-6
    assert min(r[1] for r in table) < 1e-6                       # that view reproduces the image
    # dataset mode: every 90 degrees
    azi90, _, t90 = IU.find_best_azimuth(gm, rd, image, alpha, intr, in_the_wild=False)
    assert len(t90) == 4 and azi90 in (-180, -90, 0, 90)
    # rotate in place: the best view becomes the front view (azimuth 0)
    IU.align_gaussian_to_canonical(gm, image, alpha, intr, rd)
    rd.pipe.use_mip_gaussian = False
    front = rd.render(gm, IU.orbit_extrinsics([0])[0].to(DEV), intr.to(DEV))["rgb"].clamp(0, 1)
    assert float((front - image).abs().mean()) < 2e-3


def test_ssaa_and_patch_mask_follow_the_reference():
    from gvfdiffusion_b200 import synthetic as S
    from gvfdiffusion_b200.renderers.gaussian_render import GaussianRenderer
    import torch.nn.functional as F
    gm, _, _ = _model(128, 6)
    ext, intr = S.orbit_extrinsics(1)[0].to(DEV), S.intrinsics().to(DEV)
    opts = {"resolution": 64, "near": 0.8, "far": 1.6, "bg_color": (1.0, 1.0, 1.0)}
    lo = GaussianRenderer(dict(opts, ssaa=1))
    hi = GaussianRenderer(dict(opts, ssaa=2))
    big = GaussianRenderer(dict(opts, resolution=128, ssaa=1))
    for r in (lo, hi, big):
        r.pipe.use_mip_gaussian = True
    out = hi.render(gm, ext, intr, patch_mask=torch.ones(64, 64, device=DEV))
    ref = F.interpolate(big.render(gm, ext, intr)["rgb"][None], size=(64, 64), mode="bicubic", align_corners=False,
                        antialias=True).squeeze()
    assert out["rgb"].shape == (3, 64, 64) and torch.equal(out["rgb"], ref)
    assert lo.render(gm, ext, intr)["rgb"].shape == (3, 64, 64)
