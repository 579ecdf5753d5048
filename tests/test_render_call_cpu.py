"""Pins the rasteriser BOUNDARY to the reference: tests/golden/render_call.pt holds what the reference's own
GaussianRenderer.render (renderers/gaussian_render.py:85-238,269-369) hands to the third-party
diff_gaussian_rasterization -- the GaussianRasterizationSettings fields and the tensors of the rasteriser call --
recorded by running that code on the CPU with a recording stand-in (tests/golden/make_golden.py gen_render_call).
The rasteriser arithmetic itself stays "parity unpinned" (third party, absent); everything up to its call is
checked here, for the oracle and for the host mirror that packs the cameras for libgvf_b200.so."""
import os

import torch

from gvfdiffusion_b200 import raster as R
from gvfdiffusion_b200 import synthetic as S
from oracle import gaussian as OG

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load():
    return torch.load(os.path.join(G, "render_call.pt"), weights_only=False)


def test_oracle_camera_setup_is_the_reference_call():
    g = _load()
    s = g["with_delta"]["settings"]
    vt, pt, campos, tfx, tfy = OG.camera_matrices(g["extrinsics"], g["intrinsics"], g["near"], g["far"])
    assert torch.equal(vt, s["viewmatrix"]) and torch.equal(pt, s["projmatrix"]) and torch.equal(campos, s["campos"])
    assert tfx == s["tanfovx"] and tfy == s["tanfovy"]
    assert (s["image_height"], s["image_width"], s["sh_degree"], s["prefiltered"]) == (64, 64, 0, False)
    assert s["kernel_size"] == 0.1 and s["scale_modifier"] == 1.0
    assert torch.equal(s["bg"], torch.ones(3)) and float(s["subpixel_offset"].abs().max()) == 0.0
    assert tuple(s["subpixel_offset"].shape) == (64, 64, 2)


def test_host_mirror_packs_the_same_cameras():
    g = _load()
    s = g["with_delta"]["settings"]
    cams, tfx, tfy = R.pack_cameras(g["extrinsics"], g["intrinsics"], g["near"], g["far"])
    assert cams.shape == (1, 32)
    assert torch.equal(cams[0, :16].reshape(4, 4), s["viewmatrix"])
    assert torch.equal(cams[0, 16:].reshape(4, 4), s["projmatrix"])
    assert tfx == s["tanfovx"] and tfy == s["tanfovy"]
    prm = R.make_params(64, 64, tfx, tfy, S.gaussian_constants(), 0.1, 1.0, (1.0, 1.0, 1.0))
    assert (prm.H, prm.W) == (64, 64) and abs(prm.kernel_size - 0.1) < 1e-7 and list(prm.bg) == [1.0, 1.0, 1.0]
    assert abs(prm.tanfovx - s["tanfovx"]) < 1e-7


def test_oracle_activation_is_what_the_reference_passes_to_the_rasteriser():
    g = _load()
    const = S.gaussian_constants()
    for tag, delta in (("with_delta", g["delta"]), ("no_delta", None)):
        c = g[tag]["call"]
        xyz, scales, rots, shs, opac = OG.activate(g["raw"], delta, const)
        assert c["colors_precomp"] is None and c["cov3D_precomp"] is None
        assert torch.allclose(xyz, c["means3D"], rtol=0, atol=1e-7)
        assert torch.allclose(scales, c["scales"], rtol=1e-6, atol=1e-9)
        assert torch.allclose(rots, c["rotations"], rtol=1e-6, atol=1e-7)
        assert torch.equal(shs, c["shs"]) and tuple(c["shs"].shape) == (96, 1, 3)
        assert torch.allclose(opac, c["opacities"], rtol=1e-6, atol=1e-8)
        assert tuple(c["means2D"].shape) == (96, 3) and float(c["means2D"].abs().max()) == 0.0
