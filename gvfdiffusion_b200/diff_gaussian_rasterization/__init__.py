"""Call-compatible stand-in for `diff_gaussian_rasterization` (mip-splatting fork) as the
reference uses it at renderers/gaussian_render.py:105-125,143,198-206:

    settings = GaussianRasterizationSettings(image_height, image_width, tanfovx, tanfovy, kernel_size,
                                             subpixel_offset, bg, scale_modifier, viewmatrix, projmatrix,
                                             sh_degree, campos, prefiltered, debug)
    color, radii = GaussianRasterizer(settings)(means3D, means2D, shs=..., opacities=..., scales=..., rotations=...)

The tensors are the ACTIVATED rasteriser inputs and go through gvf_raster_forward / gvf_raster_backward
(activated=1); like upstream's autograd Function the call is differentiable with respect to means3D, shs or
colors_precomp, opacities, scales and rotations, and `means2D.grad` receives the screen-space gradient."""
from typing import NamedTuple

import torch
import torch.nn as nn

from .. import raster as R

SH_C0 = 0.28209479177387814


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    kernel_size: float
    subpixel_offset: torch.Tensor
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings
        self._rz = None

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        rs = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        if cov3D_precomp is not None:
            raise NotImplementedError("cov3D_precomp is not used by the reference call site")
        if rs.sh_degree != 0:
            raise NotImplementedError("sh_degree 0 only")
        dev = means3D.device
        P = means3D.shape[0]
        cams = torch.cat([rs.viewmatrix.reshape(1, 16), rs.projmatrix.reshape(1, 16)], 1).to(dev, torch.float32).contiguous()
        prm = R.make_params(rs.image_height, rs.image_width, rs.tanfovx, rs.tanfovy, None, rs.kernel_size,
                            rs.scale_modifier, tuple(float(b) for b in rs.bg.tolist()))
        if self._rz is None:
            self._rz = R.Rasterizer(dev)
        sub = rs.subpixel_offset
        sub = None if sub is None else sub.detach().to(dev, torch.float32).contiguous()
        # colours: sh_degree 0 evaluates clamp_min(SH_C0 * dc + 0.5, 0) (renderers/sh_utils.py:57-113);
        # colors_precomp is mapped onto the same input (identical for colours >= 0)
        dc = shs if shs is not None else (colors_precomp - 0.5) / SH_C0
        diff = torch.is_grad_enabled() and any(t is not None and t.requires_grad
                                               for t in (means3D, means2D, dc, opacities, scales, rotations))
        if diff:
            return R.RasterizeActivated.apply(self._rz, prm, cams, sub, means3D, means2D, dc, opacities, scales,
                                              rotations)
        f = lambda t, n: t.detach().to(dev, torch.float32).reshape(1, P, n).contiguous()
        arrays = (f(means3D, 3), f(dc, 3), f(scales, 3), f(rotations, 4), f(opacities, 1).reshape(1, P))
        rgba, radii = self._rz.forward(prm, arrays, None, cams, activated=True, subpixel_offset=sub)
        return rgba[0, :3], radii[0]
