"""`GaussianRenderer`: drop-in for the reference's renderers/gaussian_render.py:240-369 (and its
duplicate renderers/gaussian_render_all_delta.py) on the mip-Gaussian path
(`pipe.use_mip_gaussian=True`, the only one the inference loop uses):

    renderer.render(gaussian, extrinsics, intrinsics, delta_pc=None, ...) -> {'rgb': (3,H,W), 'alpha': (H,W)}

plus `render_frames` for F (frame, camera) pairs in one launch sequence -- the loop of
utils/inference_utils.py:256-269 batched.  Camera set-up, GaussianModel activations with delta
and the rasteriser all run in libgvf_b200.so (gvf_raster_forward)."""
import numpy as np
import torch

from .. import raster as R


class edict(dict):
    """Minimal attribute dict (the reference uses easydict.EasyDict)."""
    __getattr__ = dict.get

    def __setattr__(self, k, v):
        self[k] = v


class GaussianRenderer:
    def __init__(self, rendering_options={}) -> None:
        self.pipe = edict({"use_mip_gaussian": False, "kernel_size": 0.1, "convert_SHs_python": False,
                           "compute_cov3D_python": False, "scale_modifier": 1.0, "debug": False})
        self.rendering_options = edict({"resolution": None, "near": None, "far": None, "ssaa": 1, "bg_color": "random"})
        self.rendering_options.update(rendering_options)
        self.bg_color = None
        self._rz = None

    def _bg(self):
        if self.rendering_options["bg_color"] == "random":
            return (1.0, 1.0, 1.0) if np.random.rand() < 0.5 else (0.0, 0.0, 0.0)
        return tuple(float(c) for c in self.rendering_options["bg_color"])

    def render_frames(self, gausssian, extrinsics, intrinsics, delta_pc=None, detach_static=False):
        """extrinsics (F,4,4), intrinsics (3,3)|(F,3,3), delta_pc (F,P,14)|None -> rgba (F,4,H,W), radii (F,P).
        Differentiable w.r.t. delta_pc and (unless detach_static) the GaussianModel's raw tensors."""
        # pipe.use_mip_gaussian False = the `diff_gauss` rasteriser of the reference (renderers/gaussian_render.py:126-141,
        # the alignment pre-step sets it at utils/inference_utils.py:50): no mip 2-D filter / opacity compensation, the
        # plain 3DGS screen-space dilation of 0.3 px^2 instead -- same kernels, gvf_raster_params.mip_filter = 0
        mip = bool(self.pipe.use_mip_gaussian)
        if self.pipe.convert_SHs_python or self.pipe.compute_cov3D_python:
            raise NotImplementedError("convert_SHs_python / compute_cov3D_python are not used by the reference configs")
        opt = self.rendering_options
        dev = gausssian._xyz.device
        res = int(opt["resolution"]) * int(opt["ssaa"])        # ssaa > 1: rendered large, reduced by render()
        bg = self._bg()
        from ..representations.gaussian.gaussian_model import _device_const
        self.bg_color = _device_const(bg, dev)
        cams, tfx, tfy = R.pack_cameras(extrinsics, intrinsics, opt["near"], opt["far"])
        cams = R.to_device_async(cams, dev)
        prm = R.make_params(res, res, tfx, tfy, gausssian.constants(), self.pipe.kernel_size if mip else 0.3,
                            self.pipe.scale_modifier, bg, mip_filter=mip)
        if self._rz is None:
            self._rz = R.Rasterizer(dev)
        raw = gausssian.raw()
        needs_grad = torch.is_grad_enabled() and (
            (delta_pc is not None and delta_pc.requires_grad) or
            (not detach_static and any(t.requires_grad for t in raw.values())))
        if needs_grad:
            P = raw["_xyz"].shape[0]
            t = {k: (v.detach() if detach_static else v) for k, v in raw.items()}
            return R.RasterizeFrames.apply(self._rz, prm, cams.to(dev), t["_xyz"].reshape(P, 3).float(),
                                           t["_features_dc"].reshape(P, 3).float(), t["_scaling"].reshape(P, 3).float(),
                                           t["_rotation"].reshape(P, 4).float(), t["_opacity"].reshape(P).float(),
                                           None if delta_pc is None else delta_pc.float())
        d = None if delta_pc is None else delta_pc.detach().to(dev, torch.float32).contiguous()
        return self._rz.forward(prm, R.canon_arrays(raw, dev), d, cams.to(dev))

    def render(self, gausssian, extrinsics, intrinsics, delta_pc=None, detach_static=False, colors_overwrite=None,
               patch_mask=None):
        if colors_overwrite is not None:
            raise NotImplementedError("colors_overwrite is not used on the inference path")
        d = None if delta_pc is None else delta_pc[None]
        if d is not None and d.shape[-1] == 10:        # xyz/scale/rot only (gaussian_render.py:158)
            d = torch.cat([d, torch.zeros(d.shape[:-1] + (4,), device=d.device, dtype=d.dtype)], -1)
        rgba, radii = self.render_frames(gausssian, extrinsics[None], intrinsics, d, detach_static=detach_static)
        rgb, alpha = rgba[0, :3], rgba[0, 3]
        ssaa, res = int(self.rendering_options["ssaa"]), int(self.rendering_options["resolution"])
        if ssaa > 1:
            # renderers/gaussian_render.py:355-358: the super-sampled image is reduced with torch's antialiased bicubic
            # filter (the same library call as the reference; alpha / depth are not resampled there either)
            import torch.nn.functional as F
            rgb = F.interpolate(rgb[None], size=(res, res), mode="bicubic", align_corners=False, antialias=True).squeeze()
        # patch_mask is accepted and unused, exactly like the reference's render() (it never reads the argument)
        return edict({"rgb": rgb, "alpha": alpha})
