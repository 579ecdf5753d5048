"""Canonical-pose alignment of the reference's in-the-wild inference (utils/inference_utils.py:37-177): the canonical
Gaussians are rendered from 360 azimuths (every 90 degrees for dataset objects) with the plain-3DGS rasteriser, every
render is rescaled so that its alpha bounding box matches the conditioning image's, compared with that image (L1 + 0.2 *
CLIP distance), and the object is rotated about z so that the best azimuth becomes the front view.

Here the 360 renders are ONE frame-batched rasteriser call (gvf_raster_forward, mip_filter = 0) instead of 360 Python
`render` calls; bounding boxes, scale factors and L1 distances are evaluated for all views on the device.  The bicubic
resize / pad / crop of each view is the same torch library call as the reference's.  CLIP is third-party model code and
stays outside: pass `clip_distance(render [3,512,512], canonical [3,512,512]) -> float` to include its term, otherwise the
L1 term alone decides (the reference's weighting is 1 : 0.2).  The quaternion update composes the rotation in quaternion
form (the reference goes through rotation matrices and pytorch3d's matrix_to_quaternion; same rotation, sign included
for the usual w > 0 case).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from .. import raster as R, synthetic as S


def orbit_extrinsics(azimuths_deg, elevation=0.0, radius=2.0):
    """World-to-camera matrices of the reference's loop (:56-62): kiui orbit_camera(elevation, azimuth, radius, opengl=True),
    pre-multiplied by the y/z swap, columns 1:3 negated, inverted."""
    return S.orbit_extrinsics(len(azimuths_deg), azimuths=list(azimuths_deg), elevation=elevation, radius=radius)


def _bbox_size(mask):
    """mask [F,H,W] bool -> max(height, width) extent of the true region per frame (int64), -1 where empty."""
    Fn, H, W = mask.shape
    rows, cols = mask.any(2), mask.any(1)
    ar_h, ar_w = torch.arange(H, device=mask.device), torch.arange(W, device=mask.device)
    big = 1 << 30
    y0 = torch.where(rows, ar_h, big).amin(1)
    y1 = torch.where(rows, ar_h, -1).amax(1)
    x0 = torch.where(cols, ar_w, big).amin(1)
    x1 = torch.where(cols, ar_w, -1).amax(1)
    size = torch.maximum(y1 - y0, x1 - x0)
    return torch.where(rows.any(1), size, torch.full_like(size, -1))


def _fit_512(image, target_size, out=512):
    """bicubic resize to target_size^2, then centre pad (white) or crop to out^2 -- reference :87-103."""
    image = F.interpolate(image.unsqueeze(0), size=(target_size, target_size), mode="bicubic", align_corners=False).squeeze(0)
    _, H, W = image.shape
    if H < out or W < out:
        ph, pw = max(0, (out - H) // 2), max(0, (out - W) // 2)
        image = F.pad(image, (pw, pw + (out - W - 2 * pw), ph, ph + (out - H - 2 * ph)), mode="constant", value=1.0)
    else:
        top, left = (H - out) // 2, (W - out) // 2
        image = image[:, top:top + out, left:left + out]
    return image.clamp(0.0, 1.0)


@torch.no_grad()
def find_best_azimuth(gaussian, renderer, canonical_image, canonical_alpha, intrinsics, in_the_wild=True, clip_distance=None,
                      chunk=90):
    """-> (best_azimuth_deg, best_scale_factor, per-view table [azimuth, l1, clip, total, scale]).  `renderer` is a
    gvfdiffusion_b200.renderers.GaussianRenderer whose rendering_options are set (resolution 512, near / far, bg)."""
    dev = gaussian._xyz.device
    azis = np.arange(-180, 180, 1) if in_the_wild else np.arange(-180, 180, 90)
    mip_was = renderer.pipe.use_mip_gaussian
    renderer.pipe.use_mip_gaussian = False                     # reference :50
    can_mask = canonical_alpha.to(dev) > 0.5
    can_size = int(_bbox_size(can_mask[None])[0])
    canonical_image = canonical_image.to(dev)
    rows = []
    try:
        for s in range(0, len(azis), chunk):
            az = azis[s:s + chunk]
            ext = orbit_extrinsics(az).to(dev)
            rgba, _ = renderer.render_frames(gaussian, ext, intrinsics.to(dev))
            sizes = _bbox_size(rgba[:, 3] > 0.5).tolist()
            for k, a in enumerate(az):
                if sizes[k] < 0 or can_size < 0:
                    continue                                   # empty render / empty canonical mask (:72-84)
                scale = can_size / sizes[k] if sizes[k] > 0 else float("inf")
                if not math.isfinite(scale):
                    continue
                img = _fit_512(rgba[k, :3].clamp(0.0, 1.0), int(512 * scale))
                l1 = float((img - canonical_image).abs().mean())
                cd = float(clip_distance(img, canonical_image)) if clip_distance is not None else 0.0
                rows.append((int(a), l1, cd, l1 + 0.2 * cd, scale))
    finally:
        renderer.pipe.use_mip_gaussian = mip_was
    if not rows:
        return 0, 1.0, rows
    best = min(rows, key=lambda r: r[3])                       # first minimum, like the reference's strict `<`
    return best[0], best[4], rows


def _quat_mul(a, b):
    aw, ax, ay, az = a.unbind(-1)
    bw, bx, by, bz = b.unbind(-1)
    return torch.stack([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                        aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw], -1)


@torch.no_grad()
def align_gaussian_to_canonical(static_gs_model, canonical_image, canonical_alpha, intrinsics, renderer, id=0, device="cuda",
                                in_the_wild=True, clip_distance=None):
    """Reference signature (:37) with the renderer passed directly (the reference reaches it through
    static_vae.renderers["MipGS"]).  Rotates the model in place by -best_azimuth about z; returns (model, scale)."""
    best_azi, best_scale, _ = find_best_azimuth(static_gs_model, renderer, canonical_image, canonical_alpha, intrinsics,
                                                in_the_wild, clip_distance)
    ang = math.radians(-best_azi)
    c, s = math.cos(ang), math.sin(ang)
    rot = torch.tensor([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]], dtype=torch.float32, device=static_gs_model._xyz.device)
    xyz = static_gs_model.get_xyz
    static_gs_model.from_xyz((rot @ xyz.T).T)
    qz = torch.tensor([math.cos(ang / 2), 0.0, 0.0, math.sin(ang / 2)], dtype=torch.float32, device=rot.device)
    static_gs_model.from_rotation(_quat_mul(qz.expand_as(static_gs_model.get_rotation), static_gs_model.get_rotation))
    return static_gs_model, best_scale


# ---------------------------------------------------------------------------------------------- output stage (:276-297)
def pil_resample_coeffs(in_size, out_size, support=3.0):
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc for the LANCZOS filter over the whole input range:
    -> (bounds int32 [out, 2], coeffs int32 [out, ksize], ksize).  Double arithmetic with libm's sin, like Pillow."""
    def lanczos(x):
        if -3.0 <= x < 3.0:
            if x == 0.0:
                return 1.0
            a, b = x * math.pi, x * math.pi / 3.0
            return (math.sin(a) / a) * (math.sin(b) / b)
        return 0.0
    scale = in_size / out_size
    fscale = max(scale, 1.0)
    sup = support * fscale
    ksize = int(math.ceil(sup)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / fscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - sup + 0.5), 0)
        xmax = min(int(center + sup + 0.5), in_size) - xmin
        w = [lanczos((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = sum(w)                                       # Pillow accumulates left to right in a double, as sum() does
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << 22)) if v < 0 else int(0.5 + v * (1 << 22))
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


_coef_cache = {}


def resize_pad_frames_u8(frames, scale_factor, size=512, fill=255):
    """frames uint8 [F, H, W, 3] on the device (gvf_rgba_to_u8 output) -> uint8 [F, size, size, 3]: PIL LANCZOS resize to
    int(size * scale_factor) and centre pad (white) / crop, exactly the reference's per-frame host code (:284-297)."""
    from .. import _lib
    from .._lib import check, current_stream, ptr
    if not (frames.is_cuda and frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[-1] == 3):
        raise ValueError("frames: expected a CUDA uint8 [F, H, W, 3] tensor")
    frames = frames.contiguous()
    Fn, H, W, _ = frames.shape
    t = int(size * scale_factor)
    if t <= 0:
        raise ValueError("scale factor too small")
    dev = frames.device
    key = (H, W, t, str(dev))
    if key not in _coef_cache:
        bh, kh, ksh = pil_resample_coeffs(W, t)
        bv, kv, ksv = pil_resample_coeffs(H, t)
        _coef_cache[key] = tuple(torch.from_numpy(a).to(dev) for a in (bh, kh, bv, kv)) + (ksh, ksv)
    bh, kh, bv, kv, ksh, ksv = _coef_cache[key]
    if t == W and t == H:
        resized = frames                                  # Pillow returns a copy without resampling
    else:
        tmp = torch.empty((Fn, H, t, 3), dtype=torch.uint8, device=dev)
        resized = torch.empty((Fn, t, t, 3), dtype=torch.uint8, device=dev)
        check(_lib.lib().gvf_resample_u8(ptr(frames), Fn, H, W, t, t, ptr(bh), ptr(kh), ksh, ptr(bv), ptr(kv), ksv, ptr(tmp),
                                         ptr(resized), current_stream()), "gvf_resample_u8")
    out = torch.empty((Fn, size, size, 3), dtype=torch.uint8, device=dev)
    check(_lib.lib().gvf_pad_crop_u8(ptr(resized), Fn, t, t, size, fill, ptr(out), current_stream()), "gvf_pad_crop_u8")
    return out
