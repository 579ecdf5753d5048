"""Output stage of the visualisation loop (reference utils/inference_utils.py:276-297): uint8 frames ->
PIL LANCZOS resize to int(512 * scale) -> centre pad (white) / crop to 512^2.  The oracle is Pillow itself (the library
the reference calls), byte for byte: on the CPU for the host-side coefficient tables (numpy emulation of the two integer
passes), on the GPU for the kernels."""
import numpy as np
import pytest
import torch
from PIL import Image

from gvfdiffusion_b200.utils.inference_utils import pil_resample_coeffs


def _pil_stage(frame, scale, size=512):
    t = int(size * scale)
    image = Image.fromarray(frame).resize((t, t), resample=Image.Resampling.LANCZOS)
    W, H = image.size
    if H < size or W < size:
        pad_h, pad_w = max(0, (size - H) // 2), max(0, (size - W) // 2)
        new = Image.new("RGB", (size, size), (255, 255, 255))
        new.paste(image, (pad_w, pad_h))
        image = new
    else:
        left, top = (W - size) // 2, (H - size) // 2
        image = image.crop((left, top, left + size, top + size))
    return np.asarray(image)


def test_coefficient_tables_reproduce_pillow_on_cpu():
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (96, 96, 3), dtype=np.uint8)
    for t in (40, 95, 96, 130, 33, 200):
        ref = np.asarray(Image.fromarray(img).resize((t, t), resample=Image.Resampling.LANCZOS))
        b, k, ks = pil_resample_coeffs(96, t)
        assert k.shape == (t, ks) and (b[:, 0] + b[:, 1] <= 96).all()
        tmp = np.zeros((96, t, 3), np.uint8)
        for xo in range(t):
            x0, n = b[xo]
            s = (1 << 21) + (img[:, x0:x0 + n, :].astype(np.int64) * k[xo, :n][None, :, None]).sum(1)
            tmp[:, xo, :] = np.clip(s >> 22, 0, 255)
        out = np.zeros((t, t, 3), np.uint8)
        for yo in range(t):
            y0, n = b[yo]
            s = (1 << 21) + (tmp[y0:y0 + n].astype(np.int64) * k[yo, :n][:, None, None]).sum(0)
            out[yo] = np.clip(s >> 22, 0, 255)
        assert np.array_equal(out, ref), t


@pytest.mark.gpu
@pytest.mark.parametrize("scale", [0.83, 1.0, 1.37, 0.25, 2.0])
def test_resize_pad_frames_equal_pillow(scale):
    from gvfdiffusion_b200.utils.inference_utils import resize_pad_frames_u8
    rng = np.random.default_rng(int(scale * 100))
    frames = rng.integers(0, 256, (3, 512, 512, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:512, 0:512]
    frames[1] = np.stack([(xx // 2) % 256, (yy // 2) % 256, ((xx + yy) // 4) % 256], -1).astype(np.uint8)   # smooth image
    got = resize_pad_frames_u8(torch.from_numpy(frames).cuda(), scale).cpu().numpy()
    for f in range(3):
        assert np.array_equal(got[f], _pil_stage(frames[f], scale)), (scale, f)


@pytest.mark.gpu
def test_render_views_to_output_frames():
    """rgba -> uint8 (gvf_rgba_to_u8) -> resize / pad: the whole tail of render_and_save_images for a few frames."""
    from gvfdiffusion_b200 import _lib
    from gvfdiffusion_b200._lib import check, current_stream, ptr
    from gvfdiffusion_b200.utils.inference_utils import resize_pad_frames_u8
    g = torch.Generator().manual_seed(1)
    rgba = (torch.rand(2, 4, 512, 512, generator=g) * 1.2 - 0.1).cuda()
    u8 = torch.empty((2, 512, 512, 3), dtype=torch.uint8, device="cuda")
    check(_lib.lib().gvf_rgba_to_u8(ptr(rgba), 2, 512, 512, ptr(u8), current_stream()), "gvf_rgba_to_u8")
    out = resize_pad_frames_u8(u8, 0.9).cpu().numpy()
    for f in range(2):
        rgb = (rgba[f, :3].clamp(0.0, 1.0).permute(1, 2, 0).cpu().numpy() * 255).astype("uint8")     # reference :278-283
        assert np.array_equal(out[f], _pil_stage(rgb, 0.9))
