"""`utils.loss_util` of the reference (utils/loss_util.py:17-63) on the sm_100a loss kernels
(csrc/losses.cu): `l1_loss`, `l2_loss`, `ssim`, plus `ssim_l1`, the fused form of the two pixel losses of the
training step (train_vae.py:328-330).  Differentiable with respect to the first image (the prediction), like
the way the reference uses them; CUDA fp32 tensors only -- there is no CPU fallback.
"""
import torch

from .. import _lib


def _planes(img):
    if not (img.is_cuda and img.dtype == torch.float32):
        raise ValueError(f"expected a CUDA float32 image tensor, got {img.dtype} on {img.device}")
    if img.ndim < 3:
        raise ValueError("expected [..., C, H, W]")
    H, W = img.shape[-2:]
    n = img.numel() // (H * W)
    return n, H, W


class _SsimL1(torch.autograd.Function):
    """sums [planes, 2] = (sum of the SSIM map, sum of |img1 - img2|) per (batch, channel) plane."""

    @staticmethod
    def forward(ctx, img1, img2):
        n, H, W = _planes(img1)
        if img2.shape != img1.shape:
            raise ValueError("img1 and img2 must have the same shape")
        _planes(img2)
        a, b = img1.contiguous(), img2.contiguous()
        L = _lib.lib()
        nbytes = L.gvf_ssim_l1_workspace_bytes(n, H, W)
        ws = torch.empty(nbytes // 4, dtype=torch.float32, device=a.device)
        sums = torch.empty((n, 2), dtype=torch.float32, device=a.device)
        need = ctx.needs_input_grad[0]
        dmaps = torch.empty((3, n, H, W), dtype=torch.float32, device=a.device) if need else None
        _lib.check(L.gvf_ssim_l1_fwd(_lib.ptr(a), _lib.ptr(b), n, H, W, _lib.ptr(ws), nbytes, _lib.ptr(sums),
                                     _lib.ptr(dmaps), _lib.current_stream()), "gvf_ssim_l1_fwd")
        if need:
            ctx.save_for_backward(a, b, dmaps)
        ctx.mark_non_differentiable()
        return sums

    @staticmethod
    def backward(ctx, g):
        a, b, dmaps = ctx.saved_tensors
        n, H, W = _planes(a)
        g = g.to(torch.float32).contiguous()
        cs, cl = g[:, 0].contiguous(), g[:, 1].contiguous()
        grad = torch.empty_like(a)
        _lib.check(_lib.lib().gvf_ssim_l1_bwd(_lib.ptr(a), _lib.ptr(b), _lib.ptr(dmaps), n, H, W, _lib.ptr(cs),
                                              _lib.ptr(cl), _lib.ptr(grad), _lib.current_stream()), "gvf_ssim_l1_bwd")
        return grad, None


def ssim_l1(img1, img2, size_average=True):
    """(ssim, l1) of utils/loss_util.py -- one forward and one backward kernel for both."""
    n, H, W = _planes(img1)
    sums = _SsimL1.apply(img1, img2)
    if size_average:
        tot = sums.sum(0) / float(n * H * W)
        return tot[0], tot[1]
    B = img1.shape[0]
    per = sums.view(B, -1, 2).sum(1) / float((n // B) * H * W)
    return per[:, 0], per[:, 1]


def ssim(img1, img2, window_size=11, size_average=True):
    """utils/loss_util.py:33-63.  size_average=False returns the per-batch-element mean."""
    if window_size != 11:
        raise NotImplementedError("the kernel is specialised for the reference's 11x11 window")
    return ssim_l1(img1, img2, size_average)[0]


def l1_loss(network_output, gt):
    """utils/loss_util.py:17-18"""
    return ssim_l1(network_output, gt, True)[1]


def l2_loss(network_output, gt):
    """utils/loss_util.py:20-21 (plain torch: not used by the training step)"""
    return ((network_output - gt) ** 2).mean()


_loss_fn_vgg = {}


def lpips(img1, img2, value_range=(0, 1), module=None):
    """utils/loss_util.py:66-74: LPIPS-VGG16 of images in `value_range`, mapped to [-1, 1].  The reference builds its
    (downloaded) criterion lazily and keeps it in a global; here the lazily built one has seeded random weights unless
    `module` (an LPIPS with real weights, gvfdiffusion_b200.utils.lpips.LPIPS.load_pretrained) is passed."""
    from .lpips import LPIPS
    if module is None:
        key = str(img1.device)
        if key not in _loss_fn_vgg:
            _loss_fn_vgg[key] = LPIPS(net_type="vgg").to(img1.device).eval()
        module = _loss_fn_vgg[key]
    a = (img1 - value_range[0]) / (value_range[1] - value_range[0]) * 2 - 1
    b = (img2 - value_range[0]) / (value_range[1] - value_range[0]) * 2 - 1
    return module(a, b).mean()
