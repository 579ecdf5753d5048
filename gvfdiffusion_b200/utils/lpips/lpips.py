"""LPIPS-VGG16 as the training step uses it (reference utils/lpips/lpips.py:8-34, networks.py:24-97, utils.py:6-8;
called at train_vae.py:93,329 as `LPIPS(net_type='vgg')(pred * 2 - 1, gt * 2 - 1)` and through utils/loss_util.py:66-74
from SparseVAE.training_losses).

SURVEY.md row a17 keeps this term on library kernels: it is 13 VGG16 convolutions (cuDNN) plus five 1x1 heads, nothing of
it is specific to this project.  What is mirrored here is the criterion's arithmetic and its state-dict layout, so that the
joint train step can carry the term:

    feats  = VGG16.features taps after ReLU 1_2, 2_2, 3_3, 4_3, 5_3 (module indices 4, 9, 16, 23, 30) of z-scored input
    f_hat  = f / (||f||_channel + 1e-10)
    d_l    = lin_l((f_hat_x - f_hat_y)^2).mean(H, W)          lin_l: 1x1 conv, C_l -> 1, no bias, frozen
    loss   = sum_l sum_n d_l[n] / N

The reference downloads torchvision's ImageNet VGG16 and the v0.1 linear heads; neither is reachable here, so the module
initialises both randomly unless state dicts are handed in (`load_pretrained`) -- the structure, cost and gradient path are
the ones of the real criterion, the values are not a perceptual metric until real weights are loaded.  Everything is
frozen; gradients flow to `x` only.  Under CUDA the convolutions run in fp16 autocast and channels-last (what accelerate's
mixed precision makes of them in the reference), the target's pass runs without a graph, and everything after the
taps -- normalisation, difference, 1x1 head, spatial mean, and their backward -- is one kernel per tap and direction
(csrc/losses.cu `gvf_lpips_tap_fwd / _bwd`: the torch formulation of that tail was 80 of the criterion's 146 ms per step)."""
import torch
import torch.nn as nn


class _TapFn(torch.autograd.Function):
    """d[n] of one tap from the channels-last fp16 activations of prediction and target; gradient to the prediction."""

    @staticmethod
    def forward(ctx, fx, fy, w):
        from ... import _lib
        from ..._lib import check, current_stream, ptr
        N, C, H, W = fx.shape
        fxc = fx.detach().permute(0, 2, 3, 1).contiguous()          # no copy for channels-last tensors
        fyc = fy.detach().permute(0, 2, 3, 1).contiguous()
        nb = _lib.lib().gvf_lpips_tap_blocks(H * W)
        partial = torch.empty((N, nb), dtype=torch.float32, device=fx.device)
        check(_lib.lib().gvf_lpips_tap_fwd(ptr(fxc), ptr(fyc), ptr(w), N, H * W, C, ptr(partial), current_stream()), "gvf_lpips_tap_fwd")
        ctx.save_for_backward(fxc, fyc, w)
        return partial.sum(1) / float(H * W)

    @staticmethod
    def backward(ctx, gout):
        from ... import _lib
        from ..._lib import check, current_stream, ptr
        fxc, fyc, w = ctx.saved_tensors
        N, H, W, C = fxc.shape
        g = torch.empty_like(fxc)
        check(_lib.lib().gvf_lpips_tap_bwd(ptr(fxc), ptr(fyc), ptr(w), ptr(gout.float().contiguous()), N, H * W, C, ptr(g),
                                           current_stream()), "gvf_lpips_tap_bwd")
        return g.permute(0, 3, 1, 2), None, None

_CFG = (64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M")   # torchvision vgg16 "D"
_TAPS = (4, 9, 16, 23, 30)            # number of feature modules applied before each tap
_CHANNELS = (64, 128, 256, 512, 512)


def _vgg16_features():
    layers, cin = [], 3
    for v in _CFG:
        if v == "M":
            layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
        else:
            layers += [nn.Conv2d(cin, v, kernel_size=3, padding=1), nn.ReLU(inplace=True)]
            cin = v
    return nn.Sequential(*layers)     # module indices equal torchvision.models.vgg16().features'


class _BiasReLUFn(torch.autograd.Function):
    """relu(conv_out + bias) in place on the fresh channels-last convolution output (the bias is frozen)."""

    @staticmethod
    def forward(ctx, y, bias):
        from ... import _lib
        from ..._lib import check, current_stream, ptr
        N, C, H, W = y.shape
        assert y.is_contiguous(memory_format=torch.channels_last) and y.dtype == torch.float16
        check(_lib.lib().gvf_bias_relu_nhwc_f16(ptr(y), ptr(bias), N * H * W, C, current_stream()), "gvf_bias_relu_nhwc_f16")
        ctx.mark_dirty(y)
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        return torch.ops.aten.threshold_backward(g.contiguous(memory_format=torch.channels_last), y, 0), None


class _MaxPool2Fn(torch.autograd.Function):
    """2 x 2 / stride 2 max pool on channels-last fp16 activations."""

    @staticmethod
    def forward(ctx, x):
        from ... import _lib
        from ..._lib import check, current_stream, ptr
        N, C, H, W = x.shape
        y = torch.empty((N, C, H // 2, W // 2), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
        check(_lib.lib().gvf_maxpool2_nhwc_f16(ptr(x), ptr(y), N, H, W, C, current_stream()), "gvf_maxpool2_nhwc_f16")
        ctx.save_for_backward(x, y)
        return y

    @staticmethod
    def backward(ctx, gy):
        from ... import _lib
        from ..._lib import check, current_stream, ptr
        x, y = ctx.saved_tensors
        N, C, H, W = x.shape
        gy = gy.contiguous(memory_format=torch.channels_last)
        gx = torch.empty_like(x)
        check(_lib.lib().gvf_maxpool2_nhwc_bwd_f16(ptr(x), ptr(y), ptr(gy), ptr(gx), N, H, W, C, current_stream()),
              "gvf_maxpool2_nhwc_bwd_f16")
        return gx


class LPIPS(nn.Module):
    def __init__(self, net_type="vgg", version="0.1", vgg_state_dict=None, lin_state_dict=None, seed=0):
        super().__init__()
        if net_type != "vgg" or version != "0.1":
            raise NotImplementedError("the training step uses LPIPS(net_type='vgg'), v0.1")
        g = torch.Generator().manual_seed(seed)
        self.layers = _vgg16_features()
        self.lin = nn.ModuleList([nn.Sequential(nn.Identity(), nn.Conv2d(c, 1, 1, 1, 0, bias=False)) for c in _CHANNELS])
        with torch.no_grad():
            for m in self.layers:
                if isinstance(m, nn.Conv2d):          # He init, as torchvision's un-pretrained VGG
                    m.weight.copy_(torch.randn(m.weight.shape, generator=g) * (2.0 / (m.weight.shape[0] * 9)) ** 0.5)
                    m.bias.zero_()
            for l, c in zip(self.lin, _CHANNELS):     # the published heads are non-negative
                l[1].weight.copy_(torch.rand(l[1].weight.shape, generator=g) / c)
        self.register_buffer("mean", torch.tensor([-.030, -.088, -.188])[None, :, None, None])
        self.register_buffer("std", torch.tensor([.458, .448, .450])[None, :, None, None])
        self.pretrained = False
        self._w16 = {}                                 # fp16 channels-last copies of the (frozen) convolution weights
        if vgg_state_dict is not None or lin_state_dict is not None:
            self.load_pretrained(vgg_state_dict, lin_state_dict)
        for p in self.parameters():
            p.requires_grad_(False)

    def load_pretrained(self, vgg_state_dict=None, lin_state_dict=None):
        """vgg_state_dict: torchvision vgg16 (keys `features.N.weight|bias` or `N.weight|bias`); lin_state_dict: the v0.1
        heads (`lin0.model.1.weight` ... as published, or the reference's renamed `0.1.weight` ...)."""
        if vgg_state_dict is not None:
            sd = {k.replace("features.", ""): v for k, v in vgg_state_dict.items() if not k.startswith("classifier")}
            self.layers.load_state_dict(sd)
            self._w16 = {}
        if lin_state_dict is not None:
            sd = {k.replace("lin", "").replace("model.", ""): v for k, v in lin_state_dict.items()}    # utils.py:22-27
            self.lin.load_state_dict(sd)
        self.pretrained = vgg_state_dict is not None and lin_state_dict is not None

    def taps(self, x):
        """z-score -> VGG16 features; the activations after modules 4, 9, 16, 23, 30 (not yet normalised)."""
        x = (x - self.mean) / self.std
        out = []
        for i, layer in enumerate(self.layers, 1):
            x = layer(x)
            if i in _TAPS:
                out.append(x)
        return out

    def taps_cuda(self, x):
        """`taps` for fp16 channels-last CUDA input: cuDNN convolutions without bias, bias + ReLU and the pools on the
        library's own glue kernels (csrc/losses.cu)."""
        import torch.nn.functional as F
        x = ((x - self.mean) / self.std).to(torch.float16).contiguous(memory_format=torch.channels_last)
        out = []
        for i, layer in enumerate(self.layers, 1):
            if isinstance(layer, nn.Conv2d):
                if x.shape[2] % 2 or x.shape[3] % 2 or layer.out_channels % 8:
                    return None
                w16 = self._w16.get(i)
                if w16 is None or w16.device != x.device:
                    w16 = self._w16[i] = layer.weight.detach().to(torch.float16).contiguous(memory_format=torch.channels_last)
                x = _BiasReLUFn.apply(F.conv2d(x, w16, None, padding=1), layer.bias.detach().float())
            elif isinstance(layer, nn.MaxPool2d):
                x = _MaxPool2Fn.apply(x)
            if i in _TAPS:
                out.append(x)
        return out

    @staticmethod
    def _unit(f):
        return f / (torch.sqrt(torch.sum(f ** 2, dim=1, keepdim=True)) + 1e-10)

    def forward(self, x, y):
        N = x.shape[0]
        if not x.is_cuda:                              # plain torch (CPU tests, the reference's own formulation)
            fx = self.taps(x)
            with torch.no_grad():
                fy = self.taps(y)
            res = [l((self._unit(a) - self._unit(b)) ** 2).mean((2, 3), True) for a, b, l in zip(fx, fy, self.lin)]
            return torch.sum(torch.cat(res, 0)) / N
        fx = self.taps_cuda(x)
        with torch.no_grad():                          # the target's half never enters the graph
            fy = self.taps_cuda(y)
        if fx is None or fy is None:                   # odd image sizes: torch's own pool / bias kernels
            with torch.autocast("cuda", dtype=torch.float16):
                fx = self.taps(x.to(torch.float16).contiguous(memory_format=torch.channels_last))
                with torch.no_grad():
                    fy = self.taps(y.to(torch.float16).contiguous(memory_format=torch.channels_last))
        total = 0.0
        for a, b, l in zip(fx, fy, self.lin):
            total = total + _TapFn.apply(a, b, l[1].weight.reshape(-1).float().contiguous()).sum()
        return total / N
