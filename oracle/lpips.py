"""ORACLE (test infrastructure, never on the product path): functional restatement of the reference's LPIPS-VGG16 criterion
(utils/lpips/lpips.py:29-34, networks.py:45-63,88-97, utils.py:6-8) over a state dict with torchvision's `features.N` layout
and the reference's renamed linear heads (`N.1.weight`).  "Parity unpinned" for the VALUES of a pretrained network (the
ImageNet VGG16 / v0.1 head weights are a network download); the arithmetic is what is restated and tested."""
import torch
import torch.nn.functional as F

CFG = (64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M")
TAPS = (4, 9, 16, 23, 30)


def vgg_taps(sd, x):
    x = (x - torch.tensor([-.030, -.088, -.188])[None, :, None, None]) / torch.tensor([.458, .448, .450])[None, :, None, None]
    out, idx = [], 0
    for v in CFG:
        if v == "M":
            x = F.max_pool2d(x, 2, 2)
            idx += 1
        else:
            x = F.relu(F.conv2d(x, sd[f"{idx}.weight"], sd[f"{idx}.bias"], padding=1))
            idx += 2
        if idx in TAPS:
            out.append(x / (torch.sqrt(torch.sum(x ** 2, dim=1, keepdim=True)) + 1e-10))
    return out


def lpips(vgg_sd, lin_sd, x, y):
    fx, fy = vgg_taps(vgg_sd, x), vgg_taps(vgg_sd, y)
    res = [F.conv2d((a - b) ** 2, lin_sd[f"{i}.1.weight"]).mean((2, 3), True) for i, (a, b) in enumerate(zip(fx, fy))]
    return torch.sum(torch.cat(res, 0)) / x.shape[0]
