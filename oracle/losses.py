"""CPU oracle (TEST INFRASTRUCTURE ONLY) of the training-step losses, SURVEY.md row a17:

    ssim / l1            utils/loss_util.py:17-63 (2-D 11x11 Gaussian window sigma 1.5 applied as a depthwise
                         zero-padded convolution to img1, img2, img1^2, img2^2, img1 img2), train_vae.py:328-330
    knn_points           pytorch3d.ops.knn_points as called at train_vae.py:525-530 -- third-party, absent here
                         (pinned version: whatever `pip install pytorch3d` resolved in setup.sh; semantics from its
                         docstring: squared distances, ascending, zero padding) -> numpy brute force, every
                         operation rounded to fp32 ((dx*dx + dy*dy) + dz*dz), stable sort: ties -> lowest index
    interp_deltas / interpolation_loss   train_vae.py:486-586

Pinned by tests/test_oracle_golden.py against tests/golden/losses.pt, which the reference's own ssim and
compute_interpolation_loss_delta_interp produced (tests/golden/make_golden.py::gen_losses).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def _window(channel, window_size=11, sigma=1.5):
    g = torch.Tensor([math.exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)])
    g = (g / g.sum()).unsqueeze(1)
    w2 = g.mm(g.t()).float().unsqueeze(0).unsqueeze(0)
    return w2.expand(channel, 1, window_size, window_size).contiguous()


def ssim_map(img1, img2, window_size=11):
    C = img1.size(-3)
    w = _window(C, window_size).type_as(img1)
    conv = lambda x: F.conv2d(x, w, padding=window_size // 2, groups=C)
    mu1, mu2 = conv(img1), conv(img2)
    mu1_sq, mu2_sq, mu12 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1 = conv(img1 * img1) - mu1_sq
    s2 = conv(img2 * img2) - mu2_sq
    s12 = conv(img1 * img2) - mu12
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    return ((2 * mu12 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))


def ssim(img1, img2, size_average=True):
    m = ssim_map(img1, img2)
    return m.mean() if size_average else m.mean(1).mean(1).mean(1)


def l1_loss(a, b):
    return torch.abs(a - b).mean()


def knn_points(p1, p2, lengths1=None, lengths2=None, K=1):
    """numpy brute force; returns (dists [B,P1,K] fp32, idx [B,P1,K] int64)."""
    a, b = np.asarray(p1, np.float32), np.asarray(p2, np.float32)
    B, P1, P2 = a.shape[0], a.shape[1], b.shape[1]
    dists = np.zeros((B, P1, K), np.float32)
    idx = np.zeros((B, P1, K), np.int64)
    for n in range(B):
        n1 = P1 if lengths1 is None else int(lengths1[n])
        n2 = P2 if lengths2 is None else int(lengths2[n])
        d = a[n, :n1, None, :] - b[n, None, :n2, :]
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
        order = np.argsort(d2, axis=1, kind="stable")[:, :K]
        kk = order.shape[1]
        dists[n, :n1, :kk] = np.take_along_axis(d2, order, 1)
        idx[n, :n1, :kk] = order
    return dists, idx


def interp_deltas(knn_dists, knn_idx, micro_static_pc, micro_moving_pc, lengths1, adaptive_radius=True, beta=7.0):
    """train_vae.py:532-563 -> estimated deltas [B,T,P1,3]."""
    d = torch.as_tensor(knn_dists)
    idx = torch.as_tensor(knn_idx)
    B, P1, K = d.shape
    T = micro_moving_pc.shape[1]
    radii = d.mean(dim=-1).sqrt() + 1e-6
    if adaptive_radius:
        w = torch.exp(-beta * d / radii[..., None] ** 2) * (d <= radii[..., None] ** 2).float()
    else:
        w = torch.exp(-beta * d)
    pad = torch.arange(P1).expand(B, -1) < torch.as_tensor(lengths1).unsqueeze(-1)
    w = w * pad.unsqueeze(-1)
    w = w / (w.sum(dim=-1, keepdim=True) + 1e-8)
    bi = torch.arange(B).view(-1, 1, 1)
    nb_static = micro_static_pc[bi, idx]                               # [B,P1,K,3]
    est = torch.zeros(B, T, P1, 3)
    for t in range(T):
        mv = micro_moving_pc[:, t][bi, idx] - nb_static
        est[:, t] = (w.unsqueeze(-1) * mv).sum(dim=2)
    return est


def interpolation_loss(static_gs, micro_static_pc, micro_moving_pc, output, knn_k=4, adaptive_radius=True, beta=7.0):
    B = len(static_gs)
    sizes = [s.shape[0] for s in static_gs]
    mx = max(sizes)
    padded = torch.stack([F.pad(s[:, :3], (0, 0, 0, mx - s.shape[0])) for s in static_gs])
    kd, ki = knn_points(padded.numpy(), micro_static_pc.numpy(), lengths1=sizes, K=knn_k)
    est = interp_deltas(kd, ki, micro_static_pc, micro_moving_pc, sizes, adaptive_radius, beta)
    T = micro_moving_pc.shape[1]
    mask = (torch.arange(mx).expand(B, -1) < torch.tensor(sizes).unsqueeze(1)).unsqueeze(1).expand(-1, T, -1)
    pred = torch.stack([F.pad(output[b, :, :sizes[b], :3], (0, 0, 0, mx - sizes[b])) for b in range(B)])
    loss = (torch.abs(pred - est) * mask.unsqueeze(-1)).sum() / (mask.sum() * 3)
    return loss, est, kd, ki
