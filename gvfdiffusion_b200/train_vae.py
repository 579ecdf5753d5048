"""Loss-side helpers of the reference's `train_vae.py` that sit next to the render path: `get_gaussian_tensor`
(:466-472), `pad_static_gs` (:475-483), `compute_interpolation_loss_delta_interp` (:486-586) and the
`pytorch3d.ops.knn_points` call inside it (:525-530), on the sm_100a kernels of csrc/losses.cu.  The training
loop itself (optimizer, EMA, logging) is out of scope.
"""
import torch
import torch.nn.functional as F

from . import _lib
from .pipeline import pad_static_gs  # noqa: F401  (same name as the reference's)


def get_gaussian_tensor(gaussians):
    """[xyz | features_dc | opacity | scaling | rotation] -> [P, 14] of one GaussianModel (train_vae.py:466-472);
    one `gvf_gaussian_tensor` launch instead of five activations and a cat."""
    if hasattr(gaussians, "gaussian_tensor"):
        return gaussians.gaussian_tensor()
    return torch.cat([gaussians.get_xyz, gaussians.get_features.squeeze(), gaussians.get_opacity,
                      gaussians.get_scaling, gaussians.get_rotation], dim=-1)


@torch.no_grad()
def knn_points(p1, p2, lengths1=None, lengths2=None, K=1):
    """pytorch3d.ops.knn_points(p1 [B,P1,3], p2 [B,P2,3], lengths1, lengths2, K) -> (dists [B,P1,K] squared,
    ascending; idx [B,P1,K] int64; None).  Exact search, ties -> lowest index."""
    if not (p1.is_cuda and p2.is_cuda):
        raise ValueError("knn_points: CUDA tensors only (no CPU fallback)")
    if p1.shape[-1] != 3 or p2.shape[-1] != 3 or p1.shape[0] != p2.shape[0]:
        raise ValueError("knn_points: expected [B,P1,3] and [B,P2,3]")
    a, b = p1.to(torch.float32).contiguous(), p2.to(torch.float32).contiguous()
    B, P1, P2 = a.shape[0], a.shape[1], b.shape[1]
    l1 = None if lengths1 is None else lengths1.to(device=a.device, dtype=torch.int64).contiguous()
    l2 = None if lengths2 is None else lengths2.to(device=a.device, dtype=torch.int64).contiguous()
    d = torch.empty((B, P1, K), dtype=torch.float32, device=a.device)
    i = torch.empty((B, P1, K), dtype=torch.int64, device=a.device)
    _lib.check(_lib.lib().gvf_knn(_lib.ptr(a), _lib.ptr(b), B, P1, P2, _lib.ptr(l1), _lib.ptr(l2), K, _lib.ptr(d),
                                  _lib.ptr(i), _lib.current_stream()), "gvf_knn")
    return d, i, None


@torch.no_grad()
def interpolate_deltas(knn_dists, knn_idx, micro_static_pc, micro_moving_pc, lengths1, adaptive_radius=True, beta=7.0):
    """estimated_deltas [B,T,P1,3] of train_vae.py:532-563 from the K nearest neighbours."""
    B, P1, K = knn_dists.shape
    T, P2 = micro_moving_pc.shape[1], micro_static_pc.shape[1]
    s = micro_static_pc.to(torch.float32).contiguous()
    m = micro_moving_pc.to(torch.float32).contiguous()
    l1 = None if lengths1 is None else lengths1.to(device=s.device, dtype=torch.int64).contiguous()
    est = torch.empty((B, T, P1, 3), dtype=torch.float32, device=s.device)
    _lib.check(_lib.lib().gvf_knn_interp_deltas(_lib.ptr(knn_dists), _lib.ptr(knn_idx), _lib.ptr(s), _lib.ptr(m),
                                                _lib.ptr(l1), B, P1, P2, T, K, int(bool(adaptive_radius)), float(beta),
                                                _lib.ptr(est), _lib.current_stream()), "gvf_knn_interp_deltas")
    return est


def compute_interpolation_loss_delta_interp(static_gs, micro_static_pc, micro_moving_pc, output, B, knn_k=4,
                                            adaptive_radius=True, beta=7.0):
    """train_vae.py:486-586: masked L1 between the predicted xyz deltas `output[b, :, :P_b, :3]` and the
    RBF-weighted motion of the `knn_k` nearest tracked points.  Returns (loss, {"deformation_xyz_loss"}, estimated
    deltas [B,T,max P,3]) like the reference; differentiable with respect to `output`."""
    dev = micro_static_pc.device
    xyz = [static_gs[b][:, :3] for b in range(B)]
    T = micro_moving_pc.shape[1]
    max_samples = max(x.shape[0] for x in xyz)
    with torch.no_grad():
        padded = torch.stack([F.pad(x, (0, 0, 0, max_samples - x.shape[0])) for x in xyz])
        lengths = torch.tensor([x.shape[0] for x in xyz], dtype=torch.int64, device=dev)
        knn_dists, knn_idx, _ = knn_points(padded, micro_static_pc, lengths1=lengths, K=knn_k)
        estimated = interpolate_deltas(knn_dists, knn_idx, micro_static_pc, micro_moving_pc, lengths, adaptive_radius,
                                       beta)
        mask = (torch.arange(max_samples, device=dev).expand(B, -1) < lengths.unsqueeze(1)).unsqueeze(1).expand(-1, T, -1)
    pred = torch.stack([F.pad(output[b, :, :xyz[b].shape[0], :3], (0, 0, 0, max_samples - xyz[b].shape[0]))
                        for b in range(B)])
    diff = torch.abs(pred - estimated) * mask.unsqueeze(-1)
    loss = diff.sum() / (mask.sum() * 3)
    return loss, {"deformation_xyz_loss": loss.detach().reshape(1)}, estimated
