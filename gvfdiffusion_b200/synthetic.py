"""Seeded synthetic inputs of the benchmark shapes.

There is no network for checkpoints or datasets, so the benchmark and the parity tests run
on random-init weights of the reference architecture and on synthetic canonical Gaussians
built exactly like `SparseVAE.to_representation` builds them
(reference model/sparse_voxel_diffusion/sparse_vae.py:114-182, config configs/diffusion.yml
`MipGS`: 8 Gaussians per voxel of a 64^3 grid, soft_invoxel offsets, voxel_size 1.5,
lr _rotation 0.1, Hammersley perturbation).  Everything is generated on the CPU with a
seeded torch.Generator and moved to the device by the caller.
"""
import math

import numpy as np
import torch

PRIMES = [2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53]


def _radical_inverse(base, n):
    val, inv_base = 0.0, 1.0 / base
    inv_base_n = inv_base
    while n > 0:
        val += (n % base) * inv_base_n
        n //= base
        inv_base_n *= inv_base
    return val


def _hammersley(dim, n, num):
    return [n / num] + [_radical_inverse(PRIMES[d], n) for d in range(dim - 1)]


def gaussian_constants():
    """Biases of GaussianModel for the MipGS config (gaussian_model.py:23-41)."""
    x = torch.tensor(0.004)
    p = torch.tensor(0.1)
    return {"aabb": (-0.5, -0.5, -0.5, 1.0, 1.0, 1.0),
            "scale_bias": float(x + torch.log(-torch.expm1(-x))),
            "min_kernel": 0.0009,
            "opacity_bias": float(torch.log(p / (1 - p))),
            "softplus": True}


def canonical_gaussians(num_voxels=2048, resolution=64, num_gaussians=8, seed=0,
                        shell_radius=0.35, voxel_size=1.5):
    """-> dict of raw GaussianModel tensors (P = num_voxels * num_gaussians)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    # distinct voxels on a spherical shell
    coords = set()
    while len(coords) < num_voxels:
        d = torch.randn(4 * num_voxels, 3, generator=g)
        d = d / d.norm(dim=1, keepdim=True) * shell_radius
        d = d + 0.02 * torch.randn(d.shape, generator=g)
        c = torch.clamp(((d + 0.5) * resolution).floor().long(), 0, resolution - 1)
        for row in c.tolist():
            coords.add(tuple(row))
            if len(coords) == num_voxels:
                break
    coords = torch.tensor(sorted(coords), dtype=torch.float32)
    xyz_center = (coords + 0.5) / resolution
    off = torch.tensor([_hammersley(3, i, num_gaussians) for i in range(num_gaussians)]).float() - 0.5
    perturb = torch.atanh(off / 0.5 / voxel_size)
    feats = lambda *s: torch.randn(*s, generator=g)
    offset = feats(num_voxels, num_gaussians, 3) * 1.0 + perturb
    offset = torch.tanh(offset) / resolution * 0.5 * voxel_size
    P = num_voxels * num_gaussians
    return {
        "_xyz": (xyz_center.unsqueeze(1) + offset).reshape(P, 3).contiguous(),
        "_features_dc": feats(P, 1, 3).contiguous(),
        "_scaling": feats(P, 3).contiguous(),
        "_rotation": (feats(P, 4) * 0.1).contiguous(),
        "_opacity": feats(P, 1).contiguous(),
    }


def raster_delta(F, P, seed=1):
    """Per-frame delta ~ N(0, sigma) with sigma = {xyz .01, scale .05, rot .05, rgb .05, opacity .1}."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    sig = torch.tensor([0.01] * 3 + [0.05] * 3 + [0.05] * 4 + [0.05] * 3 + [0.1])
    return torch.randn(F, P, 14, generator=g) * sig


def orbit_camera(elevation, azimuth, radius=1.0):
    """kiui.cam.orbit_camera(opengl=True) restated: camera-to-world, y up, looking at the origin."""
    el, az = np.deg2rad(elevation), np.deg2rad(azimuth)
    campos = np.array([radius * np.cos(el) * np.sin(az), -radius * np.sin(el),
                       radius * np.cos(el) * np.cos(az)], dtype=np.float32)
    nrm = lambda v: v / (np.linalg.norm(v) + 1e-20)
    fwd = nrm(campos)
    right = nrm(np.cross(np.array([0, 1, 0], dtype=np.float32), fwd))
    up = nrm(np.cross(fwd, right))
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = np.stack([right, up, fwd], axis=1)
    T[:3, 3] = campos
    return T


def orbit_extrinsics(F, elevation=0.0, radius=2.0, azimuths=None):
    """World->camera matrices of the reference render loop (utils/inference_utils.py:245-254),
    azimuth 360 * f / F (or the given list of azimuths in degrees: the alignment loop of :53-62)."""
    convert = np.array([[1, 0, 0, 0], [0, 0, -1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.float32)
    out = []
    for f in range(F):
        pose = convert @ orbit_camera(elevation, 360.0 * f / F if azimuths is None else float(azimuths[f]), radius)
        pose[:3, 1:3] *= -1
        out.append(np.linalg.inv(pose))
    return torch.from_numpy(np.stack(out)).float()


def intrinsics(fov_deg=49.1):
    f = 0.5 / math.tan(math.radians(fov_deg) / 2)
    return torch.tensor([[f, 0, 0.5], [0, f, 0.5], [0, 0, 1]], dtype=torch.float32)


def sampler_inputs(B=1, T=24, N=512, C=16, L_img=1370, C_img=1024, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return {"noise": torch.randn(B, T, N, C, generator=g),
            "cond_images": torch.randn(B, T, L_img, C_img, generator=g)}
