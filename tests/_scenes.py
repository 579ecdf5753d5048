"""Shared scene builders for the rasteriser tests (CPU tensors)."""
import numpy as np
import torch

from gvfdiffusion_b200 import synthetic as S
from oracle import gaussian as G
from oracle import raster as OR


def scene(num_voxels=256, F=3, H=128, W=128, seed=0, with_delta=True, scale_boost=0.0):
    canon = S.canonical_gaussians(num_voxels=num_voxels, seed=seed)
    if scale_boost:
        canon["_scaling"] = canon["_scaling"] + scale_boost
    P = canon["_xyz"].shape[0]
    delta = S.raster_delta(F, P, seed=seed + 1) if with_delta else None
    ext = S.orbit_extrinsics(F)
    intr = S.intrinsics()
    const = S.gaussian_constants()
    return canon, delta, ext, intr, const


def oracle_frames(canon, delta, ext, intr, const, H, W, near=0.8, far=1.6, bg=(1.0, 1.0, 1.0),
                  kernel_size=0.1):
    """Per-frame oracle outputs (list of dicts) + params."""
    outs = []
    cn = {k: v.numpy() for k, v in canon.items()}
    for f in range(ext.shape[0]):
        vt, pt, _, tfx, tfy = G.camera_matrices(ext[f], intr, near, far)
        prm = OR.make_params(H, W, tfx, tfy, const, kernel_size=kernel_size, bg=bg)
        m3, sc, rt, sh, op = OR.activate(prm, cn, None if delta is None else delta[f].numpy())
        o = OR.forward(prm, m3, sc, rt, sh, op, vt.numpy(), pt.numpy())
        o["activated"] = (m3, sc, rt, sh, op)
        outs.append(o)
    return outs
