"""CPU oracle: GaussianModel activations and GaussianRenderer camera set-up
(TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates in torch-CPU fp32:
  * representations/gaussian/gaussian_model.py:84-114  (get_xyz / get_*_with_delta)
  * representations/gaussian/gaussian_model.py:23-41   (scale / rotation / opacity biases)
  * renderers/gaussian_render.py:57-82                 (intrinsics_to_projection)
  * renderers/gaussian_render.py:302-321               (view / full-projection / campos / tanfov)
  * utils/inference_utils.py:240-254                   (orbit cameras of the render loop)
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

DELTA_SLICES = {"xyz": (0, 3), "scale": (3, 6), "rot": (6, 10), "rgb": (10, 13), "opacity": (13, 14)}


def model_constants(scaling_bias=0.004, opacity_bias=0.1, min_kernel=0.0009,
                    aabb=(-0.5, -0.5, -0.5, 1.0, 1.0, 1.0), softplus=True):
    """Biases exactly as GaussianModel.setup_functions computes them (fp32 torch)."""
    x = torch.tensor(scaling_bias)
    scale_bias = (x + torch.log(-torch.expm1(-x))) if softplus else torch.log(x)
    p = torch.tensor(opacity_bias)
    op_bias = torch.log(p / (1 - p))
    return {"aabb": tuple(float(a) for a in aabb), "scale_bias": float(scale_bias),
            "min_kernel": float(min_kernel), "opacity_bias": float(op_bias),
            "softplus": bool(softplus)}


def activate(canon, delta, const):
    """canon: dict _xyz (P,3) _features_dc (P,1,3) _scaling (P,3) _rotation (P,4) _opacity (P,1);
    delta (P,14) or None -> means3D, scales, rotations, shs (P,1,3), opacity (P,1)."""
    aabb = torch.tensor(const["aabb"], dtype=torch.float32)
    z = lambda a, b: 0 if delta is None else delta[..., a:b]
    xyz = canon["_xyz"] * aabb[None, 3:] + aabb[None, :3] + z(0, 3)
    act = F.softplus if const["softplus"] else torch.exp
    s = act(canon["_scaling"] + const["scale_bias"] + z(3, 6))
    scales = torch.sqrt(torch.square(s) + const["min_kernel"] ** 2)
    rb = torch.tensor([1.0, 0, 0, 0])
    rots = F.normalize(canon["_rotation"] + rb[None] + z(6, 10))
    shs = canon["_features_dc"] + (0 if delta is None else delta[..., 10:13].unsqueeze(1))
    opac = torch.sigmoid(canon["_opacity"] + const["opacity_bias"] + z(13, 14))
    return xyz, scales, rots, shs, opac


def gaussian_tensor(canon, const):
    """train_vae.py:466-472 get_gaussian_tensor: [xyz3 | rgb3 | opacity1 | scale3 | rot4]."""
    xyz, scales, rots, shs, opac = activate(canon, None, const)
    return torch.cat([xyz, shs.squeeze(1), opac, scales, rots], dim=-1)


def intrinsics_to_projection(intr, near, far):
    fx, fy, cx, cy = intr[0, 0], intr[1, 1], intr[0, 2], intr[1, 2]
    ret = torch.zeros((4, 4), dtype=intr.dtype)
    ret[0, 0] = 2 * fx
    ret[1, 1] = 2 * fy
    ret[0, 2] = 2 * cx - 1
    ret[1, 2] = -2 * cy + 1
    ret[2, 2] = far / (far - near)
    ret[2, 3] = near * far / (near - far)
    ret[3, 2] = 1.0
    return ret


def camera_matrices(extrinsics, intrinsics, near, far):
    """-> (viewmatrix.T, full_proj.T, campos, tanfovx, tanfovy) as handed to the rasteriser."""
    view = extrinsics
    persp = intrinsics_to_projection(intrinsics, near, far)
    campos = torch.inverse(view)[:3, 3]
    fovx = 2 * torch.atan(0.5 / intrinsics[0, 0])
    fovy = 2 * torch.atan(0.5 / intrinsics[1, 1])
    return (view.T.contiguous(), (persp @ view).T.contiguous(), campos,
            math.tan(fovx * 0.5), math.tan(fovy * 0.5))


def orbit_camera(elevation, azimuth, radius=1.0, opengl=True):
    """kiui.cam.orbit_camera (degrees, look-at origin, y-up) restated: camera-to-world 4x4."""
    el, az = np.deg2rad(elevation), np.deg2rad(azimuth)
    x = radius * np.cos(el) * np.sin(az)
    y = -radius * np.sin(el)
    z = radius * np.cos(el) * np.cos(az)
    campos = np.array([x, y, z], dtype=np.float32)
    nrm = lambda v: v / (np.linalg.norm(v) + 1e-20)
    fwd = nrm(campos) if opengl else nrm(-campos)   # camera looks along -z in OpenGL
    up = np.array([0, 1, 0], dtype=np.float32)
    right = nrm(np.cross(up, fwd))
    up = nrm(np.cross(fwd, right))
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = np.stack([right, up, fwd], axis=1)
    T[:3, 3] = campos
    return T


def inference_extrinsics(azimuth, elevation=0.0, radius=2.0):
    """utils/inference_utils.py:245-254: world->camera matrix of one orbit view."""
    convert = np.array([[1, 0, 0, 0], [0, 0, -1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.float32)
    pose = convert @ orbit_camera(elevation, azimuth, radius=radius, opengl=True)
    pose[:3, 1:3] *= -1
    return torch.from_numpy(np.linalg.inv(pose)).float()


def intrinsics_from_fov(fov_deg=49.1):
    """utils3d.torch.intrinsics_from_fov_xy normalised intrinsics (dataset_latent_inference.py:182,212)."""
    f = 0.5 / math.tan(math.radians(fov_deg) / 2)
    return torch.tensor([[f, 0, 0.5], [0, f, 0.5], [0, 0, 1]], dtype=torch.float32)
