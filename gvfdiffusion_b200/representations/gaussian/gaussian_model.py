"""`GaussianModel`: the container the renderer consumes, mirroring the reference's
representations/gaussian/gaussian_model.py:15-128 for the inference path (raw tensors `_xyz`,
`_features_dc`, `_scaling`, `_rotation`, `_opacity`, the bias constants of setup_functions :23-41
and the activated getters).  Activations are evaluated by the sm_100a kernels
(`gvf_gaussian_tensor`; the per-frame `*_with_delta` variants are fused into the rasteriser's
preprocess), not by torch.  PLY I/O and the training-time helpers are out of scope."""
import torch

from ... import ops
from ... import raster as R


_CONST_CACHE = {}


def _device_const(values, device):
    """Small constant tensors (aabb, ...) built once per (values, device): torch.tensor(..., device='cuda') is a blocking
    host-to-device copy, i.e. a stream synchronisation per GaussianModel otherwise."""
    key = (tuple(float(v) for v in values), str(device))
    t = _CONST_CACHE.get(key)
    if t is None:
        t = torch.tensor(key[0], dtype=torch.float32, device=device)
        _CONST_CACHE[key] = t
    return t


class GaussianModel:
    def __init__(self, sh_degree: int = 0, aabb=(-0.5, -0.5, -0.5, 1.0, 1.0, 1.0), mininum_kernel_size: float = 0.0,
                 scaling_bias: float = 0.01, opacity_bias: float = 0.1, scaling_activation: str = "exp",
                 device="cuda"):
        if sh_degree != 0:
            raise NotImplementedError("the inference path uses sh_degree 0")
        self.active_sh_degree, self.max_sh_degree = 0, sh_degree
        self.mininum_kernel_size = mininum_kernel_size
        self.scaling_bias, self.opacity_bias = scaling_bias, opacity_bias
        self.scaling_activation_type = scaling_activation
        self.device = device
        self._aabb = _device_const(aabb, device)
        self._aabb_host = tuple(float(a) for a in aabb)     # constants() must not read the device tensor back
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=device)
        self._xyz, self._features_dc, self._scaling = z(8, 3), z(8, 1, 3), z(8, 3)
        self._rotation, self._opacity = z(8, 4), z(8, 1)
        # setup_functions (:23-41): biases exactly as the reference derives them (fp32 torch scalars)
        key = ("bias", float(scaling_bias), float(opacity_bias), scaling_activation)
        if key not in _CONST_CACHE:
            x = torch.tensor(scaling_bias)
            sb = (x + torch.log(-torch.expm1(-x))) if scaling_activation == "softplus" else torch.log(x)
            p = torch.tensor(opacity_bias)
            _CONST_CACHE[key] = (float(sb), float(torch.log(p / (1 - p))))
        self.scale_bias, self.opacity_logit_bias = _CONST_CACHE[key]

    @property
    def aabb(self):
        return self._aabb

    @aabb.setter
    def aabb(self, value):                                  # assigning a new box re-reads it once
        self._aabb = value
        self._aabb_host = tuple(float(a) for a in value.tolist())

    def constants(self):
        return {"aabb": self._aabb_host, "scale_bias": self.scale_bias,
                "min_kernel": float(self.mininum_kernel_size), "opacity_bias": self.opacity_logit_bias,
                "softplus": self.scaling_activation_type == "softplus"}

    def raw(self):
        return {"_xyz": self._xyz, "_features_dc": self._features_dc, "_scaling": self._scaling,
                "_rotation": self._rotation, "_opacity": self._opacity}

    def from_xyz(self, xyz):
        """activated positions -> raw `_xyz` (gaussian_model.py:137-138)"""
        self._xyz = (xyz - self.aabb[None, :3]) / self.aabb[None, 3:]

    def from_rotation(self, rots):
        """activated (w-first) quaternions -> raw `_rotation` (gaussian_model.py:134-135; rots_bias = [1, 0, 0, 0])"""
        bias = torch.zeros(4, dtype=rots.dtype, device=rots.device)
        bias[0] = 1.0
        self._rotation = rots - bias[None, :]

    def gaussian_tensor(self):
        """[P,14] = [xyz3 | rgb3 | opacity1 | scale3 | rot4] (train_vae.py:466-472)."""
        prm = R.make_params(16, 16, 1.0, 1.0, self.constants())
        raw = self.raw()
        if torch.is_grad_enabled() and any(t.requires_grad for t in raw.values()):
            P = raw["_xyz"].shape[0]                     # training: the queries carry gradients back to the static VAE
            return ops.GaussianTensorFn.apply(prm, raw["_xyz"].reshape(P, 3), raw["_features_dc"].reshape(P, 3),
                                              raw["_scaling"].reshape(P, 3), raw["_rotation"].reshape(P, 4),
                                              raw["_opacity"].reshape(P))
        return ops.gaussian_tensor(prm, R.canon_arrays(raw, self._xyz.device))

    @property
    def get_xyz(self):
        return self.gaussian_tensor()[:, 0:3]

    @property
    def get_features(self):
        return self._features_dc

    @property
    def get_opacity(self):
        return self.gaussian_tensor()[:, 6:7]

    @property
    def get_scaling(self):
        return self.gaussian_tensor()[:, 7:10]

    @property
    def get_rotation(self):
        return self.gaussian_tensor()[:, 10:14]
