"""`SparseVAE.to_representation` on the device: the canonical-Gaussian end of the static VAE, mirroring the
reference's model/sparse_voxel_diffusion/sparse_vae.py:60-112 (configuration, Hammersley perturbation),
:114-180 (to_representation) and :202-227 (feature-row layout).  The trunk that produces the feature rows is
`gvfdiffusion_b200.sparse.transformer.SparseTransformerVAE.decode`; training (losses, regularisers, optimiser
phases) is out of scope.  One kernel launch per representation type covers the whole batch; the per-entry
`GaussianModel`s returned are views into its outputs."""
import copy

import torch

from ... import ops
from ...representations.gaussian import GaussianModel

_DEFAULT_GAUSSIAN_LR_CONFIG = {"_xyz": 1.0, "_features_dc": 0.0025, "_opacity": 0.05, "_scaling": 0.005, "_rotation": 0.001}
_DEFAULT_GS_CFG = {"lr": _DEFAULT_GAUSSIAN_LR_CONFIG, "perturb_offset": False, "reg_mode": "invoxel", "voxel_size": 1.1,
                   "num_gaussians": 8, "scaling_bias": 0.01, "opacity_bias": 0.1, "scaling_activation": "exp"}
_DEFAULT_MIPGS_CFG = dict(_DEFAULT_GS_CFG, **{"2d_filter_kernel_size": 0.1, "3d_filter_kernel_size": 0.0})
_DEFAULT_CONFIG = {"GS": _DEFAULT_GS_CFG, "MipGS": _DEFAULT_MIPGS_CFG}
_PRIMES = [2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53]
_ORDER = ("_xyz", "_features_dc", "_scaling", "_rotation", "_opacity")
_WIDTH = {"_xyz": 3, "_features_dc": 3, "_scaling": 3, "_rotation": 4, "_opacity": 1}
_REG = {None: 0, "none": 0, "invoxel": 1, "soft_invoxel": 2}


def _radical_inverse(base, n):
    val, inv_base = 0.0, 1.0 / base
    inv_base_n = inv_base
    while n > 0:
        val += (n % base) * inv_base_n
        n //= base
        inv_base_n *= inv_base
    return val


def hammersley_sequence(dim, n, num_samples):
    return [n / num_samples] + [_radical_inverse(_PRIMES[d], n) for d in range(dim - 1)]


class _ToRepresentationFn(torch.autograd.Function):
    """to_representation under autograd (training_losses, reference sparse_vae.py:303-362, backpropagates the render loss
    through it into the decoder trunk): forward gvf_to_representation, backward gvf_to_representation_bwd."""

    @staticmethod
    def forward(ctx, feats, coords, G, lr, resolution, reg_mode, voxel_size, perturbation):
        f = feats.detach()
        ctx.save_for_backward(f, perturbation)
        ctx.cfg = (G, lr, resolution, reg_mode, voxel_size)
        return ops.to_representation(f, coords, G, lr, resolution, reg_mode, voxel_size, perturbation)

    @staticmethod
    def backward(ctx, g_xyz, g_dc, g_scaling, g_rotation, g_opacity):
        f, perturbation = ctx.saved_tensors
        G, lr, resolution, reg_mode, voxel_size = ctx.cfg
        d = ops.to_representation_bwd(f, G, lr, resolution, reg_mode, voxel_size, perturbation, g_xyz, g_dc, g_scaling,
                                      g_rotation, g_opacity)
        return d, None, None, None, None, None, None, None


class SparseVAE:
    def __init__(self, backbones=None, resolution=64, representation_config=None, device="cuda"):
        self.backbones = backbones or {}
        self.resolution = resolution
        self.device = torch.device(device)
        self.rep_config = {}
        for k, v in (representation_config or {}).items():
            if k not in _DEFAULT_CONFIG:
                raise ValueError(f"Invalid representation type: {k}")
            self.rep_config[k] = copy.deepcopy(_DEFAULT_CONFIG[k])
            self.rep_config[k].update(v)
        self._calc_layout(self.rep_config)
        self.perturbation = {k: self._build_perturbation(v["num_gaussians"], v["reg_mode"])
                             for k, v in self.rep_config.items() if v["perturb_offset"]}

    def _build_perturbation(self, num_gaussians, reg_mode):
        offsets = torch.tensor([hammersley_sequence(3, i, num_gaussians) for i in range(num_gaussians)]).float() - 0.5
        if reg_mode == "soft_invoxel":
            # the reference divides by the MipGS voxel size whatever the representation (:110); kept
            vs = self.rep_config.get("MipGS", next(iter(self.rep_config.values())))["voxel_size"]
            offsets = offsets / 0.5 / vs
        return torch.atanh(offsets).to(self.device).contiguous()

    def _calc_layout(self, rep_config):
        self.layouts, start = {}, 0
        for k, v in rep_config.items():
            G = v["num_gaussians"]
            self.layouts[k] = {}
            for name in _ORDER:
                size = G * _WIDTH[name]
                self.layouts[k][name] = {"size": size, "range": (start, start + size)}
                start += size
        self.out_channels = start

    def to_representation(self, x):
        """x: SparseTensor-like (`feats` fp32 [N, C], `coords` int32 [N,4], `layout`, `shape`) ->
        {'GS' | 'MipGS': [GaussianModel per batch entry]}."""
        if not x.feats.is_cuda:
            raise RuntimeError("to_representation runs on the device only (no CPU fallback)")
        feats = x.feats if x.feats.dtype == torch.float32 else x.feats.float()
        coords = x.coords.to(torch.int32).contiguous()
        ret = {}
        for k, cfg in self.rep_config.items():
            G = cfg["num_gaussians"]
            lo = self.layouts[k]["_xyz"]["range"][0]
            # the GS branch hard-codes 1.25 where MipGS uses its voxel_size (:153 / :175)
            vs = 1.25 if k == "GS" else cfg["voxel_size"]
            raw = _ToRepresentationFn.apply(feats[:, lo:lo + 14 * G], coords, G, tuple(cfg["lr"][n] for n in _ORDER),
                                            self.resolution, _REG[cfg["reg_mode"]], vs, self.perturbation.get(k))
            ret[k] = []
            for i in range(x.shape[0]):
                sl = x.layout[i]
                rep = GaussianModel(sh_degree=0, aabb=[-0.5, -0.5, -0.5, 1.0, 1.0, 1.0],
                                    mininum_kernel_size=cfg.get("3d_filter_kernel_size", 0.0) if k == "MipGS" else 0.0,
                                    scaling_bias=cfg["scaling_bias"], opacity_bias=cfg["opacity_bias"],
                                    scaling_activation=cfg["scaling_activation"], device=self.device)
                gs = slice(sl.start * G, sl.stop * G)
                rep._xyz, rep._features_dc, rep._scaling, rep._rotation, rep._opacity = (t[gs] for t in raw)
                ret[k].append(rep)
        return ret
