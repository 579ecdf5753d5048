"""`SparseVAE.to_representation` on the device: the canonical-Gaussian end of the static VAE, mirroring the
reference's model/sparse_voxel_diffusion/sparse_vae.py:60-112 (configuration, Hammersley perturbation),
:114-180 (to_representation) and :202-227 (feature-row layout).  The trunk that produces the feature rows is
`gvfdiffusion_b200.sparse.transformer.SparseTransformerVAE`.  One kernel launch per representation type covers the whole
batch; the per-entry `GaussianModel`s returned are views into its outputs.  `training_losses` (:303-362) is the static-VAE
half of the reference's train step (BASELINE configs[4]): backbone forward / backward on the library's kernels as one
autograd node, to_representation, one render per sample, L1 + lambda_ssim (1 - SSIM) + lambda_lpips LPIPS (VGG16 on library
convolutions, utils/lpips; seeded-random weights unless a module with the downloaded ones is supplied) + lamda_kl KL +
volume / opacity regularisers."""
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
    def __init__(self, backbones=None, resolution=64, representation_config=None, device="cuda", loss_type="l1",
                 lambda_ssim=0.2, lambda_lpips=0.0, lamda_kl=1e-6, regularizations=None, mem_ratio=1.0, lpips=None):
        self.backbones = backbones or {}
        self.resolution = resolution
        self.loss_type, self.lambda_ssim, self.lambda_lpips, self.lamda_kl = loss_type, lambda_ssim, lambda_lpips, lamda_kl
        self.regularizations = regularizations or {}
        self.mem_ratio = mem_ratio        # accepted; every activation is kept (180 GB of HBM: no checkpointing needed)
        self.lpips = lpips
        self.device = torch.device(device)
        self.rep_config = {}
        for k, v in (representation_config or {}).items():
            if k not in _DEFAULT_CONFIG:
                raise ValueError(f"Invalid representation type: {k}")
            self.rep_config[k] = copy.deepcopy(_DEFAULT_CONFIG[k])
            self.rep_config[k].update(v)
        self._calc_layout(self.rep_config)
        self._init_renderer()
        self.perturbation = {k: self._build_perturbation(v["num_gaussians"], v["reg_mode"])
                             for k, v in self.rep_config.items() if v["perturb_offset"]}

    def get_renderer(self, type, rendering_options):
        """:184-194"""
        from ...renderers.gaussian_render import GaussianRenderer
        renderer = GaussianRenderer(rendering_options)
        if type == "MipGS":
            renderer.pipe.use_mip_gaussian = True
            renderer.pipe.kernel_size = self.rep_config["MipGS"]["2d_filter_kernel_size"]
        elif type != "GS":
            raise ValueError(f"Invalid representation type: {type}")
        return renderer

    def _init_renderer(self):
        """:196-200"""
        opts = {"near": 0.8, "far": 1.6, "bg_color": (1.0, 1.0, 1.0)}
        self.renderers = {k: self.get_renderer(k, opts) for k in self.rep_config.keys()}

    def render_batch(self, reps, extrinsics, intrinsics):
        """One render per (representation type, batch entry) (:281-301) -> {type: {'rgb' [N,3,H,W], 'alpha' [N,H,W],
        'bg_color' [N,3]}}."""
        ret = {}
        for k, v in reps.items():
            packs = [self.renderers[k].render(rep, extrinsics[i], intrinsics[i]) for i, rep in enumerate(v)]
            ret[k] = {kk: torch.stack([p[kk] for p in packs], 0) for kk in packs[0].keys()}
            ret[k]["bg_color"] = torch.stack([self.renderers[k].bg_color] * len(packs), 0)
        return ret

    def get_regularization_loss(self, x, reps):
        """Volume and opacity regularisers (:228-248).  The activated scales / opacities are re-derived from the raw
        tensors with torch expressions here so that autograd reaches to_representation's backward: two elementwise
        maps over [P, 3] / [P, 1], not part of the rendered path."""
        import torch.nn.functional as F
        loss, terms = 0.0, {}
        for k, v in reps.items():
            reg = self.regularizations.get(k)
            if not reg:
                continue
            if "lambda_vol" in reg:
                sc = []
                for g in v:
                    s = g._scaling + g.scale_bias
                    s = F.softplus(s) if g.scaling_activation_type == "softplus" else torch.exp(s)
                    sc.append(torch.sqrt(torch.square(s) + g.mininum_kernel_size ** 2))
                terms[f"reg_{k}_vol"] = torch.prod(torch.cat(sc, 0), dim=1).mean()
                loss = loss + reg["lambda_vol"] * terms[f"reg_{k}_vol"]
            if "lambda_opacity" in reg:
                op = torch.cat([torch.sigmoid(g._opacity + g.opacity_logit_bias) for g in v], 0)
                terms[f"reg_{k}_opacity"] = (op - 1).pow(2).mean()
                loss = loss + reg["lambda_opacity"] * terms[f"reg_{k}_opacity"]
        return loss, terms

    def _backbone_forward(self, feats, noise=None):
        """`self.backbones['vae'](feats)` of :318 -> (out [Nvox, out_channels], kl, mean, logvar), out and kl on the graph."""
        vae = self.backbones["vae"]
        if isinstance(vae, torch.nn.Module):             # the nn.Module mirror: parameter gradients through autograd
            out, mean, logvar = vae(feats, mem_ratio=self.mem_ratio, noise=noise)
            return out.feats, vae.kl, mean, logvar
        from ...sparse.transformer import sparse_vae_forward_autograd     # bare engine: gradients on engine.grads
        return sparse_vae_forward_autograd(vae, feats.feats, feats.coords, noise)

    def training_losses(self, feats, image, extrinsics, intrinsics, return_aux=False, noise=None, **kwargs):
        """feats: SparseTensor [Nvox, in_channels]; image [N,3,H,W]; extrinsics [N,4,4]; intrinsics [N,3,3] ->
        (terms, reps) with terms['loss'] a scalar attached to the autograd graph (:303-362).  After
        `terms['loss'].backward()` the backbone's parameter gradients are on `backbones['vae'].grads`.
        noise: the posterior's randn_like draw (host RNG when None)."""
        from ...utils.loss_util import l2_loss, ssim, ssim_l1
        out, kl, mean, logvar = self._backbone_forward(feats, noise)
        x = feats.replace(out)
        reps = self.to_representation(x)
        for v in self.renderers.values():
            v.rendering_options.resolution = image.shape[-1]
        render_results = self.render_batch(reps, extrinsics, intrinsics)
        terms = {"loss": 0.0, "rec": 0.0}
        for k, rr in render_results.items():
            rec = rr["rgb"]
            if self.loss_type == "l1":
                s, l1 = ssim_l1(rec, image)                    # one kernel for both terms
                terms[k + "_l1"] = l1
                terms["rec"] = terms["rec"] + l1
            elif self.loss_type == "l2":
                terms[k + "_l2"] = l2_loss(rec, image)
                terms["rec"] = terms["rec"] + terms[k + "_l2"]
                s = ssim(rec, image) if self.lambda_ssim > 0 else None
            else:
                raise ValueError(f"Invalid loss type: {self.loss_type}")
            if self.lambda_ssim > 0:
                terms[k + "_ssim"] = 1 - s
                terms["rec"] = terms["rec"] + self.lambda_ssim * terms[k + "_ssim"]
            if self.lambda_lpips > 0:
                # utils/loss_util.py:66-74; `lpips=` an LPIPS module with real weights, else the seeded-random one
                from ...utils.loss_util import lpips as lpips_fn
                terms[k + "_lpips"] = lpips_fn(rec, image, module=self.lpips)
                terms["rec"] = terms["rec"] + self.lambda_lpips * terms[k + "_lpips"]
            terms["loss"] = terms["loss"] + terms["rec"]
        terms["kl"] = kl
        terms["loss"] = terms["loss"] + self.lamda_kl * kl
        reg_loss, reg_terms = self.get_regularization_loss(x, reps)
        terms.update(reg_terms)
        terms["loss"] = terms["loss"] + reg_loss
        if return_aux:
            rec_image = torch.cat([v["rgb"] for v in render_results.values()])
            return terms, reps, {"rec_image": rec_image, "gt_image": torch.cat([image for _ in render_results])}
        return terms, reps

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
