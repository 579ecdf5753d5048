"""`SparseTransformerVAE`: the static (canonical-Gaussian) VAE backbone as an nn.Module with the reference's constructor,
parameter names and call surface (model/sparse_voxel_diffusion/sparse_transformer_vae.py:14-212), over the device engine
of gvfdiffusion_b200/sparse/transformer.py.

The Parameters are fp32 master weights (what the optimiser and the EMA see, as in the reference's mixed-precision set-up);
the engine holds the fp16 compute copies (+ the transposes its backward GEMMs read) and is refreshed IN PLACE -- same
buffers, one foreach cast -- whenever a parameter's version counter has moved (i.e. after every optimiser step).
`forward` under autograd is one node (csrc/sparse_trunk.cu forward / backward) returning the gradient of every parameter.
Only the shipped configuration family is built: attn_mode "swin", pe_mode "ape", head dim 64.
"""
import torch
import torch.nn as nn

from ... import _param_epoch, ops
from ...sparse.basic import SparseTensor
from ...sparse.transformer import SparseTransformerVAE as _Engine

_BLOCK = (("attn.to_qkv", 3, 1), ("attn.to_out", 1, 1), ("mlp.mlp.0", None, 1), ("mlp.mlp.2", 1, None))


class _TrainFn(torch.autograd.Function):
    """encode -> posterior sample -> decode with every activation kept; backward on the library's kernels."""

    @staticmethod
    def forward(ctx, module, feats, coords, noise, *params):
        eng = module.engine(force_refresh=not _param_epoch.HOOKED)
        out, mean, logvar, kl, saved = eng.forward_train(feats, coords, noise)
        ctx.eng, ctx.saved, ctx.names = eng, saved, module._names
        ctx.mark_non_differentiable(mean, logvar)
        return out, kl.reshape(()), mean, logvar

    @staticmethod
    def backward(ctx, dout, dkl, _dm, _dl):
        g = ctx.eng.backward(ctx.saved, dout, dkl)
        ctx.saved = None
        return (None, None, None, None) + tuple(g[n] for n in ctx.names)


class SparseTransformerVAE(nn.Module):
    def __init__(self, resolution, in_channels, model_channels, out_channels, latent_channels, num_blocks, window_size=1024,
                 num_heads=None, num_head_channels=64, mlp_ratio=4, attn_mode="swin", pe_mode="ape", use_fp16=False,
                 use_checkpoint=False, use_old_attn_impl=True, norm_output=False):
        super().__init__()
        if attn_mode != "swin" or pe_mode != "ape":
            raise NotImplementedError('shipped configs/vae.yml: attn_mode "swin", pe_mode "ape"')
        self.resolution, self.in_channels, self.model_channels = resolution, in_channels, model_channels
        self.out_channels, self.latent_channels, self.num_blocks = out_channels, latent_channels, num_blocks
        self.window_size, self.mlp_ratio = window_size, mlp_ratio
        self.num_heads = num_heads or model_channels // num_head_channels
        self.use_fp16, self.norm_output, self.use_old_attn_impl = use_fp16, norm_output, use_old_attn_impl
        self.use_checkpoint = use_checkpoint      # accepted; the engine keeps every activation (180 GB of HBM)
        C, F_ = model_channels, int(model_channels * mlp_ratio)
        shapes = {"input_layer": (C, in_channels), "to_latent": (2 * latent_channels, C), "from_latent": (C, latent_channels),
                  "out_layer": (out_channels, C)}
        for side in ("encoder", "decoder"):
            for i in range(num_blocks):
                shapes.update({f"{side}.{i}.attn.to_qkv": (3 * C, C), f"{side}.{i}.attn.to_out": (C, C),
                               f"{side}.{i}.mlp.mlp.0": (F_, C), f"{side}.{i}.mlp.mlp.2": (C, F_)})
        # reference parameter names contain dots: registered through nested ParameterDict-free plain modules
        self._names = []
        for name, (o, k) in shapes.items():
            w, b = nn.Parameter(torch.empty(o, k)), nn.Parameter(torch.zeros(o))
            nn.init.xavier_uniform_(w)                                  # initialize_weights (:120-146)
            if name in ("to_latent", "out_layer"):
                nn.init.constant_(w, 0)
            self._register(name + ".weight", w)
            self._register(name + ".bias", b)
            self._names += [name + ".weight", name + ".bias"]
        self._engine, self._sig = None, None

    def _register(self, dotted, p):
        mod = self
        parts = dotted.split(".")
        for q in parts[:-1]:
            if not hasattr(mod, q):
                mod.add_module(q, nn.Module())
            mod = getattr(mod, q)
        mod.register_parameter(parts[-1], p)

    @property
    def device(self):
        return next(self.parameters()).device

    def freeze_encoder(self):
        for n, p in self.named_parameters():
            if n.startswith("encoder."):
                p.requires_grad_(False)

    # ---------------------------------------------------------------------------------------- engine
    def train(self, mode=True):
        self._sig = None              # fused optimisers do not bump version counters: re-read the weights after a mode switch
        return super().train(mode)

    def refresh_engines(self):
        self._sig = None

    def engine(self, force_refresh=False):
        """The device engine over the current parameter values (fp16 copies refreshed in place when they changed).
        "Changed" = any optimiser step since (gvfdiffusion_b200/_param_epoch.py: torch's FUSED optimisers update parameters
        without bumping their version counters) or a moved (data_ptr, _version); force_refresh re-reads regardless."""
        named = dict(self.named_parameters())
        sig = (_param_epoch.epoch(),) + tuple((p.data_ptr(), p._version) for p in named.values())
        if self._engine is None:
            if self.device.type != "cuda":
                raise RuntimeError("SparseTransformerVAE runs on a CUDA device only (no CPU fallback)")
            self._engine = _Engine({k: v.detach() for k, v in named.items()}, self.num_blocks, self.num_heads, self.window_size,
                                   use_fp16=self.use_fp16, norm_output=self.norm_output, device=self.device,
                                   use_old_attn_impl=self.use_old_attn_impl)
        elif sig != self._sig or force_refresh:
            self._engine.refresh({k: v.detach() for k, v in named.items()})
        self._sig = sig
        return self._engine

    # ---------------------------------------------------------------------------------------- reference call surface
    def encode(self, x, sample_posterior=True, return_raw=False, noise=None):
        eng = self.engine()
        with torch.no_grad():
            mean, logvar = eng.encode(x.feats, x.coords)
            z = mean
            if sample_posterior:
                eps = torch.randn(mean.shape).to(mean.device) if noise is None else noise.to(mean.device)
                z = mean + torch.exp(0.5 * logvar) * eps
        z = x.replace(z.contiguous())
        return (z, mean, logvar) if return_raw else z

    def decode(self, latent):
        with torch.no_grad():
            return latent.replace(self.engine().decode(latent.feats, latent.coords))

    def forward(self, x, t=None, c=None, mem_ratio=1.0, noise=None):
        """`_forward_with_mem_ratio` (:204-210): -> (out SparseTensor, mean, logvar); `self.kl` holds the KL term of
        sparse_vae.py:351 attached to the graph.  mem_ratio is accepted (nothing is recomputed here)."""
        if not (torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())):
            z, mean, logvar = self.encode(x, sample_posterior=True, return_raw=True, noise=noise)
            self.kl = 0.5 * torch.mean(mean.pow(2) + logvar.exp() - logvar - 1)
            return self.decode(z), mean, logvar
        named = dict(self.named_parameters())
        out, kl, mean, logvar = _TrainFn.apply(self, x.feats, x.coords, noise, *[named[n] for n in self._names])
        self.kl = kl
        return x.replace(out), mean, logvar
