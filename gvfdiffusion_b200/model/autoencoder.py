"""`GSKLTemporalVariationalAutoEncoder`: drop-in for the DECODE side of the reference's motion
VAE (model/autoencoder.py:345-609): same constructor keywords, same state-dict names for the
decode weights (`proj`, `layers.{i}.{0,1}.fn.*`, `gs_embedding.0`, `decoder_cross_attn.fn.*`,
`to_outputs`), `decode(x, queries)` on the sm_100a engine -- forward only under no_grad, forward + hand-written
backward (vae_train.py) when autograd is recording.  `encode` (FPS + KNN interpolation + cross-attention +
DiagonalGaussian, model/autoencoder.py:502-550) runs forward on the engine of vae_encode.py.
"""
import torch
import torch.nn as nn

from ..vae_encode import VAEEncodeEngine
from .. import _param_epoch
from ..vae_engine import VAEDecodeEngine
from ..vae_train import VAEDecodeTrainEngine


class _DecodeFn(torch.autograd.Function):
    """decode() under autograd: forward with saved activations and the hand-written backward of vae_train.py (the
    reference relies on torch autograd through model/autoencoder.py:552-609 here, train_vae.py:293-353)."""

    @staticmethod
    def forward(ctx, module, z, queries, *params):
        eng = module.train_engine(force_refresh=not _param_epoch.HOOKED)
        out, saved = eng.forward_train(z, queries)
        ctx.eng, ctx.saved, ctx.names = eng, saved, module._param_names
        ctx.z_shape, ctx.q_dtype = z.shape, queries.dtype
        return out

    @staticmethod
    def backward(ctx, dout):
        grads, dz, dq = ctx.eng.backward(ctx.saved, dout)
        ctx.saved = None
        return (None, dz.view(ctx.z_shape), dq.to(ctx.q_dtype)) + tuple(grads[n] for n in ctx.names)


class _EncodeFn(torch.autograd.Function):
    """encode() under autograd: gradients of the sampled latent and of the KL term flow to the ENCODER parameters; the
    point clouds and Gaussians are data (the reference interpolates them under no_grad, model/autoencoder.py:470)."""

    @staticmethod
    def forward(ctx, module, static_pc, delta_pc, gs_list, noise, *params):
        eng = module.encode_engine(force_refresh=not _param_epoch.HOOKED)
        o, saved = eng.forward_train(static_pc, delta_pc, gs_list, noise)
        ctx.eng, ctx.saved, ctx.names = eng, saved, module._enc_names
        ctx.mark_non_differentiable(o["mean"], o["logvar"], o["sampled_static_gs"])
        return o["kl"], o["x"], o["mean"], o["logvar"], o["sampled_static_gs"]

    @staticmethod
    def backward(ctx, dkl, dx, *_):
        grads = ctx.eng.backward(ctx.saved, dx, dkl)
        ctx.saved = None
        return (None, None, None, None, None) + tuple(grads[n] for n in ctx.names)


class _Fn(nn.Module):
    def __init__(self, fn):
        super().__init__()
        self.fn = fn


class _Attn(nn.Module):
    def __init__(self, qd, cd, inner):
        super().__init__()
        self.to_q = nn.Linear(qd, inner, bias=False)
        self.to_kv = nn.Linear(cd, inner * 2, bias=False)
        self.to_out = nn.Linear(inner, qd)


class _FF(nn.Module):
    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dim, dim * mult * 2), nn.Identity(), nn.Linear(dim * mult, dim))


class GSKLTemporalVariationalAutoEncoder(nn.Module):
    def __init__(self, *, depth=24, dim=512, queries_dim=512, input_dim=3, gs_dim=14, output_dim=10,
                 num_inputs=8192, num_latents=1024, latent_dim=128, heads=8, dim_head=-1, weight_tie_layers=False,
                 decoder_ff=False, enable_flash_attn=False, num_timesteps=24, chunk_size=8192, knn_k=8, beta=7.0):
        super().__init__()
        if decoder_ff or weight_tie_layers or queries_dim != dim:
            raise NotImplementedError("shipped config: decoder_ff=False, weight_tie_layers=False, queries_dim=dim")
        if dim_head == -1:
            dim_head = dim // heads
        if dim_head * heads != dim or dim_head not in (32, 64):
            raise NotImplementedError("the sm_100a attention kernels cover head dims 32 and 64")
        self.depth, self.dim, self.heads, self.num_timesteps, self.chunk_size = depth, dim, heads, num_timesteps, chunk_size
        self.num_latents, self.num_inputs, self.knn_k, self.beta = num_latents, num_inputs, knn_k, beta
        # encoder (model/autoencoder.py:385-415): one cross-attention block + feed-forward, token embedding, heads
        self.cross_attend_blocks = nn.ModuleList([_Fn(_Attn(dim, dim, dim)), _Fn(_FF(dim))])
        self.input_embedding = nn.Sequential(nn.Linear(input_dim, dim), nn.LayerNorm(dim, elementwise_affine=False))
        self.mean_fc = nn.Linear(dim, latent_dim)
        self.logvar_fc = nn.Linear(dim, latent_dim)
        self.layers = nn.ModuleList([nn.ModuleList([_Fn(_Attn(dim, dim, dim)), _Fn(_FF(dim))]) for _ in range(depth)])
        self.gs_embedding = nn.Sequential(nn.Linear(gs_dim, dim), nn.LayerNorm(dim, elementwise_affine=False))
        self.decoder_cross_attn = _Fn(_Attn(queries_dim, dim, dim))
        self.to_outputs = nn.Linear(queries_dim, output_dim)
        self.proj = nn.Linear(latent_dim, dim)
        for m in self.modules():                       # reference _init_weights :422-436
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
        nn.init.constant_(self.to_outputs.weight, 0)
        nn.init.constant_(self.to_outputs.bias, 0)
        self._engine, self._sig = None, None
        self._train_engine, self._train_sig = None, None
        self._enc_engine, self._enc_sig = None, None
        self._encoder_loaded = True                     # False after loading a decode-only checkpoint
        self._ENC = ("cross_attend_blocks.", "input_embedding.", "mean_fc.", "logvar_fc.")
        # decode() is differentiable with respect to the DECODE parameters (the encoder is forward-only here)
        self._param_names = [n for n, _ in self.named_parameters() if not n.startswith(self._ENC)]
        self._enc_names = [n for n, _ in self.named_parameters() if n.startswith(self._ENC)]

    def load_state_dict(self, state_dict, strict=False, **kw):
        # reference checkpoints also carry the encoder; only the decode weights are consumed here
        own = self.state_dict()
        sub = {k: v for k, v in state_dict.items() if k in own}
        missing = [k for k in own if k not in sub]
        if any(not k.startswith(self._ENC) for k in missing):
            raise KeyError(f"decode weights missing from checkpoint: {[k for k in missing if not k.startswith(self._ENC)][:4]}...")
        self._encoder_loaded = not missing              # decode-only checkpoints are fine for decode()
        return super().load_state_dict(sub, strict=False)

    def train(self, mode=True):
        """Switching between training and evaluation invalidates the engines' weight copies: parameters updated by a fused
        optimiser carry no version bump (see train_engine), so the inference engine must not trust its signature."""
        self._sig = self._enc_sig = self._train_sig = None
        return super().train(mode)

    def refresh_engines(self):
        """Force every existing engine to re-read the parameters at its next use."""
        self._sig = self._enc_sig = self._train_sig = None

    def engine(self):
        sig = (_param_epoch.epoch(),) + tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._engine is None or sig != self._sig:
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("decode runs on a CUDA device only (no CPU fallback)")
            if self._engine is not None and self._engine.dev == dev:
                self._engine.refresh(self.state_dict())          # same buffers, new values (optimiser step)
            else:
                self._engine = VAEDecodeEngine(self.state_dict(), self.heads, self.num_timesteps, dev, self.chunk_size)
            self._sig = sig
        return self._engine

    def encode_engine(self, force_refresh=False):
        if not self._encoder_loaded:
            raise RuntimeError("this checkpoint carried no encoder weights (cross_attend_blocks / input_embedding / mean_fc / "
                               "logvar_fc)")
        sig = (_param_epoch.epoch(),) + tuple((p.data_ptr(), p._version) for n, p in self.named_parameters() if n.startswith(self._ENC))
        if self._enc_engine is None or sig != self._enc_sig or force_refresh:
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("encode runs on a CUDA device only (no CPU fallback)")
            if self._enc_engine is not None and self._enc_engine.dev == dev:
                self._enc_engine.refresh(self.state_dict())
            else:
                self._enc_engine = VAEEncodeEngine(self.state_dict(), self.heads, self.num_latents, self.knn_k, self.beta, dev)
            self._enc_sig = sig
        return self._enc_engine

    def train_engine(self, force_refresh=False):
        """The signature that decides whether the engine's copies are current = optimiser-step counter (_param_epoch: torch's
        FUSED optimisers update parameters without bumping `_version` -- measured with AdamW(fused=True), the engines kept
        stepping on the initial weights) + every parameter's (data_ptr, _version).  force_refresh re-reads regardless."""
        sig = (_param_epoch.epoch(),) + tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._train_engine is None or sig != self._train_sig or force_refresh:
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("decode runs on a CUDA device only (no CPU fallback)")
            if self._train_engine is not None and self._train_engine.dev == dev:
                self._train_engine.refresh(self.state_dict())
            else:
                self._train_engine = VAEDecodeTrainEngine(self.state_dict(), self.heads, self.num_timesteps, dev)
            self._train_sig = sig
        return self._train_engine

    def decode(self, x, queries):
        """x ((B*T), L, latent_dim), queries (B, Q, 14) -> (B, T, Q, output_dim) fp32.  Differentiable with respect to
        x, queries and every decode parameter when autograd is recording (training step); otherwise the inference
        engine runs."""
        named = dict(self.named_parameters())
        params = [named[n] for n in self._param_names]                       # the decode parameters, in gradient order
        if torch.is_grad_enabled() and (x.requires_grad or queries.requires_grad or any(p.requires_grad for p in params)):
            return _DecodeFn.apply(self, x, queries, *params)
        with torch.no_grad():
            return self.engine().decode(x, queries)

    def encode(self, static_pc, delta_pc, static_gs_list, noise=None):
        """static_pc (B, N, 3), delta_pc (B, T, N, 3), static_gs_list [P_b x 14] -> (kl, x, posterior, sampled_static_gs) like
        the reference (model/autoencoder.py:502-550); `posterior` is a dict with mean / logvar.  Differentiable with respect
        to the encoder parameters when autograd is recording."""
        named = dict(self.named_parameters())
        params = [named[n] for n in self._enc_names]
        if torch.is_grad_enabled() and self._encoder_loaded and any(p.requires_grad for p in params):
            kl, x, mean, logvar, sgs = _EncodeFn.apply(self, static_pc, delta_pc, static_gs_list, noise, *params)
            return kl, x, {"mean": mean, "logvar": logvar}, sgs
        with torch.no_grad():
            o = self.encode_engine().encode(static_pc, delta_pc, static_gs_list, noise)
        return o["kl"], o["x"], {"mean": o["mean"], "logvar": o["logvar"]}, o["sampled_static_gs"]

    def forward(self, static_gs, static_pc, delta_pc, noise=None):
        """model/autoencoder.py:620-627: encode -> pad_static_gs -> decode.  noise: the posterior's randn draw (the
        reference draws it on the host, :316; a device tensor here avoids that blocking copy)."""
        from ..pipeline import pad_static_gs
        kl, x, posterior, _ = self.encode(static_pc, delta_pc, static_gs, noise)
        padded, _ = pad_static_gs([g.to(x.device) for g in static_gs])
        return {"logits": self.decode(x, padded), "kl": kl, "posterior": posterior}
