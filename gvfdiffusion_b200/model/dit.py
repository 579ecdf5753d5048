"""`DiT`: drop-in for the reference's `model.dit.DiT` (model/dit.py:306-480) on the inference
path -- same constructor arguments, same parameter / state-dict names (reference checkpoints
load with `load_state_dict`), same `forward(x, t, cond_images, static_latent,
deformation_position_xyz)` -- whose forward runs on the sm_100a engine (dit_engine.py).

The nn.Module tree below only owns the parameters; none of its submodules' forwards is used.
"""
from typing import Optional

import torch
import torch.nn as nn

from ..dit_engine import DiTEngine


class _RMSGamma(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(heads, dim))


class _SelfAttnParams(nn.Module):
    def __init__(self, C, H):
        super().__init__()
        self.to_qkv = nn.Linear(C, 3 * C)
        self.q_rms_norm = _RMSGamma(C // H, H)
        self.k_rms_norm = _RMSGamma(C // H, H)
        self.to_out = nn.Linear(C, C)


class _CrossAttnParams(nn.Module):
    def __init__(self, C, Cctx):
        super().__init__()
        self.to_q = nn.Linear(C, C)
        self.to_kv = nn.Linear(Cctx, 2 * C)
        self.to_out = nn.Linear(C, C)


class _MLPParams(nn.Module):
    def __init__(self, C, ratio):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(C, int(C * ratio)), nn.GELU(approximate="tanh"), nn.Linear(int(C * ratio), C))


class _BlockParams(nn.Module):
    def __init__(self, C, H, ratio):
        super().__init__()
        self.norm3 = nn.LayerNorm(C, elementwise_affine=True, eps=1e-6)
        self.norm4 = nn.LayerNorm(C, elementwise_affine=True, eps=1e-6)
        self.spatial_self_attn = _SelfAttnParams(C, H)
        self.temporal_self_attn = _SelfAttnParams(C, H)
        self.image_cross_attn = _CrossAttnParams(C, C)
        self.static_cross_attn = _CrossAttnParams(C, C)
        self.mlp = _MLPParams(C, ratio)
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(C, 6 * C))
        self.adaLN_modulation_temporal = nn.Sequential(nn.SiLU(), nn.Linear(C, 3 * C))


class _TimestepEmbedderParams(nn.Module):
    def __init__(self, C, F=256):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(F, C), nn.SiLU(), nn.Linear(C, C))


class _FinalLayerParams(nn.Module):
    def __init__(self, C, O):
        super().__init__()
        self.linear = nn.Linear(C, O)
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(C, 2 * C))


class DiT(nn.Module):
    def __init__(self, resolution: int, in_channels: int, model_channels: int, static_cond_channels: int,
                 image_cond_channels: int, out_channels: int, num_blocks: int, num_heads: Optional[int] = None,
                 num_head_channels: Optional[int] = 64, mlp_ratio: float = 4, patch_size: int = 1,
                 pe_mode: str = "learnable", use_fp16: bool = False, use_checkpoint: bool = False,
                 use_skip_connection: bool = True, share_mod: bool = False, qk_rms_norm: bool = False,
                 qk_rms_norm_cross: bool = False, no_temporal_attn: bool = True):
        super().__init__()
        if pe_mode != "ape" or share_mod or no_temporal_attn or not qk_rms_norm or qk_rms_norm_cross:
            raise NotImplementedError(
                "the sm_100a engine implements the shipped configuration (configs/diffusion.yml): "
                "pe_mode='ape', share_mod=False, no_temporal_attn=False, qk_rms_norm=True, qk_rms_norm_cross=False")
        self.resolution, self.in_channels, self.model_channels = resolution, in_channels, model_channels
        self.out_channels, self.num_blocks = out_channels, num_blocks
        self.num_heads = num_heads or model_channels // num_head_channels
        self.mlp_ratio, self.pe_mode, self.use_fp16 = mlp_ratio, pe_mode, use_fp16
        self.dtype = torch.float16 if use_fp16 else torch.float32
        C = model_channels
        self.t_embedder = _TimestepEmbedderParams(C)
        self.input_layer = nn.Linear(in_channels, C)
        self.blocks = nn.ModuleList([_BlockParams(C, self.num_heads, mlp_ratio) for _ in range(num_blocks)])
        self.final_layer = _FinalLayerParams(C, out_channels)
        self.static_cond_proj = nn.Linear(static_cond_channels, C)
        self.image_cond_proj = nn.Linear(image_cond_channels, C)
        self.initialize_weights()
        self._engine, self._engine_sig = None, None
        self._cond_cache = {}
        self._cond_gen = 0

    @property
    def device(self):
        return next(self.parameters()).device

    def initialize_weights(self):
        # reference model/dit.py:401-427
        def _basic_init(m):
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
        self.apply(_basic_init)
        nn.init.normal_(self.t_embedder.mlp[0].weight, std=0.02)
        nn.init.normal_(self.t_embedder.mlp[2].weight, std=0.02)
        nn.init.normal_(self.static_cond_proj.weight, std=0.02)
        nn.init.normal_(self.image_cond_proj.weight, std=0.02)
        for blk in self.blocks:
            nn.init.constant_(blk.adaLN_modulation[-1].weight, 0)
            nn.init.constant_(blk.adaLN_modulation[-1].bias, 0)
        nn.init.constant_(self.final_layer.adaLN_modulation[-1].weight, 0)
        nn.init.constant_(self.final_layer.adaLN_modulation[-1].bias, 0)
        nn.init.constant_(self.final_layer.linear.weight, 0)
        nn.init.constant_(self.final_layer.linear.bias, 0)

    # ------------------------------------------------------------------ engine plumbing
    def engine(self) -> DiTEngine:
        sig = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._engine is None or sig != self._engine_sig:
            dev = self.device
            if dev.type != "cuda":
                raise RuntimeError("gvfdiffusion_b200.DiT runs on a CUDA device only (no CPU fallback)")
            self._engine = DiTEngine(self.state_dict(), self.num_heads, device=dev)
            self._engine_sig = sig
            self._cond_cache.clear()
        return self._engine

    _MAX_COND_ENTRIES = 24

    def _cached(self, kind, tensor, fn):
        """Per-conditioning-tensor projections, LRU over engine buffer slots.  An entry keeps its source tensor
        alive (so data_ptr stays unique) and owns one persistent engine slot of its kind; eviction hands the slot
        to the next miss.  Entries touched by the call in flight (same `_cond_gen`) are never evicted: their
        buffers are still referenced by the K/V lists being assembled."""
        key = (kind, tensor.data_ptr(), tuple(tensor.shape), tensor._version, str(tensor.device))
        hit = self._cond_cache.get(key)
        if hit is None:
            slot = None
            if len(self._cond_cache) >= self._MAX_COND_ENTRIES:
                for k, h in sorted(self._cond_cache.items(), key=lambda kv: kv[1][3]):      # oldest first
                    if h[3] < self._cond_gen and k[0] == kind:
                        slot = h[2]
                        del self._cond_cache[k]
                        break
            if slot is None:
                used = {h[2] for k, h in self._cond_cache.items() if k[0] == kind}
                slot = next(i for i in range(len(used) + 1) if i not in used)
            hit = [fn(tensor, slot), tensor, slot, self._cond_gen]
            self._cond_cache[key] = hit
        hit[3] = self._cond_gen
        return hit[0]

    def reset_conditioning(self):
        """Forget cached per-object projections (their device buffers are reused by the next object)."""
        self._cond_cache.clear()

    def _cond_sets(self, eng, cond_images, static_latent, xyz):
        B = cond_images.shape[0]
        kv_img = [self._cached("img", cond_images[b], eng.image_kv) for b in range(B)]
        kv_st = [self._cached("st", static_latent[b], eng.static_kv) for b in range(B)]
        pos = [self._cached("pos", xyz[b], eng.pos_embed) for b in range(B)]
        return kv_img, kv_st, pos

    @torch.no_grad()
    def forward(self, x, t, cond_images, static_latent, deformation_position_xyz=None):
        """x (B,T,N,C) fp32, t (B,), cond_images (B,T,L,Ci), static_latent (B,Ls,Cs), xyz (B,N,3)
        -> (B,T,N,Cout) fp32 (reference model/dit.py:449-480 under fp16 autocast)."""
        assert deformation_position_xyz is not None, "Deformation position xyz is required for APE mode"
        eng = self.engine()
        dev = eng.dev
        self._cond_gen += 1
        kv_img, kv_st, pos = self._cond_sets(eng, cond_images, static_latent, deformation_position_xyz)
        xt = x.to(dev, torch.float32).contiguous()
        tt = t.to(dev, torch.float32).reshape(-1).contiguous()
        if tt.numel() == 1 and xt.shape[0] > 1:
            tt = tt.expand(xt.shape[0]).contiguous()
        return eng.forward(xt, tt, kv_img, kv_st, pos).clone()

    @torch.no_grad()
    def prepare_conditioning(self, conds):
        """Hoist the per-object projections of every guidance branch now (normally the first NFE does it) and return a
        hashable key of the engine buffers they live in -- what a whole-run CUDA graph (GVFPipeline.sample) is keyed on."""
        eng = self.engine()
        self._cond_gen += 1
        key = []
        for c in conds:
            a, b, p = self._cond_sets(eng, c["cond_images"], c["static_latent"], c["deformation_position_xyz"])
            key.append((tuple(t.data_ptr() for e in a for t in e), tuple(t.data_ptr() for e in b for t in e),
                        tuple(t.data_ptr() for t in p)))
        return tuple(key), eng.mod_epoch

    def precompute_modulation(self, t_inputs):
        """Hook of DPM_Solver.sample: the model times of the whole run, known before the first NFE."""
        self.engine().precompute_modulation(t_inputs)

    @torch.no_grad()
    def forward_branches(self, x, t_input, conds):
        """All guidance branches of one NFE in a single engine pass (used by DPM_Solver):
        x (B,...) is shared, conds = list of condition dicts -> (len(conds)*B, T, N, Cout)."""
        eng = self.engine()
        dev = eng.dev
        self._cond_gen += 1
        B = x.shape[0]
        kv_img, kv_st, pos = [], [], []
        for c in conds:
            a, b, p = self._cond_sets(eng, c["cond_images"], c["static_latent"], c["deformation_position_xyz"])
            kv_img += a
            kv_st += b
            pos += p
        nb = len(conds)
        xin = x.to(dev, torch.float32).contiguous()
        if nb > 1:
            xin = xin.repeat(nb, *([1] * (x.dim() - 1)))
        return eng.forward_graphed(xin, float(t_input), kv_img, kv_st, pos)
