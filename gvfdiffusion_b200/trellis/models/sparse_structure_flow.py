"""`SparseStructureFlowModel`: the dense DiT over the 16^3 occupancy latent, first denoiser of the TRELLIS stage (reference
trellis/models/sparse_structure_flow.py:55-200; sampled by trellis/pipelines/trellis_image_to_3d.py:165-190, its decoded
occupancy gives the `coords` of the structured-latent stage).  Same constructor arguments, state-dict keys and
`forward(x [B, C, R, R, R], t [B], cond [B, L, Cc])` as the reference class.  Its blocks are the arithmetic of the
structured-latent flow model's ModulatedSparseTransformerCrossBlocks on one full sequence per batch entry
(trellis/modules/transformer/modulated.py:132-150), so the engine is the same: one GEMV for every adaLN vector, fp32 residual
stream, tcgen05 GEMMs with fused epilogues, dense tcgen05 flash attention for self- and cross-attention, image-token K / V
projected once per conditioning tensor, optional CUDA-graph replay.  patchify / unpatchify are index permutations (torch
views + one copy).  Inference only, CUDA only."""
import math

import torch

from ... import ops
from .structured_latent_flow import F16, F32, _CrossBlock, _f, _h


class SparseStructureFlowModel:
    def __init__(self, resolution, in_channels, model_channels, cond_channels, out_channels, num_blocks, num_heads=None,
                 num_head_channels=64, mlp_ratio=4, patch_size=2, pe_mode="ape", use_fp16=False, use_checkpoint=False,
                 share_mod=False, qk_rms_norm=False, qk_rms_norm_cross=False, device="cuda"):
        if pe_mode != "ape" or share_mod:
            raise NotImplementedError("rope / share_mod are not used by the shipped sparse-structure flow checkpoints")
        self.resolution, self.in_channels, self.model_channels = resolution, in_channels, model_channels
        self.cond_channels, self.out_channels, self.num_blocks = cond_channels, out_channels, num_blocks
        self.num_heads = num_heads or model_channels // num_head_channels
        self.mlp_ratio, self.patch_size = mlp_ratio, patch_size
        self.qk_rms_norm, self.qk_rms_norm_cross = qk_rms_norm, qk_rms_norm_cross
        self.dtype, self.device = F16, torch.device(device)
        if model_channels > 1024 or model_channels % 64 or (model_channels // self.num_heads) not in (32, 64):
            raise ValueError("unsupported widths (model_channels <= 1024, head dim 32 / 64)")
        self.tokens = (resolution // patch_size) ** 3
        self.blocks, self._kv_cache, self._ws = [], {}, {}
        self._loaded, self.use_graphs = False, False

    def load_state_dict(self, sd, strict=True):
        dev, C = self.device, self.model_channels
        self.t_w0, self.t_b0 = _h(sd["t_embedder.mlp.0.weight"], dev), _f(sd["t_embedder.mlp.0.bias"], dev)
        self.t_w2, self.t_b2 = _h(sd["t_embedder.mlp.2.weight"], dev), _f(sd["t_embedder.mlp.2.bias"], dev)
        self.in_w, self.in_b = _h(sd["input_layer.weight"], dev), _f(sd["input_layer.bias"], dev)
        self.blocks = [_CrossBlock(sd, f"blocks.{i}.", C, self.num_heads, self.qk_rms_norm, self.qk_rms_norm_cross, dev)
                       for i in range(self.num_blocks)]
        for i, blk in enumerate(self.blocks):
            blk.mod_off = 6 * C * i
        self.mod_w = _h(torch.cat([b.mod_w for b in self.blocks], 0), dev)
        self.mod_b = _f(torch.cat([b.mod_b for b in self.blocks], 0), dev)
        self.R = 6 * C * self.num_blocks
        wo, bo = sd["out_layer.weight"].detach().float().to(dev), sd["out_layer.bias"].detach().float().to(dev)
        self.out_features = wo.shape[0]
        pad = (-self.out_features) % 8
        self.out_w = _h(torch.cat([wo, torch.zeros(pad, wo.shape[1], device=dev)], 0), dev)
        self.out_b = _f(torch.cat([bo, torch.zeros(pad, device=dev)], 0), dev)
        # the buffer `pos_emb` of the reference (:94-99): APE of the patch grid, recomputed here
        r = self.resolution // self.patch_size
        grid = torch.stack(torch.meshgrid(*[torch.arange(r, device=dev)] * 3, indexing="ij"), dim=-1).reshape(-1, 3)
        self.pos_emb = ops.ape(grid.to(F32).contiguous(), C)
        self._kv_cache.clear()
        self._loaded = True
        return self

    def _context_entry(self, cond):
        key = (cond.data_ptr(), cond._version, tuple(cond.shape), cond.dtype)
        hit = self._kv_cache.get(key)
        if hit is None:
            if len(self._kv_cache) >= 4:
                self._kv_cache.pop(next(iter(self._kv_cache)))
            c16 = cond.detach().to(self.device, F16).contiguous()
            hit = {"cond": cond, "kv": [blk.context_kv(c16) for blk in self.blocks], "graphs": {}}
            self._kv_cache[key] = hit
        return hit

    def reset_conditioning(self):
        self._kv_cache.clear()

    def _workspace(self, n):
        ws = self._ws.get(n)
        if ws is None:
            C, dev = self.model_channels, self.device
            ws = dict(A=torch.empty((n, C), dtype=F16, device=dev), QKV=torch.empty((n, 3 * C), dtype=F16, device=dev),
                      AO=torch.empty((n, C), dtype=F16, device=dev),
                      H1=torch.empty((n, int(C * self.mlp_ratio)), dtype=F16, device=dev))
            self._ws[n] = ws
        return ws

    @torch.no_grad()
    def forward(self, x, t, cond):
        if not self._loaded:
            raise RuntimeError("load_state_dict first")
        if not (x.is_cuda and cond.is_cuda):
            raise RuntimeError("SparseStructureFlowModel runs on CUDA tensors only (no CPU fallback)")
        B, C, ps, L, dev = x.shape[0], self.model_channels, self.patch_size, self.tokens, self.device
        assert list(x.shape) == [B, self.in_channels] + [self.resolution] * 3, f"Input shape mismatch, got {tuple(x.shape)}"
        if B > 8:
            raise ValueError("at most 8 batch entries per call")
        r = self.resolution // ps
        # patchify (:178-179): channel-major patch features, tokens in (x, y, z) order
        tok = x.to(F32).reshape(B, self.in_channels, r, ps, r, ps, r, ps).permute(0, 2, 4, 6, 1, 3, 5, 7)
        tok = tok.reshape(B * L, self.in_channels * ps ** 3).contiguous()
        tt = t.to(dev, F32).reshape(-1).contiguous()
        if tt.numel() == 1 and B > 1:
            tt = tt.expand(B).contiguous()
        temb = torch.empty((B, C), dtype=F16, device=dev)
        stemb = torch.empty((B, C), dtype=F16, device=dev)
        mod = torch.empty((B, self.R), dtype=F16, device=dev)
        ops.dit_modulation(tt, self.t_w0, self.t_b0, self.t_w2, self.t_b2, self.mod_w, self.mod_b, temb, stemb, mod)
        kvs = self._context_entry(cond)["kv"]
        if tok.shape[1] <= 32:                       # input_layer + pos_emb -> the fp32 residual stream (:180-181)
            X = ops.small_linear(tok, self.in_w, self.in_b, out_f16=False, add=self.pos_emb, add_rows=L)
        else:
            X = ops.gemm(ops.cast_f16(tok), self.in_w, self.in_b, ops.EPI_F32)
            X = X.view(B, L, C).add_(self.pos_emb[None]).view(B * L, C)
        layout = [slice(b * L, (b + 1) * L) for b in range(B)]
        ws = self._workspace(B * L)
        for blk, kv in zip(self.blocks, kvs):
            blk.forward(X, layout, mod, kv, ws)
        a = ops.ln_mod(X, eps=1e-5)                  # F.layer_norm (:194)
        out = torch.empty((B * L, self.out_features), dtype=F32, device=dev)
        ops.gemm(a, self.out_w, self.out_b, ops.EPI_F32_COMPACT, out=out)
        # unpatchify (:197-198)
        co = self.out_features // ps ** 3
        out = out.view(B, r, r, r, co, ps, ps, ps).permute(0, 4, 1, 5, 2, 6, 3, 7)
        return out.reshape(B, co, self.resolution, self.resolution, self.resolution).contiguous()

    @torch.no_grad()
    def forward_graphed(self, x, t, cond):
        """forward() replayed from a CUDA graph keyed on (input shape, conditioning entry)."""
        ent = self._context_entry(cond)
        B = x.shape[0]
        key = tuple(x.shape)
        tt = t.to(self.device, F32).reshape(-1)
        if tt.numel() == 1 and B > 1:
            tt = tt.expand(B)
        g = ent["graphs"].get(key)
        if g is None:
            xs, ts = x.to(F32).clone(), tt.clone()
            self.forward(xs, ts, cond)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self.forward(xs, ts, cond)
            g = (graph, xs, ts, out)
            ent["graphs"][key] = g
        graph, xs, ts, out = g
        xs.copy_(x)
        ts.copy_(tt)
        graph.replay()
        return out.clone()

    def __call__(self, x, t, cond):
        return self.forward_graphed(x, t, cond) if self.use_graphs else self.forward(x, t, cond)
