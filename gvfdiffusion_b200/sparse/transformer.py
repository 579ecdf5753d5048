"""Device engine of a stack of un-modulated SparseTransformerBlock -- the encoder / decoder trunk of the static
SparseTransformerVAE (reference model/sparse_voxel_diffusion/sparse_transformer.py:126-192 with modulated=False;
stacked at sparse_transformer_vae.py:55-91; "swin" attention config at sparse_transformer.py:24-25: block i
attends inside 8^3 windows shifted by 4 * (i % 2)).

Host orchestration only; every tensor operation is a call into libgvf_b200.so: LayerNorm -> fp16, QKV GEMM,
windowed attention with the gather / scatter fused (csrc/sparse_attn.cu), out-projection + fp32 residual,
LayerNorm, fc1 + GELU, fc2 + residual -- seven launches per block.  Same precision regime as the DiT engine
(Linear and attention in fp16 with fp32 accumulation, LayerNorm and the residual stream in fp32).
"""
import torch

from .. import _lib, ops
from .._lib import check, current_stream, ptr
from .attention import sparse_windowed_scaled_dot_product_self_attention
from .attention.windowed_attn import windowed_attention_bwd, windowed_attention_fwd_lse

F16, F32 = torch.float16, torch.float32


def qkv_rows_from_old_attn_impl(weight, bias, num_heads):
    """`use_old_attn_impl=True` (the class default of SparseTransformerVAE; the shipped configs/vae.yml:30 and
    configs/diffusion.yml:57 set it to false) lays the to_qkv output channels out as [H][3][d]
    (sparse/attention/modules.py:161-164: reshape to (H, 3 d), chunk 3) instead of [3][H][d].  The kernels consume
    [3][H][d]; permuting the ROWS of to_qkv once at load time gives exactly the same q, k, v."""
    c3 = weight.shape[0]
    d = c3 // (3 * num_heads)
    idx = torch.arange(c3, device=weight.device).reshape(num_heads, 3, d).permute(1, 0, 2).reshape(-1)
    return weight[idx], bias[idx]


class SparseTransformerBlocks:
    def __init__(self, state_dict, prefix, num_blocks, num_heads, window_size, device="cuda", fp16_residual=False,
                 use_old_attn_impl=False):
        """fp16_residual: the residual stream is fp16 (`h.type(self.dtype)` with use_fp16=True,
        sparse_transformer_vae.py:155,181) instead of fp32.  use_old_attn_impl: see qkv_rows_from_old_attn_impl."""
        self.prefix, self.qkv_perm = prefix, None
        if use_old_attn_impl:
            c3 = state_dict[f"{prefix}0.attn.to_qkv.weight"].shape[0]
            self.qkv_perm = torch.arange(c3).reshape(num_heads, 3, c3 // (3 * num_heads)).permute(1, 0, 2).reshape(-1)
            state_dict = dict(state_dict)
            for i in range(num_blocks):
                kw, kb = f"{prefix}{i}.attn.to_qkv.weight", f"{prefix}{i}.attn.to_qkv.bias"
                state_dict[kw], state_dict[kb] = qkv_rows_from_old_attn_impl(state_dict[kw], state_dict[kb], num_heads)
        dev = torch.device(device)
        self.fp16_residual = fp16_residual
        h = lambda t: t.detach().to(device=dev, dtype=F16).contiguous()
        b = lambda t: t.detach().to(device=dev, dtype=F16).to(F32).contiguous()     # fp16-valued fp32 bias
        self.dev, self.H, self.window = dev, num_heads, window_size
        self.blocks = []
        for i in range(num_blocks):
            p = f"{prefix}{i}."
            self.blocks.append(dict(
                w_qkv=h(state_dict[p + "attn.to_qkv.weight"]), b_qkv=b(state_dict[p + "attn.to_qkv.bias"]),
                w_out=h(state_dict[p + "attn.to_out.weight"]), b_out=b(state_dict[p + "attn.to_out.bias"]),
                w1=h(state_dict[p + "mlp.mlp.0.weight"]), b1=b(state_dict[p + "mlp.mlp.0.bias"]),
                w2=h(state_dict[p + "mlp.mlp.2.weight"]), b2=b(state_dict[p + "mlp.mlp.2.bias"])))
        self.C = self.blocks[0]["w_out"].shape[0]
        self.F = self.blocks[0]["w1"].shape[0]
        self._scr = None
        if self.C // num_heads != 64:
            raise NotImplementedError("windowed sparse attention is built for head dim 64 (768 / 12 on the shipped config)")

    def refresh(self, state_dict):
        """New parameter values into the SAME device buffers (after an optimiser step): one foreach cast for the fp16
        weights, one for the fp16-valued fp32 biases, and the transposes of the backward if they exist."""
        dst_w, src_w, dst_b, src_b = [], [], [], []
        for i, blk in enumerate(self.blocks):
            p = f"{self.prefix}{i}."
            for n, key in (("w_qkv", "attn.to_qkv"), ("w_out", "attn.to_out"), ("w1", "mlp.mlp.0"), ("w2", "mlp.mlp.2")):
                w, b = state_dict[p + key + ".weight"], state_dict[p + key + ".bias"]
                if n == "w_qkv" and self.qkv_perm is not None:
                    perm = self.qkv_perm.to(w.device)
                    w, b = w[perm], b[perm]
                dst_w.append(blk[n]); src_w.append(w)
                dst_b.append(blk["b" + n[1:]]); src_b.append(b)
        with torch.no_grad():
            torch._foreach_copy_(dst_w, src_w)
            torch._foreach_copy_(dst_b, [b.to(F16) for b in src_b])

    # ---------------------------------------------------------------------------------------- native driver
    def _parts(self, coords):
        """gvf_window_partition[2]: the un-shifted and the shifted window partition of `coords` (cached per tensor)."""
        from .attention.windowed_attn import _partition, _seq_of_pos
        arr = (_lib.WindowPartition * 2)()
        keep = []
        for k, sh in enumerate((0, self.window // 2)):
            fwd, _bwd, cu, max_len = _partition(coords, self.window, (sh,) * 3)
            # short windows (object surfaces: ~16 voxels each): the packed tiling, several windows per 64-row tile
            sop = _seq_of_pos(coords, self.window, (sh,) * 3) if max_len < 256 else None
            arr[k] = _lib.WindowPartition(ptr(fwd), ptr(cu), cu.shape[0] - 1, max_len, ptr(sop))
            keep += [fwd, cu, sop]
        return arr, keep

    def _block_structs(self, grads=None):
        """gvf_sparse_block[num_blocks] over the engine's weights (+ transposes and gradient views for the backward)."""
        arr = (_lib.SparseBlock * len(self.blocks))()
        for i, blk in enumerate(self.blocks):
            vals = {n: ptr(blk[n]) for n in ("w_qkv", "w_out", "w1", "w2", "b_qkv", "b_out", "b1", "b2")}
            if grads is not None:
                vals.update({"g_" + n: ptr(t) for n, t in grads[i].items()})
            arr[i] = _lib.SparseBlock(**vals)
        return arr

    def _scratch(self, T):
        need = _lib.lib().gvf_sparse_trunk_scratch_bytes(T, self.C, self.H, self.F)
        if self._scr is None or self._scr.numel() < need:
            self._scr = torch.empty(need, dtype=torch.uint8, device=self.dev)
        return self._scr

    def forward(self, feats, coords):
        """feats [T, C] fp32 CUDA, coords [T, 4] int32 CUDA (batch, x, y, z) -> [T, C] fp32 (fp16 with fp16_residual).
        One call into the native trunk driver (csrc/sparse_trunk.cu: seven launches per block)."""
        if not (feats.is_cuda and coords.is_cuda):
            raise ValueError("SparseTransformerBlocks runs on CUDA tensors only (no CPU fallback)")
        T = feats.shape[0]
        X = feats.to(F16 if self.fp16_residual else F32).contiguous().clone()
        coords = coords.int().contiguous()
        parts, _keep = self._parts(coords)
        scr = self._scratch(T)
        check(_lib.lib().gvf_sparse_trunk_forward(self._block_structs(), len(self.blocks), T, self.C, self.H, self.F,
                                                  int(self.fp16_residual), parts, ptr(X), None, 0, ptr(scr), scr.numel(), ptr(X),
                                                  current_stream()), "gvf_sparse_trunk_forward")
        return X

    # ---------------------------------------------------------------------------------------- training (cfg 5)
    def forward_train(self, feats, coords):
        """Same blocks, keeping what the backward needs (block inputs, LayerNorm outputs, QKV, attention output + LSE, the
        MLP pre-activation and activation) in one arena (gvf_sparse_trunk_arena_bytes: ~25 KB per token and block).  GELU
        stays in the fc1 epilogue, which also stores the fp16 pre-activation (autocast's rounding order).
        -> (X [T, C], saved)"""
        if not (feats.is_cuda and coords.is_cuda):
            raise ValueError("SparseTransformerBlocks runs on CUDA tensors only (no CPU fallback)")
        T, nb = feats.shape[0], len(self.blocks)
        x_in = feats.detach().to(F16 if self.fp16_residual else F32).contiguous()
        coords = coords.int().contiguous()
        parts, keep = self._parts(coords)
        nbytes = _lib.lib().gvf_sparse_trunk_arena_bytes(T, self.C, self.H, self.F, nb, int(self.fp16_residual))
        arena = torch.empty(nbytes, dtype=torch.uint8, device=self.dev)
        X = torch.empty_like(x_in)
        check(_lib.lib().gvf_sparse_trunk_forward(self._block_structs(), nb, T, self.C, self.H, self.F, int(self.fp16_residual),
                                                  parts, ptr(x_in), ptr(arena), nbytes, None, 0, ptr(X), current_stream()),
              "gvf_sparse_trunk_forward")
        return X, {"coords": coords, "arena": arena, "parts": parts, "keep": keep, "T": T}

    def backward(self, saved, dX):
        """dX [T, C] (gradient of forward_train's output) -> ({reference parameter name: fp32 gradient}, d feats fp16).
        Per block, in reverse (csrc/sparse_trunk.cu): fc2 dgrad with GELU' in its epilogue / wgrad -> fc1 dgrad / wgrad ->
        LayerNorm backward (+ residual) -> to_out dgrad / wgrad -> window attention backward -> to_qkv dgrad / wgrad ->
        LayerNorm backward (+ residual); the dgrads read the [out, in] weights directly (no transposed copies); bias
        gradients = one-launch column sums.  Activation gradients are fp16 (fp32
        accumulation inside every kernel), parameter gradients fp32 views of one flat buffer."""
        T, C, F_, nb = saved["T"], self.C, self.F, len(self.blocks)
        dx = dX.detach().to(F16).contiguous()
        shapes = (("w_qkv", (3 * C, C)), ("b_qkv", (3 * C,)), ("w_out", (C, C)), ("b_out", (C,)), ("w1", (F_, C)), ("b1", (F_,)),
                  ("w2", (C, F_)), ("b2", (C,)))
        per = sum(int(torch.Size(s).numel()) for _, s in shapes)
        flat = torch.empty(nb * per, dtype=F32, device=self.dev)
        grads, off = [], 0
        for _ in range(nb):
            d = {}
            for n, s in shapes:
                k = int(torch.Size(s).numel())
                d[n] = flat[off:off + k].view(s)
                off += k
            grads.append(d)
        scr = self._scratch(T)
        ws = ops._reduce_ws(self.dev, _lib.lib().gvf_colsum_workspace_bytes(T, max(3 * C, F_), 0))
        d_in = torch.empty((T, C), dtype=F16, device=self.dev)
        arena = saved["arena"]
        check(_lib.lib().gvf_sparse_trunk_backward(self._block_structs(grads), nb, T, C, self.H, F_, int(self.fp16_residual),
                                                   saved["parts"], ptr(arena), arena.numel(), ptr(dx), ptr(scr), scr.numel(), ptr(ws),
                                                   ws.numel() * 4, ptr(d_in), current_stream()), "gvf_sparse_trunk_backward")
        names = {"w_qkv": "attn.to_qkv.weight", "b_qkv": "attn.to_qkv.bias", "w_out": "attn.to_out.weight", "b_out": "attn.to_out.bias",
                 "w1": "mlp.mlp.0.weight", "b1": "mlp.mlp.0.bias", "w2": "mlp.mlp.2.weight", "b2": "mlp.mlp.2.bias"}
        g = {}
        for i, d in enumerate(grads):
            if self.qkv_perm is not None:                 # rows were permuted at load time: hand the gradient back in
                inv = torch.empty_like(self.qkv_perm)     # the checkpoint's [H][3][d] order
                inv[self.qkv_perm] = torch.arange(inv.numel())
                inv = inv.to(self.dev)
                d["w_qkv"], d["b_qkv"] = d["w_qkv"][inv], d["b_qkv"][inv]
            for n, t in d.items():
                g[f"{self.prefix}{i}.{names[n]}"] = t
        return g, d_in


class SparseTransformerVAE:
    """encode / decode trunks of the static SparseTransformerVAE (reference
    model/sparse_voxel_diffusion/sparse_transformer_vae.py:151-188) over its state dict: input_layer /
    from_latent + AbsolutePositionEmbedder -> swin blocks (fp16 residual stream when use_fp16) -> optional
    affine-free LayerNorm (norm_output, eps 1e-5) -> to_latent / out_layer.  `to_representation`
    (sparse_vae.py:114-180) stays host-side torch (gvfdiffusion_b200/synthetic.py builds the same tensors)."""

    def __init__(self, state_dict, num_blocks, num_heads, window_size=8, use_fp16=True, norm_output=False, device="cuda",
                 use_old_attn_impl=False):
        dev = torch.device(device)
        self.dev, self.norm_output, self.use_fp16 = dev, norm_output, use_fp16
        h = lambda t: t.detach().to(device=dev, dtype=F16).contiguous()
        b = lambda t: t.detach().to(device=dev, dtype=F16).to(F32).contiguous()
        sd = state_dict
        self.lin = {n: (h(sd[n + ".weight"]), b(sd[n + ".bias"])) for n in ("input_layer", "to_latent", "from_latent", "out_layer")
                    if n + ".weight" in sd}
        self.C = sd["from_latent.weight"].shape[0] if "from_latent.weight" in sd else sd["input_layer.weight"].shape[0]
        mk = lambda prefix: SparseTransformerBlocks(sd, prefix, num_blocks, num_heads, window_size, dev, fp16_residual=use_fp16,
                                                    use_old_attn_impl=use_old_attn_impl)
        self.encoder = mk("encoder.") if "encoder.0.attn.to_qkv.weight" in sd else None
        self.decoder = mk("decoder.") if "decoder.0.attn.to_qkv.weight" in sd else None

    def refresh(self, state_dict):
        """In-place update of every device copy from new parameter values (see SparseTransformerBlocks.refresh)."""
        with torch.no_grad():
            for n in ("input_layer", "to_latent", "from_latent", "out_layer"):
                if n in self.lin:
                    w, b = self.lin[n]
                    w.copy_(state_dict[n + ".weight"])
                    b.copy_(state_dict[n + ".bias"].to(F16))
                    if n + "_t" in self.lin:
                        ops.transpose(w, out=self.lin[n + "_t"])
        for blocks in (self.encoder, self.decoder):
            if blocks is not None:
                blocks.refresh(state_dict)

    def _linear(self, name, x, add=None):
        """nn.Linear under autocast: fp16 operands, fp16-rounded result returned as fp32 (+ fp32 `add` rows)."""
        w, bias = self.lin[name]
        M, K = x.shape
        if K <= 32:
            return ops.small_linear(x.to(F32).contiguous(), w, bias, out_f16=False, add=add, add_rows=M if add is not None else 0)
        y = ops.gemm(ops.cast_f16(x.to(F32)) if x.dtype != F16 else x.contiguous(), w, bias, ops.EPI_F16).float()
        return y if add is None else y + add

    def _trunk(self, blocks, first, last, feats, coords):
        coords = coords.int().contiguous()
        pos = ops.ape(coords[:, 1:].float().contiguous(), self.C)
        hcur = self._linear(first, feats, add=pos)
        hcur = blocks.forward(hcur, coords).float()
        if self.norm_output:
            hcur = ops.ln_mod(hcur.contiguous(), eps=1e-5).float()
        return self._linear(last, hcur)

    # ---------------------------------------------------------------------------------------- training (cfg 5)
    def _trunk_train(self, blocks, first, feats, coords):
        coords = coords.int().contiguous()
        feats = feats.detach().to(self.dev, F32).contiguous()
        pos = ops.ape(coords[:, 1:].float().contiguous(), self.C)
        w, bias = self.lin[first]
        if w.shape[1] <= 16:
            h0 = ops.small_linear(feats, w, bias, out_f16=False, add=pos, add_rows=feats.shape[0])
        else:
            feats = ops.cast_f16(feats)
            h0 = ops.gemm(feats, w, bias, ops.EPI_F16).float() + pos
        X, saved = blocks.forward_train(h0, coords)
        saved["first_in"] = feats
        return X, saved

    def _first_backward(self, first, saved, dx, g, need_input_grad):
        """gradients of the trunk's first Linear (from_latent: K <= 32 on the skinny kernels; input_layer: GEMMs)."""
        w, _ = self.lin[first]
        x = saved["first_in"]
        g[first + ".bias"] = ops.colsum(dx)
        if w.shape[1] <= 16:
            g[first + ".weight"] = ops.skinny_outer(x, dx).t().contiguous()
            return ops.small_linear_bwd_input(dx, w) if need_input_grad else None
        g[first + ".weight"] = ops.gemm_tn(dx, x)
        if not need_input_grad:
            return None
        return ops.gemm_nn(dx, w).float()

    def _last_backward(self, last, saved, dout, g):
        """gradients of the trunk's last Linear (+ the affine-free LayerNorm in front of it) -> d block output fp16."""
        w, _ = self.lin[last]
        N, T = w.shape[0], dout.shape[0]
        N8 = (N + 7) // 8 * 8
        d16 = torch.zeros((T, N8), dtype=F16, device=self.dev)
        d16[:, :N] = dout.detach()
        g[last + ".weight"] = ops.gemm_tn(d16, saved["hn"])[:N]
        g[last + ".bias"] = ops.colsum(d16)[:N]
        if N == N8:
            dh = ops.gemm_nn(d16, w)                                          # d h = d out W, W read as it lies in memory
        else:                                                                 # odd widths: a zero-padded transposed copy
            if last + "_t" not in self.lin:
                self.lin[last + "_t"] = ops.transpose(w)                      # [C, N8]
            dh = ops.gemm(d16, self.lin[last + "_t"], None, ops.EPI_F16)
        if self.norm_output:
            dh = ops.ln_bwd(saved["x_last"], dh, None, eps=1e-5)
        return dh

    def _last_forward(self, last, X, saved):
        saved["x_last"] = X
        hn = ops.ln_mod(X, eps=1e-5) if self.norm_output else (X if X.dtype == F16 else ops.cast_f16(X))
        saved["hn"] = hn
        w, bias = self.lin[last]
        return ops.gemm(hn, w, bias, ops.EPI_F16).float()

    def encode_train(self, feats, coords, noise=None, sample_posterior=True):
        """encode keeping the activations (:151-176): -> (z, mean, logvar, kl, saved); kl = 0.5 mean(mean^2 + var - logvar
        - 1) over all voxels (sparse_vae.py:351).  noise: the randn_like draw of :165 ([T, latent]; host RNG when None)."""
        X, saved = self._trunk_train(self.encoder, "input_layer", feats, coords)
        ml = self._last_forward("to_latent", X, saved)
        lat = ml.shape[1] // 2
        mean, logvar = ml[:, :lat].contiguous(), ml[:, lat:].contiguous()
        if sample_posterior and noise is None:
            noise = torch.randn(mean.shape)
        noise = noise.to(self.dev, F32).contiguous() if sample_posterior else None
        z, kl = torch.empty_like(mean), torch.empty(1, dtype=F32, device=self.dev)
        check(_lib.lib().gvf_diag_gaussian(ptr(mean), ptr(logvar), ptr(noise), 1, mean.numel(), ptr(z), ptr(kl), current_stream()),
              "gvf_diag_gaussian")
        saved.update(mean=mean, logvar=logvar, noise=noise)
        return z, mean, logvar, kl[0], saved

    def encode_backward(self, saved, dz=None, dkl=None, need_input_grad=False):
        """-> ({parameter name: fp32 gradient}, d feats or None)."""
        mean, logvar = saved["mean"], saved["logvar"]
        dmean, dlogvar = torch.empty_like(mean), torch.empty_like(mean)
        dzc = None if dz is None else dz.detach().to(self.dev, F32).contiguous()
        dkc = None if dkl is None else dkl.detach().to(self.dev, F32).reshape(1).contiguous()
        check(_lib.lib().gvf_diag_gaussian_bwd(ptr(mean), ptr(logvar), ptr(saved["noise"]), ptr(dzc), ptr(dkc), 1, mean.numel(),
                                               ptr(dmean), ptr(dlogvar), current_stream()), "gvf_diag_gaussian_bwd")
        g = {}
        dh = self._last_backward("to_latent", saved, torch.cat([dmean, dlogvar], 1), g)
        gb, dx = self.encoder.backward(saved, dh)
        g.update(gb)
        return g, self._first_backward("input_layer", saved, dx, g, need_input_grad)

    def forward_train(self, feats, coords, noise=None):
        """SparseTransformerVAE.forward with sample_posterior=True (:206-210): -> (out, mean, logvar, kl, saved)."""
        z, mean, logvar, kl, se = self.encode_train(feats, coords, noise)
        out, sd_ = self.decode_train(z, coords)
        return out, mean, logvar, kl, {"enc": se, "dec": sd_}

    def backward(self, saved, dout, dkl=None):
        """gradients of every parameter of the VAE from d out [T, out_channels] and d kl (scalar tensor)."""
        g, dz = self.decode_backward(saved["dec"], dout)
        ge, _ = self.encode_backward(saved["enc"], dz, dkl)
        g.update(ge)
        return g

    def decode_train(self, latent_feats, coords):
        """decode keeping the activations: -> (out [T, out_channels] fp32, saved)."""
        X, saved = self._trunk_train(self.decoder, "from_latent", latent_feats, coords)
        return self._last_forward("out_layer", X, saved), saved

    def decode_backward(self, saved, dout):
        """dout [T, out_channels] -> ({parameter name: fp32 gradient}, d latent fp32 [T, latent_channels])."""
        g = {}
        dh = self._last_backward("out_layer", saved, dout, g)
        gb, dx = self.decoder.backward(saved, dh)
        g.update(gb)
        return g, self._first_backward("from_latent", saved, dx, g, True)

    def decode(self, latent_feats, coords):
        """latent [T, latent_channels] fp32, coords [T, 4] int32 -> [T, out_channels] fp32 (:178-188)."""
        return self._trunk(self.decoder, "from_latent", "out_layer", latent_feats, coords)

    def encode(self, feats, coords):
        """feats [T, in_channels] -> (mean, logvar) [T, latent_channels] each (:151-176, sample_posterior=False)."""
        out = self._trunk(self.encoder, "input_layer", "to_latent", feats, coords)
        return out.chunk(2, dim=-1)


class _SparseVAETrainFn(torch.autograd.Function):
    """The whole static VAE (encode -> posterior sample -> decode) as one autograd node: forward_train / backward of the
    engine above.  Parameter gradients are left on `engine.grads` ({reference parameter name: fp32 tensor}); `anchor` is any
    tensor that requires grad, so that autograd calls the node."""

    @staticmethod
    def forward(ctx, engine, feats, coords, noise, anchor):
        out, mean, logvar, kl, saved = engine.forward_train(feats, coords, noise)
        ctx.engine, ctx.saved = engine, saved
        ctx.mark_non_differentiable(mean, logvar)
        return out, kl.reshape(()), mean, logvar

    @staticmethod
    def backward(ctx, dout, dkl, _dm, _dl):
        ctx.engine.grads = ctx.engine.backward(ctx.saved, dout, dkl)
        ctx.saved = None
        return None, None, None, None, None


def sparse_vae_forward_autograd(engine, feats, coords, noise=None, anchor=None):
    """(out [T, out_channels], kl, mean, logvar) with `out` and `kl` attached to the autograd graph (what
    `SparseVAE.training_losses`, reference sparse_vae.py:318, consumes)."""
    if anchor is None:
        anchor = torch.zeros((), device=engine.dev, requires_grad=True)
    return _SparseVAETrainFn.apply(engine, feats, coords, noise, anchor)
