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

from .. import ops
from .attention import sparse_windowed_scaled_dot_product_self_attention

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
        if use_old_attn_impl:
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
        if self.C // num_heads != 64:
            raise NotImplementedError("windowed sparse attention is built for head dim 64 (768 / 12 on the shipped config)")

    def forward(self, feats, coords):
        """feats [T, C] fp32 CUDA, coords [T, 4] int32 CUDA (batch, x, y, z) -> [T, C] fp32 (fp16 with fp16_residual)."""
        if not (feats.is_cuda and coords.is_cuda):
            raise ValueError("SparseTransformerBlocks runs on CUDA tensors only (no CPU fallback)")
        T, C, H = feats.shape[0], self.C, self.H
        X = feats.to(F16 if self.fp16_residual else F32).contiguous().clone()
        epi = ops.EPI_RESID_F16 if self.fp16_residual else ops.EPI_RESID_F32
        A16 = torch.empty((T, C), dtype=F16, device=self.dev)
        QKV = torch.empty((T, 3 * C), dtype=F16, device=self.dev)
        H1 = torch.empty((T, self.blocks[0]["w1"].shape[0]), dtype=F16, device=self.dev)
        coords = coords.int().contiguous()
        for i, blk in enumerate(self.blocks):
            shift = (self.window // 2 * (i % 2),) * 3
            ops.ln_mod(X, out=A16)
            ops.gemm(A16, blk["w_qkv"], blk["b_qkv"], ops.EPI_F16, out=QKV)
            ao = sparse_windowed_scaled_dot_product_self_attention(QKV.view(T, 3, H, 64), coords, self.window, shift)
            ops.gemm(ao.view(T, C), blk["w_out"], blk["b_out"], epi, out=X)
            ops.ln_mod(X, out=A16)
            ops.gemm(A16, blk["w1"], blk["b1"], ops.EPI_GELU_F16, out=H1)
            ops.gemm(H1, blk["w2"], blk["b2"], epi, out=X)
        return X


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

    def decode(self, latent_feats, coords):
        """latent [T, latent_channels] fp32, coords [T, 4] int32 -> [T, out_channels] fp32 (:178-188)."""
        return self._trunk(self.decoder, "from_latent", "out_layer", latent_feats, coords)

    def encode(self, feats, coords):
        """feats [T, in_channels] -> (mean, logvar) [T, latent_channels] each (:151-176, sample_posterior=False)."""
        out = self._trunk(self.encoder, "input_layer", "to_latent", feats, coords)
        return out.chunk(2, dim=-1)
