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


class SparseTransformerBlocks:
    def __init__(self, state_dict, prefix, num_blocks, num_heads, window_size, device="cuda"):
        dev = torch.device(device)
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
        """feats [T, C] fp32 CUDA, coords [T, 4] int32 CUDA (batch, x, y, z) -> [T, C] fp32."""
        if not (feats.is_cuda and coords.is_cuda):
            raise ValueError("SparseTransformerBlocks runs on CUDA tensors only (no CPU fallback)")
        T, C, H = feats.shape[0], self.C, self.H
        X = feats.to(F32).contiguous().clone()
        A16 = torch.empty((T, C), dtype=F16, device=self.dev)
        QKV = torch.empty((T, 3 * C), dtype=F16, device=self.dev)
        H1 = torch.empty((T, self.blocks[0]["w1"].shape[0]), dtype=F16, device=self.dev)
        coords = coords.int().contiguous()
        for i, blk in enumerate(self.blocks):
            shift = (self.window // 2 * (i % 2),) * 3
            ops.ln_mod(X, out=A16)
            ops.gemm(A16, blk["w_qkv"], blk["b_qkv"], ops.EPI_F16, out=QKV)
            ao = sparse_windowed_scaled_dot_product_self_attention(QKV.view(T, 3, H, 64), coords, self.window, shift)
            ops.gemm(ao.view(T, C), blk["w_out"], blk["b_out"], ops.EPI_RESID_F32, out=X)
            ops.ln_mod(X, out=A16)
            ops.gemm(A16, blk["w1"], blk["b1"], ops.EPI_GELU_F16, out=H1)
            ops.gemm(H1, blk["w2"], blk["b2"], ops.EPI_RESID_F32, out=X)
        return X
