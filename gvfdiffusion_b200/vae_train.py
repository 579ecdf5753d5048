"""Training-mode engine of the motion-VAE decoder: forward with saved activations + hand-written backward
(BASELINE configs[2] "VAE decode + gaussian_render 24f x 512^2, 16k Gaussians, fwd+bwd" and the decode part of
configs[4]; reference train_vae.py:293-353 -> model/autoencoder.py:552-609 under torch autograd + fp16 autocast).

Host orchestration only; every contraction / reduction is a libgvf_b200.so kernel:
  * Linear dgrad  dX = dY W          -> gvf_gemm_f16(A = dY, "W" = W^T)            (W^T made once per weight version)
  * Linear wgrad  dW = dY^T X (fp32) -> gvf_gemm_tn_f16: both operands MN-major from 64 x 64 TMA boxes, split-K with
    TMA reduce-add (no transposed activation copies)
  * bias grads gvf_colsum; LayerNorm / GEGLU / query-embedding backward, K <= 16 Linears: csrc/backward.cu
  * the decoder's output side (to_out followed by to_outputs, no non-linearity) composes into rank-14 products
  * attention forward with LSE + backward (dQ, dK, dV): csrc/attn.cu / csrc/attn_bwd.cu
Activation gradients are fp16 (as under the reference's autocast), parameter gradients fp32.  The decoder queries are
not chunked here (the reference chunks at 8192 queries with gradient checkpointing to save memory,
model/autoencoder.py:591-607; the result is the same and 180 GB of HBM hold the 16k-query activations: 1.3 GB).
"""
import torch

from . import ops
from .vae_engine import VAEDecodeEngine

F16, F32 = torch.float16, torch.float32


class VAEDecodeTrainEngine(VAEDecodeEngine):
    def __init__(self, state_dict, heads, num_timesteps, device="cuda"):
        super().__init__(state_dict, heads, num_timesteps, device, chunk_size=1 << 30)
        # no transposed weight copies: the dgrad GEMMs read the [out, in] weights as MN-major B operands (ops.gemm_nn)
        self.w_o_t = self.w_o[:self.out_dim].t().contiguous()            # [dim, out_dim]: d lat = d out @ w_o

    def refresh(self, sd):
        super().refresh(sd)
        with torch.no_grad():
            self.w_o_t.copy_(self.w_o[:self.out_dim].t())

    # ------------------------------------------------------------------------------------------------ forward
    def forward_train(self, z, queries):
        """z [(B*T), L, latent] fp32, queries [B, Q, 14] fp32 -> (out [B, T, Q, out_dim] fp32, saved activations)."""
        dev = self.dev
        z = z.detach().to(dev, F32).contiguous()
        queries = queries.detach().to(dev, F32).contiguous()
        B, Q, _ = queries.shape
        T, H, d, dim = self.T, self.H, self.d, self.dim
        BT, L, Cl = z.shape
        assert BT == B * T
        M = BT * L
        scale = d ** -0.5
        sv = {"z": z.reshape(M, Cl), "queries": queries, "layers": [], "shape": (B, Q, BT, L)}
        x = ops.small_linear(sv["z"], self.w_proj, self.b_proj, out_f16=True)
        for ly in self.layers:
            A = ops.ln_mod(x, eps=1e-6)
            QKV = ops.gemm(A, ly["w_qkv"], None, ops.EPI_F16)
            q5 = QKV.view(BT, L, 3, H, d)
            AO, lse = ops.attention_fwd_lse(q5[:, :, 0], q5[:, :, 1], q5[:, :, 2], scale)
            x1 = x.clone()
            ops.gemm(AO.view(M, dim), ly["w_out"], ly["b_out"], ops.EPI_RESID_F16, out=x1)
            A2 = ops.ln_mod(x1, eps=1e-6)
            Hf = ops.gemm(A2, ly["w1"], ly["b1"], ops.EPI_F16)
            G = ops.geglu(Hf)
            x2 = x1.clone()
            ops.gemm(G, ly["w2"], ly["b2"], ops.EPI_RESID_F16, out=x2)
            sv["layers"].append(dict(x0=x, A=A, QKV=QKV, AO=AO, lse=lse, x1=x1, A2=A2, Hf=Hf, G=G))
            x = x2
        ctx = ops.ln_mod(x, eps=1e-6)
        KV = ops.gemm(ctx, self.w_dkv, None, ops.EPI_F16)
        kv4 = KV.view(B, T, L, 2, H, d)
        out = torch.empty((B, T, Q, self.out_dim), dtype=F32, device=dev)
        # decoder side, all batch entries stacked (only the attention itself is per entry: its queries are shared by the
        # T frames of ONE object)
        q2 = queries.view(B * Q, -1)
        gs = ops.small_linear(q2, self.w_gs, self.b_gs, out_f16=True)
        qe = ops.vae_query_embed(q2, gs)
        qd = ops.gemm(qe, self.w_dq, None, ops.EPI_F16)
        ao = torch.empty((B, T, Q, H, d), dtype=F16, device=dev)
        lses = []
        for b in range(B):
            _, lse = ops.attention_fwd_lse(qd[b * Q:(b + 1) * Q].view(Q, H, d), kv4[b, :, :, 0], kv4[b, :, :, 1], scale,
                                           out=ao[b], q_shared=True)
            lses.append(lse)
        lat = ops.gemm(ao.view(B * T * Q, dim), self.w_dout, self.b_dout, ops.EPI_F16)
        ops.gemm(lat, self.w_o, self.b_o, ops.EPI_F32_COMPACT, out=out.view(B * T * Q, self.out_dim))
        sv.update(x_last=x, ctx=ctx, KV=KV, gs=gs, qe=qe, qd=qd, ao=ao, lse=lses, lat=lat)
        return out, sv

    # ------------------------------------------------------------------------------------------------ backward
    @staticmethod
    def _wgrad(dy, x):
        """dW fp32 [N_out, K_in] = dy[M, N_out]^T x[M, K_in]: both activations enter the tensor core MN-major as they
        lie in memory (gvf_gemm_tn_f16), split over the token dimension."""
        return ops.gemm_tn(dy, x)

    def backward(self, sv, dout):
        """dout [B, T, Q, out_dim] fp32 -> (grads {reference parameter name: fp32 tensor}, dz, dqueries)."""
        dev = self.dev
        B, Q, BT, L = sv["shape"]
        T, H, d, dim = self.T, self.H, self.d, self.dim
        M = BT * L
        scale = d ** -0.5
        dout = dout.detach().to(dev, F32).contiguous()
        g = {}
        c = "decoder_cross_attn.fn."
        dKV = torch.empty((M, 2 * dim), dtype=F16, device=dev)
        dkv4 = dKV.view(B, T, L, 2, H, d)
        kv4 = sv["KV"].view(B, T, L, 2, H, d)
        q2 = sv["queries"].view(B * Q, -1)
        do2 = dout.view(B * T * Q, self.out_dim)
        # to_outputs o to_out (model/autoencoder.py:562-574): out = (ao W_d^T + b_d) W_o^T + b_o with no non-linearity in
        # between, so the backward composes into rank-14 products -- no [B*T*Q, dim] x [dim, dim] dgrad / wgrad GEMM, no
        # d lat tensor:  d ao = d out (W_o W_d),  dW_d = W_o^T (d out^T ao),  db_d = (sum d out) W_o,  dW_o = d out^T lat
        wo = self.w_o[:self.out_dim].float()                                        # [14, dim]
        wc = ops.skinny_outer(self.w_o_t.float(), self.w_dout)                      # W_o W_d  [14, dim]
        g["to_outputs.weight"] = ops.skinny_outer(do2, sv["lat"])
        g["to_outputs.bias"] = ops.colsum(do2)
        s1 = ops.skinny_outer(do2, sv["ao"].view(B * T * Q, dim))                   # d out^T ao  [14, dim]
        g[c + "to_out.weight"] = ops.skinny_expand(self.w_o_t.float(), s1, out_f16=False)
        g[c + "to_out.bias"] = ops.skinny_expand(g["to_outputs.bias"].view(1, -1), wo, out_f16=False).view(-1)
        dao = ops.skinny_expand(do2, wc).view(B, T, Q, H, d)
        dqd = torch.empty((B * Q, dim), dtype=F16, device=dev)
        for b in range(B):
            ops.attention_bwd(sv["qd"][b * Q:(b + 1) * Q].view(Q, H, d), kv4[b, :, :, 0], kv4[b, :, :, 1], sv["ao"][b], dao[b],
                              sv["lse"][b], scale, dqd[b * Q:(b + 1) * Q].view(Q, H, d), dkv4[b, :, :, 0], dkv4[b, :, :, 1],
                              q_shared=True)
        dqe = ops.gemm_nn(dqd, self.w_dq)
        g[c + "to_q.weight"] = self._wgrad(dqd, sv["qe"])
        # query embedding: LN(LN(gs_embedding(q)) + LN(PointEmbed(q.xyz)))
        dgs, _ = ops.vae_query_embed_bwd(q2, sv["gs"], dqe)
        dq2 = ops.small_linear_bwd_input(dgs, self.w_gs)                          # [B*Q, 14]
        ops.vae_query_embed_bwd(q2, sv["gs"], dqe, dxyz=dq2, accumulate=True)       # + d PointEmbed / d xyz
        dqueries = dq2.view(B, Q, -1)
        g["gs_embedding.0.weight"] = ops.skinny_outer(q2, dgs).t().contiguous()
        g["gs_embedding.0.bias"] = ops.colsum(dgs)
        # to_kv of the decoder over all batch entries, PreNorm.norm_context
        dctx = ops.gemm_nn(dKV, self.w_dkv)
        g[c + "to_kv.weight"] = self._wgrad(dKV, sv["ctx"])
        dx = ops.ln_bwd(sv["x_last"], dctx, None, eps=1e-6)
        for i in reversed(range(self.depth)):
            ly, s = self.layers[i], sv["layers"][i]
            a, f = f"layers.{i}.0.fn.", f"layers.{i}.1.fn."
            # x2 = x1 + net.2(GEGLU(net.0(LN x1)))
            dG = ops.gemm_nn(dx, ly["w2"])
            g[f + "net.2.weight"] = self._wgrad(dx, s["G"])
            g[f + "net.2.bias"] = ops.colsum(dx)
            dHf = ops.geglu_bwd(s["Hf"], dG)
            dA2 = ops.gemm_nn(dHf, ly["w1"])
            g[f + "net.0.weight"] = self._wgrad(dHf, s["A2"])
            g[f + "net.0.bias"] = ops.colsum(dHf)
            dx1 = ops.ln_bwd(s["x1"], dA2, dx, eps=1e-6)
            # x1 = x0 + to_out(attention(to_q(LN x0), to_kv(LN x0)))
            dAO = ops.gemm_nn(dx1, ly["w_out"])
            g[a + "to_out.weight"] = self._wgrad(dx1, s["AO"].view(M, dim))
            g[a + "to_out.bias"] = ops.colsum(dx1)
            q5 = s["QKV"].view(BT, L, 3, H, d)
            dQKV = torch.empty_like(s["QKV"])
            d5 = dQKV.view(BT, L, 3, H, d)
            ops.attention_bwd(q5[:, :, 0], q5[:, :, 1], q5[:, :, 2], s["AO"], dAO.view(BT, L, H, d), s["lse"], scale,
                              d5[:, :, 0], d5[:, :, 1], d5[:, :, 2])
            dA = ops.gemm_nn(dQKV, ly["w_qkv"])
            wqkv = self._wgrad(dQKV, s["A"])
            g[a + "to_q.weight"], g[a + "to_kv.weight"] = wqkv[:dim], wqkv[dim:]
            dx = ops.ln_bwd(s["x0"], dA, dx1, eps=1e-6)
        # proj (model/autoencoder.py:585)
        dz = ops.small_linear_bwd_input(dx, self.w_proj)
        g["proj.weight"] = ops.skinny_outer(sv["z"], dx).t().contiguous()
        g["proj.bias"] = ops.colsum(dx)
        return g, dz.view(BT, L, -1), dqueries
