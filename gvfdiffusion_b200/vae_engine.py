"""Device engine of the motion-VAE decoder (reference model/autoencoder.py:579-609,
`GSKLTemporalVariationalAutoEncoder.decode`).

Host orchestration only; all arithmetic is in libgvf_b200.so.  Same math as the reference under
its fp16 autocast, with two structural changes:
  * `to_q` / `to_kv` of the latent self-attention are one GEMM over the concatenated weights;
  * the per-Gaussian query embedding and its `to_q` projection do not depend on the frame, so
    they are computed once per object instead of once per frame (autoencoder.py:557-561 repeats
    the queries over T first); the decoder cross-attention reads them through `q_shared`.
"""
import torch

from . import ops

F16, F32 = torch.float16, torch.float32


def _h(t, dev):
    return t.detach().to(device=dev, dtype=F16).contiguous()


def _b(t, dev):
    return t.detach().to(device=dev, dtype=F16).to(F32).contiguous()


class VAEDecodeEngine:
    def __init__(self, state_dict, heads, num_timesteps, device="cuda", chunk_size=8192):
        sd, dev = state_dict, torch.device(device)
        self.dev, self.H, self.T, self.chunk = dev, heads, num_timesteps, chunk_size
        self.dim = sd["proj.weight"].shape[0]
        self.d = self.dim // heads
        self.depth = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("layers."))
        self.w_proj, self.b_proj = _h(sd["proj.weight"], dev), _b(sd["proj.bias"], dev)
        self.layers = []
        for i in range(self.depth):
            a, f = f"layers.{i}.0.fn.", f"layers.{i}.1.fn."
            self.layers.append(dict(
                w_qkv=_h(torch.cat([sd[a + "to_q.weight"].detach().float().to(dev),
                                    sd[a + "to_kv.weight"].detach().float().to(dev)], 0), dev),
                w_out=_h(sd[a + "to_out.weight"], dev), b_out=_b(sd[a + "to_out.bias"], dev),
                w1=_h(sd[f + "net.0.weight"], dev), b1=_b(sd[f + "net.0.bias"], dev),
                w2=_h(sd[f + "net.2.weight"], dev), b2=_b(sd[f + "net.2.bias"], dev)))
            ly = self.layers[-1]
            if (ly["w1"].shape[0] // 2) % 128 == 0:     # GEGLU in the ff1 epilogue needs whole [128 value | 128 gate] tiles
                ly["w1g"], ly["b1g"] = ops.geglu_interleave(ly["w1"], ly["b1"])
        self.w_gs, self.b_gs = _h(sd["gs_embedding.0.weight"], dev), _b(sd["gs_embedding.0.bias"], dev)
        c = "decoder_cross_attn.fn."
        self.w_dq, self.w_dkv = _h(sd[c + "to_q.weight"], dev), _h(sd[c + "to_kv.weight"], dev)
        self.w_dout, self.b_dout = _h(sd[c + "to_out.weight"], dev), _b(sd[c + "to_out.bias"], dev)
        wo, bo = sd["to_outputs.weight"].detach().float().to(dev), sd["to_outputs.bias"].detach().float().to(dev)
        self.out_dim = wo.shape[0]
        pad = (-self.out_dim) % 8
        self.w_o = _h(torch.cat([wo, torch.zeros(pad, wo.shape[1], device=dev)], 0), dev)      # rows padded to 16
        self.b_o = _b(torch.cat([bo, torch.zeros(pad, device=dev)], 0), dev)

    def refresh(self, sd):
        """New parameter values into the SAME device buffers (after an optimiser step): device-side foreach casts, no host
        round trip, no allocation (the constructor's cost per training step otherwise)."""
        dim = self.dim
        dw, sw, db, sb = [self.w_proj], [sd["proj.weight"]], [self.b_proj], [sd["proj.bias"]]
        for i, ly in enumerate(self.layers):
            a, f = f"layers.{i}.0.fn.", f"layers.{i}.1.fn."
            dw += [ly["w_qkv"][:dim], ly["w_qkv"][dim:], ly["w_out"], ly["w1"], ly["w2"]]
            sw += [sd[a + "to_q.weight"], sd[a + "to_kv.weight"], sd[a + "to_out.weight"], sd[f + "net.0.weight"], sd[f + "net.2.weight"]]
            db += [ly["b_out"], ly["b1"], ly["b2"]]
            sb += [sd[a + "to_out.bias"], sd[f + "net.0.bias"], sd[f + "net.2.bias"]]
        c = "decoder_cross_attn.fn."
        dw += [self.w_gs, self.w_dq, self.w_dkv, self.w_dout, self.w_o[:self.out_dim]]
        sw += [sd["gs_embedding.0.weight"], sd[c + "to_q.weight"], sd[c + "to_kv.weight"], sd[c + "to_out.weight"], sd["to_outputs.weight"]]
        db += [self.b_gs, self.b_dout, self.b_o[:self.out_dim]]
        sb += [sd["gs_embedding.0.bias"], sd[c + "to_out.bias"], sd["to_outputs.bias"]]
        with torch.no_grad():
            torch._foreach_copy_(dw, [t.detach() for t in sw])
            torch._foreach_copy_(db, [t.detach().to(F16) for t in sb])              # fp16-valued fp32 biases
            for ly in self.layers:
                if "w1g" in ly:
                    ly["w1g"], ly["b1g"] = ops.geglu_interleave(ly["w1"], ly["b1"])

    def latent_layers(self, z):
        """z [(B*T), L, latent] fp32 -> x [(B*T)*L, dim] fp16 after proj + depth x (attn, GEGLU FF)."""
        BT, L, Cl = z.shape
        M, dim, H, d = BT * L, self.dim, self.H, self.d
        x = ops.small_linear(z.reshape(M, Cl).contiguous(), self.w_proj, self.b_proj, out_f16=True)
        A = torch.empty((M, dim), dtype=F16, device=self.dev)
        QKV = torch.empty((M, 3 * dim), dtype=F16, device=self.dev)
        AO = torch.empty((M, dim), dtype=F16, device=self.dev)
        Hf = torch.empty((M, self.layers[0]["w1"].shape[0]), dtype=F16, device=self.dev)
        G = torch.empty((M, Hf.shape[1] // 2), dtype=F16, device=self.dev)
        q5 = QKV.view(BT, L, 3, H, d)
        scale = d ** -0.5
        for ly in self.layers:
            ops.ln_mod(x, out=A, eps=1e-6)
            ops.gemm(A, ly["w_qkv"], None, ops.EPI_F16, out=QKV)            # [q | k | v] along the channels
            ops.attention(q5[:, :, 0], q5[:, :, 1], q5[:, :, 2], scale, out=AO.view(BT, L, H, d))
            ops.gemm(AO, ly["w_out"], ly["b_out"], ops.EPI_RESID_F16, out=x)
            ops.ln_mod(x, out=A, eps=1e-6)
            if "w1g" in ly:
                ops.gemm_geglu(A, ly["w1g"], ly["b1g"], out=G)                 # Linear + GEGLU, one kernel
            else:
                ops.gemm(A, ly["w1"], ly["b1"], ops.EPI_F16, out=Hf)
                ops.geglu(Hf, out=G)
            ops.gemm(G, ly["w2"], ly["b2"], ops.EPI_RESID_F16, out=x)
        return x

    def decode(self, z, queries):
        """z [(B*T), L, latent] fp32, queries [B, Q, 14] fp32 -> [B, T, Q, out_dim] fp32."""
        z = z.to(self.dev, F32)
        queries = queries.to(self.dev, F32).contiguous()
        B, Q, _ = queries.shape
        T, H, d, dim = self.T, self.H, self.d, self.dim
        BT, L, _ = z.shape
        assert BT == B * T
        x = self.latent_layers(z)
        ctx = ops.ln_mod(x, eps=1e-6)                                         # PreNorm.norm_context
        KV = ops.gemm(ctx, self.w_dkv, None, ops.EPI_F16)                     # [(B*T)*L, 2*dim]
        kv4 = KV.view(B, T, L, 2, H, d)
        out = torch.empty((B, T, Q, self.out_dim), dtype=F32, device=self.dev)
        scale = d ** -0.5
        for b in range(B):
            for s in range(0, Q, self.chunk):
                qc = queries[b, s:s + self.chunk]
                n = qc.shape[0]
                gs = ops.small_linear(qc, self.w_gs, self.b_gs, out_f16=True)
                qe = ops.vae_query_embed(qc, gs)                              # frame independent
                qd = ops.gemm(qe, self.w_dq, None, ops.EPI_F16).view(n, H, d)
                ao = ops.attention(qd, kv4[b, :, :, 0], kv4[b, :, :, 1], scale, q_shared=True)   # [T,n,H,d]
                lat = ops.gemm(ao.view(T * n, dim), self.w_dout, self.b_dout, ops.EPI_F16)
                if n == Q:      # single chunk: write straight into the result
                    ops.gemm(lat, self.w_o, self.b_o, ops.EPI_F32_COMPACT, out=out[b].view(T * Q, self.out_dim))
                else:
                    dl = torch.empty((T, n, self.out_dim), dtype=F32, device=self.dev)
                    ops.gemm(lat, self.w_o, self.b_o, ops.EPI_F32_COMPACT, out=dl.view(T * n, self.out_dim))
                    out[b, :, s:s + n].copy_(dl)
        return out
