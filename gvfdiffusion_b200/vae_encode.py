"""Device engine of the motion-VAE ENCODER (reference model/autoencoder.py:502-550 `encode`, :451-500
`compute_delta_interp`): FPS of every object's Gaussians to `num_latents` anchors, KNN + RBF interpolation of the tracked
points' motion onto the anchors, token embeddings, one cross-attention block (anchor tokens x point tokens) + GEGLU
feed-forward on an fp32 residual stream, mean / logvar heads, DiagonalGaussian sample and KL.

Host orchestration only.  Same math as the reference under its fp16 autocast (Linear / attention fp16 with fp32
accumulation, LayerNorms and the residual stream fp32, mean / logvar rounded to fp16); `torch_cluster.fps` starts at a
random point -- gvf_fps starts at index 0 (deterministic), like `pipeline.sample_gs`.  The posterior noise is drawn with
`torch.randn(mean.shape)` on the host exactly like the reference (:316), so seeds carry over.
"""
import torch

from . import _lib, ops
from . import train_vae as TV
from ._lib import check, current_stream, ptr

F16, F32 = torch.float16, torch.float32


class VAEEncodeEngine:
    def __init__(self, state_dict, heads, num_latents, knn_k=8, beta=7.0, device="cuda"):
        sd, dev = state_dict, torch.device(device)
        h = lambda t: t.detach().to(device=dev, dtype=F16).contiguous()
        b = lambda t: t.detach().to(device=dev, dtype=F16).to(F32).contiguous()
        self.dev, self.H, self.L, self.knn_k, self.beta = dev, heads, num_latents, knn_k, beta
        a, f = "cross_attend_blocks.0.fn.", "cross_attend_blocks.1.fn."
        self.w_in, self.b_in = h(sd["input_embedding.0.weight"]), b(sd["input_embedding.0.bias"])
        self.dim = self.w_in.shape[0]
        self.d = self.dim // heads
        self.w_q, self.w_kv = h(sd[a + "to_q.weight"]), h(sd[a + "to_kv.weight"])
        self.w_out, self.b_out = h(sd[a + "to_out.weight"]), b(sd[a + "to_out.bias"])
        self.w1, self.b1 = h(sd[f + "net.0.weight"]), b(sd[f + "net.0.bias"])
        self.w2, self.b2 = h(sd[f + "net.2.weight"]), b(sd[f + "net.2.bias"])
        if (self.w1.shape[0] // 2) % 128 == 0:
            self.w1g, self.b1g = ops.geglu_interleave(self.w1, self.b1)
        # mean_fc and logvar_fc as one GEMM
        self.latent = sd["mean_fc.weight"].shape[0]
        wm = torch.cat([sd["mean_fc.weight"].detach().float().to(dev), sd["logvar_fc.weight"].detach().float().to(dev)], 0)
        bm = torch.cat([sd["mean_fc.bias"].detach().float().to(dev), sd["logvar_fc.bias"].detach().float().to(dev)], 0)
        pad = (-wm.shape[0]) % 8
        self.w_ml = h(torch.cat([wm, torch.zeros(pad, wm.shape[1], device=dev)], 0))
        self.b_ml = b(torch.cat([bm, torch.zeros(pad, device=dev)], 0))

    def refresh(self, sd):
        """New parameter values into the same device buffers (see VAEDecodeEngine.refresh)."""
        a, f, lat = "cross_attend_blocks.0.fn.", "cross_attend_blocks.1.fn.", self.latent
        dw = [self.w_in, self.w_q, self.w_kv, self.w_out, self.w1, self.w2, self.w_ml[:lat], self.w_ml[lat:2 * lat]]
        sw = [sd["input_embedding.0.weight"], sd[a + "to_q.weight"], sd[a + "to_kv.weight"], sd[a + "to_out.weight"],
              sd[f + "net.0.weight"], sd[f + "net.2.weight"], sd["mean_fc.weight"], sd["logvar_fc.weight"]]
        db = [self.b_in, self.b_out, self.b1, self.b2, self.b_ml[:lat], self.b_ml[lat:2 * lat]]
        sb = [sd["input_embedding.0.bias"], sd[a + "to_out.bias"], sd[f + "net.0.bias"], sd[f + "net.2.bias"], sd["mean_fc.bias"],
              sd["logvar_fc.bias"]]
        with torch.no_grad():
            torch._foreach_copy_(dw, [t.detach() for t in sw])
            torch._foreach_copy_(db, [t.detach().to(F16) for t in sb])
            if hasattr(self, "w1g"):
                self.w1g, self.b1g = ops.geglu_interleave(self.w1, self.b1)

    def _embed(self, disp, xyz, rows_per_xyz_row):
        """disp [R, 3] fp32 per (batch, frame, point) row, xyz [B * n, 3] per point -> fp32 [R, dim]."""
        R = disp.shape[0]
        lin = ops.small_linear(disp, self.w_in, self.b_in, out_f16=True)
        out = torch.empty((R, self.dim), dtype=F32, device=self.dev)
        check(_lib.lib().gvf_vae_embed_sum(ptr(xyz), xyz.stride(0), ptr(rows_per_xyz_row), ptr(lin), R, self.dim, ptr(out),
                                           current_stream()), "gvf_vae_embed_sum")
        return out

    @torch.no_grad()
    def encode(self, static_pc, delta_pc, static_gs_list, noise=None, save=None):
        """static_pc [B, N, 3], delta_pc [B, T, N, 3], static_gs_list: B tensors [P_b, 14]
        -> dict(kl [(B T)], x [(B T), L, latent], mean, logvar, sampled_static_gs [B, L, 14]).
        save: a dict that receives the activations `backward` needs (training step)."""
        dev, L, H, d, dim = self.dev, self.L, self.H, self.d, self.dim
        static_pc = static_pc.to(dev, F32).contiguous()
        delta_pc = delta_pc.to(dev, F32).contiguous()
        B, N, _ = static_pc.shape
        T = delta_pc.shape[1]
        sampled = []
        for g in static_gs_list:                                              # :519-527 (torch_cluster.fps per entry)
            g = g.to(dev, F32).contiguous()
            sampled.append(g.index_select(0, ops.fps(g, L).long()))
        sampled = torch.stack(sampled)                                        # [B, L, 14]
        gs_xyz = sampled[:, :, :3].contiguous()
        moving = delta_pc + static_pc.unsqueeze(1)                            # :529 (the reference's own fp32 add)
        kd, ki, _ = TV.knn_points(gs_xyz, static_pc, K=self.knn_k)
        est = TV.interpolate_deltas(kd, ki, static_pc, moving, None, True, self.beta)     # [B, T, L, 3]
        # row (b, t, n) of the token tensors takes its position from row (b, n)
        bt = torch.arange(B * T, device=dev, dtype=torch.int32) // T
        rows_a = (bt[:, None] * L + torch.arange(L, device=dev, dtype=torch.int32)[None]).reshape(-1).contiguous()
        rows_c = (bt[:, None] * N + torch.arange(N, device=dev, dtype=torch.int32)[None]).reshape(-1).contiguous()
        A = self._embed(est.reshape(B * T * L, 3), gs_xyz.view(B * L, 3), rows_a)          # fp32 [(B T) L, dim]
        Cx = self._embed(delta_pc.reshape(B * T * N, 3), static_pc.view(B * N, 3), rows_c)
        # cross_attend_blocks[0]: x = Attention(LN a, LN ctx) + a ; [1]: x = FF(LN x) + x   (:538-539)
        q = ops.gemm(ops.ln_mod(A, eps=1e-6), self.w_q, None, ops.EPI_F16)
        kv = ops.gemm(ops.ln_mod(Cx, eps=1e-6), self.w_kv, None, ops.EPI_F16).view(B * T, N, 2, H, d)
        if save is not None:                                                   # training: the forward that also leaves LSE2
            ao, save["lse"] = ops.attention_fwd_lse(q.view(B * T, L, H, d), kv[:, :, 0], kv[:, :, 1], d ** -0.5)
            return self._encode_train(save, A, Cx, q, kv, ao, est, delta_pc, sampled, noise, B, T, N)
        ao = ops.attention(q.view(B * T, L, H, d), kv[:, :, 0], kv[:, :, 1], d ** -0.5)
        x = A
        ops.gemm(ao.view(B * T * L, dim), self.w_out, self.b_out, ops.EPI_RESID_F32, out=x)
        hmid = ops.ln_mod(x, eps=1e-6)
        if hasattr(self, "w1g"):
            G = ops.gemm_geglu(hmid, self.w1g, self.b1g)
        else:
            G = ops.geglu(ops.gemm(hmid, self.w1, self.b1, ops.EPI_F16))
        ops.gemm(G, self.w2, self.b2, ops.EPI_RESID_F32, out=x)
        return self._heads(x, sampled, noise, B, T)

    def _heads(self, x, sampled, noise, B, T, save=None):
        dev, L = self.dev, self.L
        # heads: Linear under autocast -> fp16-rounded values
        ml = torch.empty((B * T * L, self.w_ml.shape[0]), dtype=F32, device=dev)
        x16 = ops.cast_f16(x)
        ops.gemm(x16, self.w_ml, self.b_ml, ops.EPI_F32_COMPACT, out=ml)
        mean = ml[:, :self.latent].contiguous().view(B * T, L, self.latent)
        logvar = ml[:, self.latent:2 * self.latent].contiguous().view(B * T, L, self.latent)
        if noise is None:
            noise = torch.randn(mean.shape)                                    # host RNG, like the reference (:316)
        noise = noise.to(dev, F32).contiguous()
        sample, kl = torch.empty_like(mean), torch.empty(B * T, dtype=F32, device=dev)
        check(_lib.lib().gvf_diag_gaussian(ptr(mean), ptr(logvar), ptr(noise), B * T, L * self.latent, ptr(sample), ptr(kl),
                                           current_stream()), "gvf_diag_gaussian")
        if save is not None:
            save.update(x16=x16, mean=mean, logvar=logvar, noise=noise)
        return {"kl": kl, "x": sample, "mean": mean, "logvar": logvar.clamp(-30.0, 20.0), "sampled_static_gs": sampled}

    # ------------------------------------------------------------------------------------------------ training step
    def _encode_train(self, sv, A, Cx, q, kv, ao, est, delta_pc, sampled, noise, B, T, N):
        """Tail of `encode` with every activation the backward needs kept (un-fused GEGLU, out-of-place residuals)."""
        L, dim = self.L, self.dim
        M = B * T * L
        qn, cn = ops.ln_mod(A, eps=1e-6), ops.ln_mod(Cx, eps=1e-6)            # recomputed: cheap, keeps `encode` linear
        x1 = A.clone()
        ops.gemm(ao.view(M, dim), self.w_out, self.b_out, ops.EPI_RESID_F32, out=x1)
        hmid = ops.ln_mod(x1, eps=1e-6)
        Hf = ops.gemm(hmid, self.w1, self.b1, ops.EPI_F16)
        G = ops.geglu(Hf)
        x2 = x1.clone()
        ops.gemm(G, self.w2, self.b2, ops.EPI_RESID_F32, out=x2)
        sv.update(A=A, Cx=Cx, qn=qn, cn=cn, q=q, kv=kv, ao=ao, x1=x1, hmid=hmid, Hf=Hf, G=G, x2=x2, est=est, delta_pc=delta_pc,
                  shape=(B, T, N))
        return self._heads(x2, sampled, noise, B, T, save=sv)

    def forward_train(self, static_pc, delta_pc, static_gs_list, noise=None):
        sv = {}
        return self.encode(static_pc, delta_pc, static_gs_list, noise, save=sv), sv

    def backward(self, sv, dx, dkl):
        """dx [(B T), L, latent] (gradient of the sampled latent), dkl [(B T)] or None -> {encoder parameter name: fp32 grad}.
        static_pc / delta_pc / the Gaussians are data (the reference interpolates them under no_grad, :470)."""
        # dgrad GEMMs read the [out, in] weights directly (ops.gemm_nn): no transposed copies
        dev, L, H, d, dim, lat = self.dev, self.L, self.H, self.d, self.dim, self.latent
        B, T, N = sv["shape"]
        M = B * T * L
        g = {}
        a, f = "cross_attend_blocks.0.fn.", "cross_attend_blocks.1.fn."
        wg = lambda dy, x: ops.gemm_tn(dy, x)
        # DiagonalGaussian + heads
        dmean, dlogvar = torch.empty_like(sv["mean"]), torch.empty_like(sv["mean"])
        dxc = None if dx is None else dx.detach().to(dev, F32).contiguous()
        dkc = None if dkl is None else dkl.detach().to(dev, F32).contiguous()
        check(_lib.lib().gvf_diag_gaussian_bwd(ptr(sv["mean"]), ptr(sv["logvar"]), ptr(sv["noise"]), ptr(dxc), ptr(dkc), B * T,
                                               L * lat, ptr(dmean), ptr(dlogvar), current_stream()), "gvf_diag_gaussian_bwd")
        dml = torch.zeros((M, self.w_ml.shape[0]), dtype=F16, device=dev)
        dml[:, :lat] = dmean.view(M, lat)
        dml[:, lat:2 * lat] = dlogvar.view(M, lat)
        wml = wg(dml, sv["x16"])                                                 # [32, dim]
        g["mean_fc.weight"], g["logvar_fc.weight"] = wml[:lat], wml[lat:2 * lat]
        bml = ops.colsum(dml)
        g["mean_fc.bias"], g["logvar_fc.bias"] = bml[:lat], bml[lat:2 * lat]
        dx2 = ops.gemm_nn(dml, self.w_ml)                      # [M, dim]
        # feed-forward block: x2 = x1 + net.2(GEGLU(net.0(LN x1)))
        dG = ops.gemm_nn(dx2, self.w2)
        g[f + "net.2.weight"], g[f + "net.2.bias"] = wg(dx2, sv["G"]), ops.colsum(dx2)
        dHf = ops.geglu_bwd(sv["Hf"], dG)
        dh = ops.gemm_nn(dHf, self.w1)
        g[f + "net.0.weight"], g[f + "net.0.bias"] = wg(dHf, sv["hmid"]), ops.colsum(dHf)
        dx1 = ops.ln_bwd(sv["x1"], dh, dx2, eps=1e-6)
        # cross-attention block: x1 = a + to_out(attention(to_q(LN a), to_kv(LN ctx)))
        dao = ops.gemm_nn(dx1, self.w_out)
        g[a + "to_out.weight"], g[a + "to_out.bias"] = wg(dx1, sv["ao"].view(M, dim)), ops.colsum(dx1)
        kv = sv["kv"]
        dq, dkv = torch.empty_like(sv["q"]), torch.empty_like(kv)
        ops.attention_bwd(sv["q"].view(B * T, L, H, d), kv[:, :, 0], kv[:, :, 1], sv["ao"], dao.view(B * T, L, H, d), sv["lse"],
                          d ** -0.5, dq.view(B * T, L, H, d), dkv[:, :, 0], dkv[:, :, 1])
        dkv2 = dkv.view(B * T * N, 2 * dim)
        g[a + "to_q.weight"] = wg(dq, sv["qn"])
        g[a + "to_kv.weight"] = wg(dkv2, sv["cn"])
        dqn = ops.gemm_nn(dq, self.w_q)
        dcn = ops.gemm_nn(dkv2, self.w_kv)
        dA = ops.ln_bwd(sv["A"], dqn, dx1, eps=1e-6)                              # + the residual branch
        dC = ops.ln_bwd(sv["Cx"], dcn, None, eps=1e-6)
        # token embeddings: LN_1e-5(Linear(3 -> dim)(disp)) + LN_1e-5(PointEmbed) -- only the Linear has parameters
        est2, dpc2 = sv["est"].reshape(M, 3), sv["delta_pc"].reshape(B * T * N, 3)
        lin_a = ops.small_linear(est2, self.w_in, self.b_in, out_f16=True)
        lin_c = ops.small_linear(dpc2, self.w_in, self.b_in, out_f16=True)
        dlin_a = ops.ln_bwd(lin_a, dA, None, eps=1e-5)
        dlin_c = ops.ln_bwd(lin_c, dC, None, eps=1e-5)
        wt = ops.skinny_outer(est2, dlin_a)
        wt = ops.skinny_outer(dpc2, dlin_c, out=wt, accumulate=True)             # [3, dim] = dW^T
        g["input_embedding.0.weight"] = wt.t().contiguous()
        bsum = ops.colsum(dlin_a)
        g["input_embedding.0.bias"] = ops.colsum(dlin_c, out=bsum, accumulate=True)
        return g
