"""BASELINE configs[4]: "main_vae.py train step (L1 + LPIPS + SSIM render loss), bs = 2 per GPU, DDP".

One step = reference train_vae.py:263-375 (`forward_backward` + `optimize`) on synthetic data of the shipped shapes:

  static VAE   SparseVAE.training_losses: 2 objects x 2048 voxels x 1024 features -> SparseTransformerVAE (12 + 12 swin
               blocks, 768 ch) -> to_representation (16384 Gaussians / object) -> one 512^2 render each -> L1 + 0.2 (1 - SSIM)
               + 1e-6 KL + volume / opacity regularisers
  motion VAE   get_gaussian_tensor (differentiable) -> GSKLTemporalVariationalAutoEncoder.forward: encode (FPS, KNN
               interpolation, cross attention, KL) + decode (12 layers, 24 x 512 latents, 16384 queries / object)
  losses       + 1e-5 KL + 0.1 x interpolation (xyz) loss (KNN 8) + 24 frames x 2 objects rendered at 512^2 with the predicted
               deltas (gradients to the deltas AND the canonical Gaussians) -> L1 + 0.2 (1 - SSIM)
  backward     through everything above into both models (loss scale 2^16, fp16 activation gradients, fp32 parameter grads)
  optimize     [DDP: one flat NCCL all-reduce of all gradients] -> un-scale -> clip_grad_norm 1.0 over both models -> AdamW
               (lr, 0.1 lr) -> EMA 0.9999 of both models; the engines' fp16 weight copies are refreshed from the fp32
               masters at the next forward (inside the timed steps, which run back to back)

LPIPS-VGG16 (0.2 x, both the static and the 48-frame term) runs with seeded-random weights -- the ImageNet VGG16 and the
v0.1 heads are a network download -- i.e. the criterion's full cost and gradient path (13 cuDNN convolutions on 2 x 50 images
of 512^2 per step, forward + backward) is inside the timed step.  Timed with CUDA events over whole steps.

    python tools/cfg5_step_bench.py [--steps 5]         -> one JSON line
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

LOSS_SCALE = 65536.0
STANDIN_LOSS_SCALE = 1024.0
KL_W, XYZ_W, L1_W, SSIM_W, LPIPS_W, KNN_K, BETA, LR, EMA = 1e-5, 0.1, 1.0, 0.2, 0.2, 8, 7.0, 1e-4, 0.9999
N_PC = 8192


def build(dev, seed=0, lpips=True):
    import bench as BN
    from gvfdiffusion_b200 import synthetic as S
    from gvfdiffusion_b200.model.sparse_voxel_diffusion import SparseTransformerVAE, SparseVAE
    from gvfdiffusion_b200.sparse.basic import SparseTensor
    from tools import static_vae_step_bench as SB
    B, T = SB.BATCH, BN.T_FRAMES
    torch.manual_seed(seed)
    static = SparseTransformerVAE(SB.RES_GRID, SB.CIN, SB.C, SB.COUT, SB.LAT, SB.NB, window_size=SB.WIN, use_fp16=True,
                                  use_old_attn_impl=False, norm_output=True)
    static.load_state_dict(SB.state_dict(seed))                       # random init incl. the zero-initialised output layers
    static = static.to(dev).train()
    _, vae = BN.build_models(dev, seed=seed)
    vae.train()
    from gvfdiffusion_b200.utils.lpips import LPIPS
    vgg = LPIPS(net_type="vgg").to(dev).eval() if lpips else None                      # train_vae.py:93
    fw = SparseVAE({"vae": static}, resolution=SB.RES_GRID, representation_config=SB.REP, device=dev, lambda_ssim=0.2,
                   lambda_lpips=0.2 if lpips else 0.0, lamda_kl=1e-6, regularizations=SB.REG, lpips=vgg)
    coords = torch.cat([torch.cat([torch.full((SB.NVOX, 1), b), SB.surface_voxels(seed + 1 + b)], 1) for b in range(B)]).int().to(dev)
    g = torch.Generator().manual_seed(seed + 9)
    x = SparseTensor(torch.randn(coords.shape[0], SB.CIN, generator=g).to(dev), coords)
    ext_s, intr_s = S.orbit_extrinsics(B, radius=1.2), S.intrinsics(40.0)[None].repeat(B, 1, 1)
    ext = torch.stack([S.orbit_extrinsics(T, radius=1.2) for _ in range(B)])           # [B, T, 4, 4], host
    intr = S.intrinsics(40.0)
    # tracked points: N_PC voxel centres jittered inside their voxels, moving on a smooth field
    pcs, deltas = [], []
    for b in range(B):
        c = coords[coords[:, 0] == b][:, 1:].float().cpu()
        idx = torch.randint(0, c.shape[0], (N_PC,), generator=g)
        p = (c[idx] + torch.rand(N_PC, 3, generator=g)) / SB.RES_GRID - 0.5
        t = torch.linspace(0, 1, T)[:, None, None]
        d = 0.05 * torch.sin(6.28 * t + 4.0 * p[None, :, [1, 2, 0]]) * t
        pcs.append(p)
        deltas.append(d)
    static_pc, delta_pc = torch.stack(pcs).to(dev), torch.stack(deltas).to(dev)
    moving_pc = static_pc.unsqueeze(1) + delta_pc
    S_ = dict(static=static, vae=vae, fw=fw, x=x, ext_s=ext_s, intr_s=intr_s, ext=ext, intr=intr, static_pc=static_pc,
              delta_pc=delta_pc, moving_pc=moving_pc, B=B, T=T, dev=dev, gen=torch.Generator(device=dev).manual_seed(seed + 3),
              lat_shape=(B * T, BN.N_LAT, BN.C_LAT), vgg=vgg)
    # targets: renders of the un-trained models under a different posterior draw
    with torch.no_grad():
        fw.renderers["MipGS"].rendering_options.resolution = SB.RES_IMG
        z = static.encode(x, noise=torch.randn(coords.shape[0], SB.LAT, generator=g).to(dev))
        reps = fw.to_representation(static.decode(z))["MipGS"]
        S_["image_s"] = fw.render_batch({"MipGS": reps}, ext_s, intr_s)["MipGS"]["rgb"].clone()
        gs = [m.gaussian_tensor() for m in reps]
        out = vae(gs, static_pc, delta_pc)["logits"]
        S_["image_m"] = torch.cat([fw.renderers["MipGS"].render_frames(reps[b], ext[b], intr, 1.2 * out[b])[0][:, :3] for b in range(B)]).clone()
    S_["opt"] = torch.optim.AdamW(vae.parameters(), lr=LR, weight_decay=0.0, fused=True)
    S_["opt_s"] = torch.optim.AdamW(static.parameters(), lr=LR * 0.1, weight_decay=0.0, fused=True)
    S_["params"] = list(vae.parameters()) + list(static.parameters())
    S_["ema"] = [p.detach().clone() for p in S_["params"]]
    return S_


def step(S_, world=1, dry=False):
    from gvfdiffusion_b200.train_vae import compute_interpolation_loss_delta_interp, get_gaussian_tensor
    from gvfdiffusion_b200.utils.loss_util import ssim_l1
    fw, vae, dev, B, T = S_["fw"], S_["vae"], S_["dev"], S_["B"], S_["T"]
    for p in S_["params"]:
        p.grad = None
    # ---- forward_backward (train_vae.py:263-353)
    noise_s = torch.randn(S_["x"].feats.shape[0], 8, device=dev, generator=S_["gen"])
    terms, reps = fw.training_losses(S_["x"], S_["image_s"], S_["ext_s"], S_["intr_s"], noise=noise_s)
    models = reps["MipGS"]
    static_gs = [get_gaussian_tensor(m) for m in models]
    loss = terms["loss"]
    out = vae(static_gs, S_["static_pc"], S_["delta_pc"], noise=torch.randn(S_["lat_shape"], device=dev, generator=S_["gen"]))
    loss = loss + out["kl"].mean() * KL_W
    pred = out["logits"]
    interp, _, _ = compute_interpolation_loss_delta_interp(static_gs, S_["static_pc"], S_["moving_pc"], pred, B, KNN_K, beta=BETA)
    loss = loss + interp * XYZ_W
    imgs = [fw.renderers["MipGS"].render_frames(models[b], S_["ext"][b], S_["intr"], pred[b], detach_static=False)[0][:, :3]
            for b in range(B)]
    pred_img = torch.cat(imgs)
    ssim_v, l1 = ssim_l1(pred_img, S_["image_m"])
    loss = loss + l1 * L1_W + (1.0 - ssim_v) * SSIM_W
    if S_["vgg"] is not None:                                                         # train_vae.py:329
        loss = loss + S_["vgg"](pred_img * 2 - 1.0, S_["image_m"] * 2 - 1.0) * LPIPS_W
    if dry:                                                  # forward only: the loss both arms must agree on
        return loss.detach()
    (loss * LOSS_SCALE).backward()
    # ---- optimize (train_vae.py:355-375)
    grads = [p.grad for p in S_["params"] if p.grad is not None]
    if world > 1:
        import torch.distributed as dist
        flat = torch._utils._flatten_dense_tensors(grads)
        dist.all_reduce(flat)
        flat.mul_(1.0 / (world * LOSS_SCALE))
        for g_, f in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
            g_.copy_(f)
        S_["allreduce_bytes"] = flat.numel() * 4
    else:
        torch._foreach_mul_(grads, 1.0 / LOSS_SCALE)
    torch.nn.utils.clip_grad_norm_(S_["params"], 1.0, foreach=True)
    S_["opt"].step()
    S_["opt_s"].step()
    live = [(e, p) for e, p in zip(S_["ema"], S_["params"])]
    torch._foreach_mul_([e for e, _ in live], EMA)
    torch._foreach_add_([e for e, _ in live], [p.detach() for _, p in live], alpha=1 - EMA)
    return loss


# ------------------------------------------------------------------------------------------ stand-in of the reference's step
class Standin:
    """The same joint step the way the reference executes it on a GPU, as far as it can be restated without its absent
    dependencies (labelled a stand-in wherever its number is printed):
      * static VAE: tools/gpu_reference.sparse_vae_forward (fp16 block weights converted once, fp16 residual stream, flash_attn
        2.8.3 varlen window attention with index gather / scatter, cuBLAS Linear, torch autograd, torch to_representation); no
        activation checkpointing although the reference trains with mem_ratio 0.2;
      * motion VAE decode: tools/gpu_reference.vae_decode (fp16 autocast, flash_attn forward / backward, cuBLAS, autograd);
        its ENCODE stays on this repo's kernels (the reference's needs torch_cluster / pytorch3d) -- in the stand-in's favour;
      * LPIPS: the plain torch modules under fp16 autocast (13 cuDNN convolutions, torch ReLU / pool / normalisation / head);
      * rasteriser, SSIM / L1, KNN interpolation loss, get_gaussian_tensor: this repo's kernels in both arms;
      * optimiser: fp32 masters for everything, fp16 block copies re-cast after the step, same clip / 2 x AdamW / EMA."""

    def __init__(self, S_):
        from gvfdiffusion_b200.model.sparse_voxel_diffusion import SparseVAE
        from gvfdiffusion_b200.representations.gaussian import GaussianModel
        from gvfdiffusion_b200.sparse.attention import calc_window_partition
        from tools import gpu_reference as GR
        from tools import static_vae_step_bench as SB
        dev, coords = S_["dev"], S_["x"].coords
        self.S = S_
        self.sd = {k: v.detach().clone().requires_grad_(True) for k, v in S_["static"].state_dict().items()}
        self.sd16 = {k: v.detach().half().clone().requires_grad_(True) for k, v in self.sd.items() if "coder." in k}
        vae = S_["vae"]
        self.enc_names = list(vae._enc_names)
        self.vsd = {k: v.detach().clone().requires_grad_(True) for k, v in vae.state_dict().items() if k not in self.enc_names}
        parts = []
        for shift in (0, SB.WIN // 2):
            fwd, bwd, seq, _ = calc_window_partition(coords, SB.WIN, shift)
            cu = torch.zeros(seq.shape[0] + 1, dtype=torch.int32, device=dev)
            cu[1:] = torch.cumsum(seq, 0)
            parts.append((fwd, bwd, cu, int(seq.max())))
        outer = self

        class FW(SparseVAE):
            def _backbone_forward(self, feats, noise=None):
                out, mean, logvar = GR.sparse_vae_forward(outer.sd, outer.sd16, feats.feats, feats.coords, noise, SB.H, SB.NB, parts)
                return out, 0.5 * torch.mean(mean.pow(2) + logvar.exp() - logvar - 1), mean, logvar

            def to_representation(self, x):
                cfg = self.rep_config["MipGS"]
                raw = GR.to_representation_torch(x.feats, x.coords, 8, [cfg["lr"][n] for n in ("_xyz", "_features_dc", "_scaling", "_rotation", "_opacity")],
                                                 self.resolution, cfg["voxel_size"], self.perturbation["MipGS"])
                reps = []
                for sl in x.layout:
                    rep = GaussianModel(sh_degree=0, aabb=[-0.5, -0.5, -0.5, 1.0, 1.0, 1.0], mininum_kernel_size=cfg["3d_filter_kernel_size"],
                                        scaling_bias=cfg["scaling_bias"], opacity_bias=cfg["opacity_bias"],
                                        scaling_activation=cfg["scaling_activation"], device=self.device)
                    gs = slice(sl.start * 8, sl.stop * 8)
                    rep._xyz, rep._features_dc, rep._scaling, rep._rotation, rep._opacity = (t[gs] for t in raw)
                    reps.append(rep)
                return {"MipGS": reps}

        vgg = S_["vgg"]

        def lpips_plain(a, b):                              # the reference's modules as they are, under autocast
            with torch.autocast("cuda", dtype=torch.float16):
                fa = vgg.taps(a)
                with torch.no_grad():
                    fb = vgg.taps(b)
                res = [l((vgg._unit(p) - vgg._unit(q)) ** 2).mean((2, 3), True) for p, q, l in zip(fa, fb, vgg.lin)]
            return torch.sum(torch.cat(res, 0)).float() / a.shape[0]

        class _Plain:
            def __call__(self, a, b):
                return lpips_plain(a, b)

        self.lpips_plain = lpips_plain if vgg is not None else None
        self.fw = FW({}, resolution=SB.RES_GRID, representation_config=SB.REP, device=dev, lambda_ssim=0.2,
                     lambda_lpips=0.2 if vgg is not None else 0.0, lamda_kl=1e-6, regularizations=SB.REG,
                     lpips=_Plain() if vgg is not None else None)
        self.fw.renderers = S_["fw"].renderers             # same rasteriser objects (and learned workspace sizes) in both arms
        import copy
        vae._engine = vae._enc_engine = vae._train_engine = None      # engines hold device buffers: rebuilt per module
        self.vae_enc = copy.deepcopy(vae)                   # this arm trains its OWN copy of the encoder
        named = dict(self.vae_enc.named_parameters())
        self.enc_params = [named[n] for n in self.enc_names]
        self.masters_s = list(self.sd.values())
        self.masters_m = list(self.vsd.values()) + self.enc_params
        self.opt = torch.optim.AdamW(self.masters_m, lr=LR, weight_decay=0.0, fused=True)
        self.opt_s = torch.optim.AdamW(self.masters_s, lr=LR * 0.1, weight_decay=0.0, fused=True)
        self.params = self.masters_m + self.masters_s
        self.ema = [p.detach().clone() for p in self.params]
        self.GR = GR

    def step(self, dry=False):
        from gvfdiffusion_b200.train_vae import compute_interpolation_loss_delta_interp, get_gaussian_tensor
        from gvfdiffusion_b200.pipeline import pad_static_gs
        from gvfdiffusion_b200.utils.loss_util import ssim_l1
        S_, fw, dev = self.S, self.fw, self.S["dev"]
        vae, B, T = S_["vae"], S_["B"], S_["T"]
        for p in self.params + list(self.sd16.values()):
            p.grad = None
        noise_s = torch.randn(S_["x"].feats.shape[0], 8, device=dev, generator=S_["gen"])
        terms, reps = fw.training_losses(S_["x"], S_["image_s"], S_["ext_s"], S_["intr_s"], noise=noise_s)
        models = reps["MipGS"]
        static_gs = [get_gaussian_tensor(m) for m in models]
        loss = terms["loss"]
        kl, z, _, _ = self.vae_enc.encode(S_["static_pc"], S_["delta_pc"], static_gs,
                                          torch.randn(S_["lat_shape"], device=dev, generator=S_["gen"]))
        padded, _ = pad_static_gs(static_gs)
        pred = self.GR.vae_decode(self.vsd, z, padded, vae.heads, T, vae.depth)
        loss = loss + kl.mean() * KL_W
        interp, _, _ = compute_interpolation_loss_delta_interp(static_gs, S_["static_pc"], S_["moving_pc"], pred, B, KNN_K, beta=BETA)
        loss = loss + interp * XYZ_W
        imgs = [fw.renderers["MipGS"].render_frames(models[b], S_["ext"][b], S_["intr"], pred[b], detach_static=False)[0][:, :3]
                for b in range(B)]
        pred_img = torch.cat(imgs)
        ssim_v, l1 = ssim_l1(pred_img, S_["image_m"])
        loss = loss + l1 * L1_W + (1.0 - ssim_v) * SSIM_W
        if self.lpips_plain is not None:
            loss = loss + self.lpips_plain(pred_img * 2 - 1.0, S_["image_m"] * 2 - 1.0) * LPIPS_W
        if dry:
            return loss.detach()
        # the fp16 block weights receive fp16 GRADIENTS here (this arm has no fp32-accumulating wgrad): 2^16 x the weight
        # gradients overflows fp16, what a GradScaler answers by backing off -- a fixed 2^10 for this arm
        scale = STANDIN_LOSS_SCALE
        (loss * scale).backward()
        # fp16 block gradients -> fp32 masters (the reference's model_grads_to_master_grads)
        for k, t in self.sd16.items():
            g16 = t.grad
            if g16 is not None:
                m = self.sd[k]
                m.grad = g16.float() if m.grad is None else m.grad + g16.float()
        grads = [p.grad for p in self.params if p.grad is not None]
        torch._foreach_mul_(grads, 1.0 / scale)
        torch.nn.utils.clip_grad_norm_(self.params, 1.0, foreach=True)
        self.opt.step()
        self.opt_s.step()
        torch._foreach_mul_(self.ema, EMA)
        torch._foreach_add_(self.ema, [p.detach() for p in self.params], alpha=1 - EMA)
        with torch.no_grad():                               # master_params_to_model_params
            torch._foreach_copy_(list(self.sd16.values()), [self.sd[k] for k in self.sd16])
        return loss


def measure(steps=5, warmup=3, seed=0, world=1, device=None, lpips=True, standin=False):
    dev = device if device is not None else torch.device("cuda", 0)
    S_ = build(dev, seed, lpips)
    first = None
    if standin and world == 1:
        # both arms from the SAME initial weights and the same posterior noise: their forward losses must agree.  Two dry
        # passes each: the first one teaches the rasteriser its workspace size
        st = Standin(S_)
        first = {}
        for name, fn in (("ours", lambda: step(S_, 1, dry=True)), ("standin", lambda: st.step(dry=True))):
            for _ in range(2):
                S_["gen"].manual_seed(seed + 3)
                with torch.enable_grad():
                    first[name] = float(fn())
        S_["gen"].manual_seed(seed + 3)
    for _ in range(warmup):
        step(S_, world)
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step(S_, world)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    B, T = S_["B"], S_["T"]
    n_static = sum(p.numel() for p in S_["static"].parameters())
    n_motion = sum(p.numel() for p in S_["vae"].parameters())
    res = {"metric": "train-step samples/s (main_vae.py joint step: static VAE + motion VAE + renders, fwd + bwd + optimizer)",
           "value": world * B / (ms / 1e3), "unit": "samples/s", "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "loss": float(loss.detach()),
           "dtype": "f16 (fp32 accumulate, fp32 master weights / gradients / AdamW)", "data": "synthetic",
           "config": {"workload": f"BASELINE.json configs[4]: per-GPU batch {B}; static VAE 2 x 2048 voxels (12+12 swin blocks, 768 ch) + "
                                  f"one 512^2 render each; motion VAE encode + decode (12 layers, {T} x 512 latents, 16384 queries / "
                                  f"object) + {B * T} renders at 512^2 with predicted deltas; L1 + SSIM + KL + interpolation (KNN 8) + "
                                  "regularisers" + (" + 0.2 LPIPS-VGG16 (seeded-random weights) on all 50 renders" if lpips else "; no LPIPS") +
                                  "; backward; grad clip; 2 x AdamW; EMA",
                      "parameters": {"static_vae": n_static, "motion_vae": n_motion}}}
    if world > 1:
        res["ddp"] = {"allreduce_bytes_per_step": S_.get("allreduce_bytes"), "where": "inside the step, after backward"}
    if standin and world == 1:
        for _ in range(warmup):
            st.step()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            loss2 = st.step()
        e1.record()
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / steps
        res["gpu_reference"] = {"kind": "stand-in", "ms_per_step": ms2, "value": B / (ms2 / 1e3), "unit": "samples/s",
                                "speedup": ms2 / ms, "loss": float(loss2.detach()),
                                "first_step_loss": {"ours": first["ours"], "standin": first["standin"],
                                                    "note": "same initial weights and posterior noise, forward only"},
                                "what": " ".join(Standin.__doc__.split())}
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--one-step", action="store_true", help="two un-timed steps (for ncu launch lists)")
    ap.add_argument("--no-lpips", action="store_true", help="leave the LPIPS terms out (A/B: what the cuDNN convolutions cost)")
    ap.add_argument("--standin", action="store_true", help="also time the stand-in of the reference's execution of the step")
    a = ap.parse_args()
    if a.one_step:
        S0 = build(torch.device("cuda", 0))
        for _ in range(2):
            step(S0)
            torch.cuda.synchronize()
        sys.exit(0)
    print(json.dumps(measure(a.steps, a.warmup, lpips=not a.no_lpips, standin=a.standin)))
