#!/usr/bin/env python
"""inference_dpm_latent.py -- the reference script's flag surface (reference
inference_dpm_latent.py:276-316) over the B200-native hot path:

    canonical Gaussians -> FPS conditioning -> DPM-Solver++ over the DiT -> motion-VAE decode
    -> frame-batched canonical+delta rasterisation            (reference :205-272)

Out of scope here and therefore replaced by inputs (TRELLIS image->3D stage,
CLIP alignment, dataset loading, PNG/ffmpeg output): the canonical Gaussians and the DINOv2
conditioning come from `--data_dir/<name>.pt` files ({"gaussian": dict of raw GaussianModel tensors,
"cond_images": (T,1370,1024)}) or, when absent, from the seeded synthetic generator; frames are
written as one uint8 tensor per object (`rgb_<id>.pt`, (T, cameras, H, W, 3)) instead of 4096 PNGs.

Multi-GPU: launch with torchrun; objects are sharded round-robin across ranks (the reference runs
every object on every rank)."""
import argparse
import os
from collections import OrderedDict

import torch
import yaml

from gvfdiffusion_b200 import parallel, synthetic
from gvfdiffusion_b200 import raster as R
from gvfdiffusion_b200.model.autoencoder import GSKLTemporalVariationalAutoEncoder
from gvfdiffusion_b200.model.dit import DiT
from gvfdiffusion_b200.pipeline import GVFPipeline

DEFAULT_CFG = {
    "model": dict(resolution=512, in_channels=16, out_channels=16, model_channels=512, static_cond_channels=14,
                  image_cond_channels=1024, num_blocks=12, num_heads=16, mlp_ratio=4, pe_mode="ape", qk_rms_norm=True,
                  use_fp16=True, no_temporal_attn=False),
    "motion_vae": dict(depth=12, dim=768, queries_dim=768, output_dim=14, num_inputs=8192, num_latents=512,
                       latent_dim=16, heads=12, dim_head=-1, weight_tie_layers=False, decoder_ff=False,
                       enable_flash_attn=True),
}


def strip_module(sd):
    out = OrderedDict()
    for k, v in sd.items():
        out[k[7:] if k.startswith("module.") else k] = v
    return out


def cosine_betas(steps=1000):
    import math
    import numpy as np
    f = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
    b = np.array([min(1 - f((i + 1) / steps) / f(i / steps), 0.999) for i in range(steps)], dtype=np.float64)
    ac, out, last = np.cumprod(1.0 - b), [], 1.0
    for a in ac:                               # SpacedDiffusion re-derivation, reference model/respace.py:123-131
        out.append(1 - a / last)
        last = a
    return torch.from_numpy(np.array(out, dtype=np.float64))


def create_argparser():
    def none_or_str(value):
        return None if value.lower() == "none" else value
    p = argparse.ArgumentParser()
    p.add_argument("--exp_name", type=str, default="/tmp/output/")
    p.add_argument("--model_name", type=str, default=None, help="(HF download is out of scope; ignored)")
    p.add_argument("--download_assets", action="store_true")
    p.add_argument("--assets_dir", type=str, default="./assets")
    p.add_argument("--ckpt", type=str, default=None)
    p.add_argument("--batch_size", type=int, default=1)
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--use_fp16", action="store_true")
    p.add_argument("--config", type=str, default="configs/diffusion.yml")
    p.add_argument("--data_dir", type=str, default="data/")
    p.add_argument("--start_idx", type=int, default=0)
    p.add_argument("--end_idx", type=int, default=10)
    p.add_argument("--txt_file", type=str, default="video_names.txt")
    p.add_argument("--static_mean_file", type=none_or_str, default=None)
    p.add_argument("--static_std_file", type=none_or_str, default=None)
    p.add_argument("--deformation_mean_file", type=none_or_str, default=None)
    p.add_argument("--deformation_std_file", type=none_or_str, default=None)
    p.add_argument("--load_camera", type=int, default=1)
    p.add_argument("--num_timesteps", type=int, default=24)
    p.add_argument("--num_samples", type=int, default=10)
    p.add_argument("--rescale_timesteps", type=int, default=100)
    p.add_argument("--guidance_scale", type=float, default=1.0)
    p.add_argument("--guidance_scale2", type=float, default=1.0)
    p.add_argument("--adaptive", action="store_true")
    p.add_argument("--vae_ckpt", type=str, default=None)
    p.add_argument("--static_vae_ckpt", type=str, default=None)
    p.add_argument("--in_the_wild", action="store_true")
    p.add_argument("--num_cameras", type=int, default=1, help="orbit cameras per frame (the reference loop uses 128)")
    return p


def main(argv=None):
    args = create_argparser().parse_args(argv)
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    cfg = DEFAULT_CFG
    if os.path.exists(args.config):
        cfg = yaml.safe_load(open(args.config))
    torch.manual_seed(args.seed)
    dit = DiT(**cfg["model"])
    if args.ckpt is not None:
        dit.load_state_dict(strip_module(torch.load(args.ckpt, map_location="cpu")))
    vcfg = dict(cfg["motion_vae"])
    vcfg["num_timesteps"] = args.num_timesteps
    vae = GSKLTemporalVariationalAutoEncoder(**vcfg)
    if args.vae_ckpt is not None:
        vae.load_state_dict(strip_module(torch.load(args.vae_ckpt, map_location="cpu")))
    if args.ckpt is None or args.vae_ckpt is None:       # no checkpoints: make the zero-initialised layers non-trivial
        g = torch.Generator().manual_seed(args.seed + 1)
        for m in (dit, vae):
            for p in m.parameters():
                if p.abs().sum() == 0:
                    p.data = torch.randn(p.shape, generator=g) * 0.02
    dit, vae = dit.to(dev).eval(), vae.to(dev).eval()
    load = lambda f, d: torch.load(f).to(torch.float32).to(dev) if f is not None else d
    static_mean, static_std = load(args.static_mean_file, None), load(args.static_std_file, None)
    d_mean, d_std = load(args.deformation_mean_file, None), load(args.deformation_std_file, None)
    pipe = GVFPipeline(dit, vae, cosine_betas(1000), device=dev, resolution=512,
                       num_latents=cfg["motion_vae"]["num_latents"])
    os.makedirs(args.exp_name, exist_ok=True)
    T = args.num_timesteps
    intr = synthetic.intrinsics()
    for i in parallel.object_shard(args.num_samples, rank, world):
        f = os.path.join(args.data_dir, f"{args.start_idx + i:06d}.pt")
        if os.path.exists(f):
            d = torch.load(f, map_location="cpu")
            canon, cond = d["gaussian"], d["cond_images"][None].float()
        else:
            canon = synthetic.canonical_gaussians(seed=args.seed + i)
            cond = synthetic.sampler_inputs(1, T, cfg["model"]["resolution"], cfg["model"]["in_channels"], seed=args.seed + i)["cond_images"]
        obj = pipe.prepare_object({k: v.to(dev) for k, v in canon.items()})
        if static_mean is not None or static_std is not None:
            from gvfdiffusion_b200 import ops
            inv = (1.0 / static_std) if static_std is not None else None
            off = (-(static_mean if static_mean is not None else 0) * (inv if inv is not None else 1.0))
            obj.fps4096 = ops.affine_lastdim(obj.fps4096.contiguous(), a=inv, b=off if torch.is_tensor(off) else None)
        noise = torch.randn((1, T, cfg["model"]["resolution"], cfg["model"]["in_channels"]),
                            generator=torch.Generator().manual_seed(args.seed + i)).to(dev)
        dit.reset_conditioning()
        lat = pipe.sample(obj, cond.to(dev), noise, steps=args.rescale_timesteps, guidance_scale=args.guidance_scale,
                          guidance_scale2=args.guidance_scale2, adaptive=args.adaptive)
        delta = pipe.decode(lat, obj, d_mean, d_std)
        if args.num_cameras > 1:
            # the reference's visualisation loop (utils/inference_utils.py:243-283): every timestep from every orbit
            # camera, clamp * 255 -> uint8 on the device -> (T, cams, H, W, 3)
            out = pipe.render_views(obj, delta, synthetic.orbit_extrinsics(args.num_cameras), intr)
        else:
            out = R.rgba_to_u8(pipe.render(obj, delta, synthetic.orbit_extrinsics(T), intr))[:, None]
        torch.save(out.cpu(), os.path.join(args.exp_name, f"rank_{rank:02d}_rgb_{i:06d}.pt"))
        rgba = out
        print(f"[rank {rank}] object {i}: {tuple(rgba.shape)} frames written")


if __name__ == "__main__":
    main()
