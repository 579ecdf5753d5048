#!/usr/bin/env python
"""bench.py -- 4D frames/sec of the GVFDiffusion sampling + decode + render hot path.

    python bench.py --gpus N --steps K --warmup W           (torchrun for N > 1)
    python bench.py --impl reference ...                     (CPU oracle port of the same path)

One "step" = one object end to end: 32-step DPM-Solver++(2M) over the 12-block DiT
(B=1, T=24 frames, 512 latent tokens, 1370 image / 4096 static context tokens), motion-VAE
decode to 16 384 x 14 deltas per frame, and rasterisation of the 24 frames at 512 x 512.
value = objects * 24 frames / time, whole job over all ranks (one object per rank and step:
weak scaling, no data-path collective).  Synthetic inputs and random-init weights of the
reference architecture (no network for checkpoints), see gvfdiffusion_b200/synthetic.py.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_FRAMES, N_LAT, C_LAT, L_IMG, C_IMG, N_STATIC, VOXELS, RES, NFE = 24, 512, 16, 1370, 1024, 4096, 2048, 512, 32
DIT_CFG = dict(resolution=512, in_channels=16, model_channels=512, static_cond_channels=14, image_cond_channels=1024,
               out_channels=16, num_blocks=12, num_heads=16, mlp_ratio=4, pe_mode="ape", qk_rms_norm=True,
               use_fp16=True, no_temporal_attn=False)
VAE_CFG = dict(depth=12, dim=768, queries_dim=768, output_dim=14, num_inputs=8192, num_latents=512, latent_dim=16,
               heads=12, dim_head=-1, weight_tie_layers=False, decoder_ff=False, enable_flash_attn=True,
               num_timesteps=T_FRAMES)
WORKLOAD = ("inference_dpm_latent 32-step DPM-Solver++(2M), 1 object/GPU, 24f x 512^2, 16384 Gaussians "
            "(BASELINE.json configs[1])")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, False, []

    def _run_nvml(self):
        """In-process NVML sampling (nvidia_ml_py): ~0.1 ms per sample with the GIL released.  A `nvidia-smi` process per
        sample (the fallback below) initialises NVML every time and holds driver locks long enough to slow a launch-heavy
        step: the joint train step measured 129 ms / step under it against 88 ms without."""
        import pynvml as N
        N.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[self.index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else self.index
        h = N.nvmlDeviceGetHandleByIndex(idx)
        mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
        bits = {"hw_slowdown": N.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": N.nvmlClocksEventReasonHwThermalSlowdown,
                "sw_thermal_slowdown": N.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": N.nvmlClocksEventReasonSwPowerCap}
        while not self.stop_flag:
            r = N.nvmlDeviceGetCurrentClocksEventReasons(h)
            self.rows.append([str(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)), str(mx)] +
                             ["Active" if r & b else "Not Active" for b in bits.values()])
            time.sleep(0.1)

    def run(self):
        if os.environ.get("GVF_BENCH_NO_CLOCKS"):          # A/B switch for the sampler's own cost
            return
        try:
            return self._run_nvml()
        except Exception:
            self.rows = []
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = max((int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()), default=0)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons,
                "samples": len(sm)}


def build_models(device, seed=0):
    from gvfdiffusion_b200.model.autoencoder import GSKLTemporalVariationalAutoEncoder
    from gvfdiffusion_b200.model.dit import DiT
    torch.manual_seed(seed)
    dit, vae = DiT(**DIT_CFG), GSKLTemporalVariationalAutoEncoder(**VAE_CFG)
    g = torch.Generator().manual_seed(seed + 1)
    for m, std in ((dit, 0.02), (vae, 0.02)):          # zero-initialised layers re-drawn
        for p in m.parameters():
            if p.abs().sum() == 0:
                p.data = torch.randn(p.shape, generator=g) * std
    return dit.to(device).eval(), vae.to(device).eval()


def host_inputs(seed):
    from gvfdiffusion_b200 import synthetic as S
    canon = S.canonical_gaussians(num_voxels=VOXELS, seed=seed)
    si = S.sampler_inputs(1, T_FRAMES, N_LAT, C_LAT, L_IMG, C_IMG, seed=seed)
    pin = lambda t: t.contiguous().pin_memory()
    return {"canon": {k: pin(v) for k, v in canon.items()}, "noise": pin(si["noise"]),
            "cond_images": pin(si["cond_images"]), "ext": S.orbit_extrinsics(T_FRAMES), "intr": S.intrinsics()}


def reference_betas():
    import numpy as np
    import math
    f = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
    b = np.array([min(1 - f((i + 1) / 1000) / f(i / 1000), 0.999) for i in range(1000)], dtype=np.float64)
    ac = np.cumprod(1.0 - b)
    out, last = [], 1.0
    for a in ac:
        out.append(1 - a / last)
        last = a
    return torch.from_numpy(np.array(out, dtype=np.float64))


# ------------------------------------------------------------------------------------------
def cpu_sample(threads=None):
    """ONE measured sample of the oracle port on the host cores, nothing sliced: a full NFE (12 DiT blocks at the
    full shape incl. the image / static projections the reference recomputes every call), the full motion-VAE
    decode (12 latent layers, the decoder cross-attention of all 24 frames x 16 384 queries) and all 24 raster
    frames.  Only the factor 32 over identical NFEs is extrapolated.  -> (t_nfe, t_decode, t_raster) seconds."""
    import numpy as np
    from gvfdiffusion_b200 import synthetic as S
    from oracle import dit as ODIT, gaussian as OG, raster as OR, vae as OVAE
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    dit, vae = build_models("cpu", seed=0)
    sd, vsd = dit.state_dict(), vae.state_dict()
    si = S.sampler_inputs(1, T_FRAMES, N_LAT, C_LAT, L_IMG, C_IMG, seed=0)
    canon = S.canonical_gaussians(num_voxels=VOXELS, seed=0)
    g = torch.Generator().manual_seed(0)
    static_latent = torch.randn(1, N_STATIC, 14, generator=g)
    xyz = torch.rand(1, N_LAT, 3, generator=g) - 0.5
    t0 = time.perf_counter()
    with torch.no_grad():
        ODIT.dit_forward(sd, si["noise"], torch.tensor([500.0]), si["cond_images"], static_latent, xyz,
                         DIT_CFG["num_heads"], "fp32")
    t_nfe = time.perf_counter() - t0
    queries = torch.randn(1, VOXELS * 8, 14, generator=g)
    z = torch.randn(T_FRAMES, N_LAT, C_LAT, generator=g)
    t0 = time.perf_counter()
    with torch.no_grad():
        OVAE.vae_decode(vsd, z, queries, VAE_CFG["heads"], T_FRAMES, "fp32")
    t_dec = time.perf_counter() - t0
    delta = S.raster_delta(T_FRAMES, VOXELS * 8).numpy()
    ext, intr, const = S.orbit_extrinsics(T_FRAMES), S.intrinsics(), S.gaussian_constants()
    vt, pt = [], []
    for f in range(T_FRAMES):
        v, p_, _, tfx, tfy = OG.camera_matrices(ext[f], intr, 0.8, 1.6)
        vt.append(v.numpy())
        pt.append(p_.numpy())
    prm = OR.make_params(RES, RES, tfx, tfy, const)
    t0 = time.perf_counter()
    OR.render_frames(prm, {k: v.numpy() for k, v in canon.items()}, delta, np.stack(vt), np.stack(pt))
    t_r = time.perf_counter() - t0
    return t_nfe, t_dec, t_r


def cpu_baseline(threads=None, samples=1):
    """Median of `samples` cpu_sample() runs -> the cpu_baseline object of the bench line."""
    threads = threads or os.cpu_count()
    runs = sorted((cpu_sample(threads) for _ in range(samples)), key=lambda r: NFE * r[0] + r[1] + r[2])
    t_nfe, t_dec, t_r = runs[len(runs) // 2]
    total = NFE * t_nfe + t_dec + t_r
    return {"value": T_FRAMES / total, "unit": "frames/s", "cores": threads, "kind": "port",
            "extrapolated": {"nfe": NFE}, "samples": samples, "measured_s": {"one_nfe": t_nfe, "vae_decode": t_dec,
                                                                             "raster_24f": t_r},
            "seconds_per_object": total,
            "sample": (f"measured on the host ({threads} threads), median of {samples}: ONE full NFE of the 12-block DiT at the "
                       f"full shape ({t_nfe:.2f}s), the full 12-layer motion-VAE decode with all 24 x 16384 decoder queries "
                       f"({t_dec:.2f}s), all 24 raster frames ({t_r:.2f}s); only x{NFE} over identical NFEs is extrapolated "
                       f"-> {total:.0f}s/object; oracle port (torch fp32 + C rasteriser; the reference has no CPU rasteriser)")}


def run_reference(args, rank):
    """`--impl reference`: the CPU arm.  Every timed "step" is one bounded sample (cpu_sample: ~10 s of host work);
    at most three are run whatever --steps says, the median is reported."""
    if rank != 0:
        return
    if args.warmup > 0:
        cpu_sample()                                           # page the oracle / MKL in
    cb = cpu_baseline(samples=max(1, min(args.steps, 3)))
    line = {"impl": "reference", "metric": "4D frames/sec (32-step DPM, 24f x 512^2, 16k Gaussians)",
            "value": cb["value"], "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * cb["seconds_per_object"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD},
            "samples_timed": cb["samples"], "extrapolated": cb["extrapolated"],
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "frames/s", "h2d_bytes_per_step": 0,
                                        "d2h_bytes_per_step": 0}}
    args._out.write(json.dumps(line) + "\n")
    args._out.flush()


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true",
                    help="skip the GPU stand-in of the reference's own execution (tools/gpu_reference.py, N = 1 only)")
    ap.add_argument("--scatter", action="store_true", help="(default for N > 1; kept for explicit command lines)")
    ap.add_argument("--no-scatter", action="store_true",
                    help="N > 1: by default e2e goes through the real data-parallel path (SURVEY 8e): rank 0 owns all objects "
                         "in pinned host memory, uploads and sends every rank its conditioning each step and collects the "
                         "frames (NCCL point-to-point, overlapped with compute: gvfdiffusion_b200.parallel.PipelinedExchange); "
                         "this flag makes every rank copy its own object instead (no collective on the data path)")
    ap.add_argument("--prefetch", action="store_true",
                    help="A/B: issue sample_gs (farthest point sampling) of the next object on a side stream next to this "
                         "object's sampling.  MEASURED on B200: resident 272.1 vs 273.4 ms / object (+0.5 %: the persistent "
                         "GEMMs lose the SM the single-CTA FPS sits on), end to end 83.1 vs 88.0 frames/s -- default off")
    ap.add_argument("--attn-dbg", type=lambda v: int(v, 0), default=0, help="gvf_attn_set_debug value (kernel-variant A/B)")
    ap.add_argument("--pdl", action="store_true", help="launch with the programmatic-dependent-launch attribute (A/B; default off)")
    ap.add_argument("--sampler-graph", type=int, default=None, help="1 / 0: the whole fixed-step sampling run as one CUDA graph "
                    "(GVFPipeline.sampler_graph) on / off for A/B runs; default = the pipeline's default")
    ap.add_argument("--config", default="cfg1", choices=["cfg1", "cfg3", "cfg5", "cfg5-static"],
                    help="cfg1 (default): BASELINE.json configs[1], the headline inference metric.  cfg3: configs[2], one "
                         "training step of the motion-VAE decoder + 24-frame render, forward + backward (tools/train_step_bench.py).  "
                         "cfg5: configs[4], the joint main_vae.py train step with optimiser, DDP all-reduce for N > 1 "
                         "(tools/cfg5_step_bench.py).  cfg5-static: its static-VAE half next to a flash-attn / cuBLAS / autograd "
                         "stand-in (tools/static_vae_step_bench.py, N = 1).  Each prints its own JSON line")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    args.scatter = world > 1 and not args.no_scatter
    # stdout carries exactly ONE JSON line: native libraries (NCCL prints its version banner there) write to fd 1,
    # so fd 1 is pointed at stderr for the duration of the run and the line goes to the saved descriptor
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    args._out = real_stdout
    if args.impl == "reference":
        return run_reference(args, rank)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    if args.config in ("cfg5", "cfg5-static"):
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        mon = ClockSampler(local)
        mon.start()
        if args.config == "cfg5":
            from tools import cfg5_step_bench as C5
            res = C5.measure(steps=args.steps, warmup=max(args.warmup, 3), seed=rank, world=world, device=dev,
                             standin=(not args.no_gpu_reference) and world == 1)
        else:
            from tools import static_vae_step_bench as SVB
            res = SVB.measure(steps=args.steps, warmup=max(args.warmup, 3), standin=not args.no_gpu_reference, seed=rank, device=dev)
            res.update({"n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "higher_is_better": True,
                        "scaling": "weak", "vs_baseline": None})
        mon.stop_flag = True
        mon.join(timeout=2)
        if rank == 0:
            res["clocks"] = mon.summary()
            real_stdout.write(json.dumps(res) + "\n")
            real_stdout.flush()
        return
    if args.config == "cfg3":
        # BASELINE configs[2] (and, for N > 1, the data-parallel half of configs[4]: every rank steps its own object and the
        # fp32 parameter gradients are averaged with one NCCL all-reduce inside the timed step)
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools"))
        import train_step_bench as TSB
        mon = ClockSampler(local)
        mon.start()
        res = TSB.measure(steps=args.steps, warmup=max(args.warmup, 3), standin=(not args.no_gpu_reference) and world == 1,
                          seed=rank, world=world, device=dev)
        mon.stop_flag = True
        mon.join(timeout=2)
        if rank == 0:
            res.update({"n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "higher_is_better": True,
                        "scaling": "weak", "vs_baseline": None, "clocks": mon.summary()})
            real_stdout.write(json.dumps(res) + "\n")
            real_stdout.flush()
        return
    from gvfdiffusion_b200 import _lib, ops
    from gvfdiffusion_b200.pipeline import GVFPipeline
    if args.pdl:
        _lib.lib().gvf_set_pdl(1)
    if args.attn_dbg:
        _lib.lib().gvf_attn_set_debug(args.attn_dbg)

    dit, vae = build_models(dev, seed=0)                      # replicated weights
    pipe = GVFPipeline(dit, vae, reference_betas(), device=dev, resolution=RES)
    if args.sampler_graph is not None:
        pipe.sampler_graph = bool(args.sampler_graph)
    hin = host_inputs(seed=rank)                              # a different object per rank
    out_host = torch.empty((T_FRAMES, 4, RES, RES), dtype=torch.float32).pin_memory()
    out_dev = torch.empty((T_FRAMES, 4, RES, RES), dtype=torch.float32, device=dev)

    # resident inputs for the device-timed run
    canon_d = {k: v.to(dev, non_blocking=True) for k, v in hin["canon"].items()}
    noise_d, cond_d = hin["noise"].to(dev), hin["cond_images"].to(dev)
    obj = pipe.prepare_object(canon_d)

    timer = ops.LaunchTimer()
    launches = [0]

    prefetch = args.prefetch
    nxt = {"resident": None, "e2e": None}

    def step_resident():
        dit.reset_conditioning()                              # per-object projections are part of the step
        # ... and so is sample_gs (farthest point sampling): one prepare_object per step.  Steady-state loop over
        # objects: the preparation of the NEXT object is issued on a side stream at the start of this step (its
        # single-CTA FPS then runs next to this object's sampling), this step consumes the one issued a step ago.
        if prefetch:
            o = pipe.wait_object(nxt["resident"] or pipe.prepare_object_async(canon_d))
            nxt["resident"] = pipe.prepare_object_async(canon_d)
        else:
            o = pipe.prepare_object(canon_d)
        lat = pipe.sample(o, cond_d, noise_d, steps=NFE)
        delta = pipe.decode(lat, o)
        pipe.render(o, delta, hin["ext"], hin["intr"], out=out_dev, check_overflow="defer")

    def step_profiled():
        """One more step with eager launches (no graph replay) and CUDA events around every attention
        launch and around the three stages: the per-kernel numbers behind `roofline`."""
        eng = dit.engine()
        orig, eng.use_graphs = ops.attention, False
        ops.attention = _tagged_attention(orig, timer)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        try:
            dit.reset_conditioning()
            ev[4].record()
            pipe.prepare_object(canon_d)
            ev[0].record()
            lat = pipe.sample(obj, cond_d, noise_d, steps=NFE)
            ev[1].record()
            delta = pipe.decode(lat, obj)
            ev[2].record()
            pipe.render(obj, delta, hin["ext"], hin["intr"], out=out_dev)
            ev[3].record()
        finally:
            ops.attention, eng.use_graphs = orig, True
        torch.cuda.synchronize()
        return [ev[i].elapsed_time(ev[i + 1]) for i in range(3)] + [ev[4].elapsed_time(ev[0])]

    # End to end: every step uploads its inputs from pinned host memory and reads its RGBA back.  The copies run
    # on a second stream, double buffered: the upload of step k+1 and the read-back of step k-1 overlap the
    # compute of step k (what a serving loop does); both still happen once per step inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    slots = []
    for _ in range(2):
        slots.append({
            "canon": {k: torch.empty_like(v, device=dev) for k, v in hin["canon"].items()},
            "noise": torch.empty_like(hin["noise"], device=dev), "cond": torch.empty_like(hin["cond_images"], device=dev),
            "out_dev": torch.empty((T_FRAMES, 4, RES, RES), dtype=torch.float32, device=dev),
            "out_host": torch.empty((T_FRAMES, 4, RES, RES), dtype=torch.float32).pin_memory(),
            "ready": torch.cuda.Event(), "free": torch.cuda.Event(), "rendered": torch.cuda.Event()})
    e2e_state = {"k": 0, "primed": False}

    def upload(slot):
        st = slots[slot]
        copy_stream.wait_event(st["free"])                     # the compute that last read these buffers is done
        with torch.cuda.stream(copy_stream):
            for k, v in hin["canon"].items():
                st["canon"][k].copy_(v, non_blocking=True)
            st["noise"].copy_(hin["noise"], non_blocking=True)
            st["cond"].copy_(hin["cond_images"], non_blocking=True)
            st["ready"].record(copy_stream)

    def step_e2e():
        cur = torch.cuda.current_stream()
        if not e2e_state["primed"]:                            # very first call: nothing has been prefetched yet
            for st in slots:
                st["free"].record(cur)
            upload(0)
            e2e_state["primed"] = True
        slot = e2e_state["k"] & 1
        e2e_state["k"] += 1
        st = slots[slot]
        cur.wait_event(st["ready"])
        upload(slot ^ 1)                                       # inputs of the NEXT step travel during this one
        dit.reset_conditioning()
        if prefetch:                                           # ... and are prepared (FPS) next to this step's sampling
            o = pipe.wait_object(nxt["e2e"] or pipe.prepare_object_async(st["canon"], after=st["ready"]))
            nxt["e2e"] = pipe.prepare_object_async(slots[slot ^ 1]["canon"], after=slots[slot ^ 1]["ready"])
        else:
            o = pipe.prepare_object(st["canon"])
        lat = pipe.sample(o, st["cond"], st["noise"], steps=NFE)
        delta = pipe.decode(lat, o)
        pipe.render(o, delta, hin["ext"], hin["intr"], out=st["out_dev"], check_overflow="defer")
        st["free"].record(cur)
        st["rendered"].record(cur)
        copy_stream.wait_event(st["rendered"])
        with torch.cuda.stream(copy_stream):
            st["out_host"].copy_(st["out_dev"], non_blocking=True)
            st["free"].record(copy_stream)                     # out_dev may be overwritten once it has been read back

    # ---- N > 1 with --scatter: the e2e number goes through rank 0 (scatter of conditioning, gather of frames)
    ex = None
    if world > 1 and args.scatter:
        from gvfdiffusion_b200.parallel import PipelinedExchange
        specs = {("canon." + k): (tuple(v.shape), v.dtype) for k, v in hin["canon"].items()}
        specs["noise"] = (tuple(hin["noise"].shape), hin["noise"].dtype)
        specs["cond_images"] = (tuple(hin["cond_images"].shape), hin["cond_images"].dtype)
        ex = PipelinedExchange(specs, (T_FRAMES, 4, RES, RES), dev)
        root_objects = None
        if rank == 0:                                          # one distinct object per rank, all owned by rank 0
            root_objects = []
            for r in range(world):
                h = hin if r == 0 else host_inputs(seed=r)
                d = {("canon." + k): v for k, v in h["canon"].items()}
                d["noise"], d["cond_images"] = h["noise"], h["cond_images"]
                root_objects.append(d)
        sc_state = {"k": 0}

        def step_scatter():
            k = sc_state["k"]
            sc_state["k"] += 1
            ex.post(k, root_objects)
            inp, out = ex.inputs(k), ex.output(k)
            dit.reset_conditioning()
            o = pipe.prepare_object({n[6:]: t for n, t in inp.items() if n.startswith("canon.")})
            lat = pipe.sample(o, inp["cond_images"], inp["noise"], steps=NFE)
            delta = pipe.decode(lat, o)
            pipe.render(o, delta, hin["ext"], hin["intr"], out=out, check_overflow="defer")
            ex.done(k)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        # the last step's read-back runs on a side stream: it belongs inside the interval
        if ex is not None and fn is step_scatter:
            ex.flush(sc_state["k"])
        torch.cuda.current_stream().wait_stream(copy_stream)
        e1.record()
        barrier()
        pipe.confirm_render()                                  # raises if any timed render dropped splats
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local)
    sampler.start()
    ms_total = timed(step_resident, args.steps)
    sampler.stop_flag = True
    sampler.join()
    Rn, overflow, _ = pipe.rz.status()
    if overflow:
        raise SystemExit("rasteriser tile-instance capacity overflowed: result invalid")
    scatter_info = None
    if ex is not None:
        ex.prime(root_objects)
        step_scatter()                                          # untimed: fills the pipeline
        ex.flush(sc_state["k"])
        b0 = (ex.bytes_h2d, ex.bytes_d2h, ex.bytes_p2p)
        ms_e2e = timed(step_scatter, args.steps)
        per = [(b - a) / args.steps for a, b in zip(b0, (ex.bytes_h2d, ex.bytes_d2h, ex.bytes_p2p))]
        # the exchange alone (no compute between the posts): what one step's scatter + gather costs on the side stream
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(3):
            k = sc_state["k"]
            sc_state["k"] += 1
            ex.post(k, root_objects)
            ex.inputs(k)
            ex.done(k)
        ex.flush(sc_state["k"])
        c1.record()
        torch.cuda.synchronize()
        last = ex.host_results[0] if rank == 0 else None
        e2e_max_diff = float((last - out_dev.cpu()).abs().max()) if rank == 0 else 0.0
        scatter_info = {"root_h2d_bytes_per_step": per[0], "root_d2h_bytes_per_step": per[1],
                        "nccl_p2p_bytes_per_step_at_root": per[2], "exchange_alone_ms_per_step": c0.elapsed_time(c1) / 3,
                        "path": "rank 0 pinned host -> H2D -> NCCL isend/irecv (one batched group per step, side stream) -> "
                                "sample/decode/render -> NCCL -> rank 0 -> D2H; scatter of step k+1 and gather of step k-1 "
                                "overlap compute of step k"}
    else:
        step_e2e()
        ms_e2e = timed(step_e2e, args.steps)
        # the frames that reached the host through the overlapped path are the frames of the resident run
        last = slots[(e2e_state["k"] - 1) & 1]["out_host"]
        e2e_max_diff = float((last - out_dev.cpu()).abs().max())
    stage_ms = step_profiled()
    # rasteriser alone (24 frames), graph-free, for the HBM-side roofline
    torch.cuda.synchronize()
    lat = pipe.sample(obj, cond_d, noise_d, steps=2)
    delta = pipe.decode(lat, obj)
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record()
    for _ in range(5):
        pipe.render(obj, delta, hin["ext"], hin["intr"], out=out_dev)
    r1.record()
    torch.cuda.synchronize()
    raster_ms = r0.elapsed_time(r1) / 5
    # BASELINE.json configs[2] (render forward + backward): gradients of the 24 frames to the raw Gaussian
    # parameters and delta (gvf_raster_backward: blend backward + preprocess backward)
    from gvfdiffusion_b200 import raster as R
    cams_b, tfx_b, tfy_b = R.pack_cameras(hin["ext"], hin["intr"], pipe.near, pipe.far)
    prm_b = R.make_params(pipe.res, pipe.res, tfx_b, tfy_b, pipe.const, pipe.kernel_size, 1.0, pipe.bg)
    cams_b, delta_b = cams_b.to(dev), delta.contiguous()
    rz_b = R.Rasterizer(dev)
    rz_b.forward(prm_b, obj.arrays, delta_b, cams_b, want_radii=False)
    grad_b = torch.randn(out_dev.shape, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    rz_b.backward(prm_b, obj.arrays, delta_b, cams_b, grad_b)
    r0.record()
    for _ in range(5):
        rz_b.backward(prm_b, obj.arrays, delta_b, cams_b, grad_b)
    r1.record()
    torch.cuda.synchronize()
    raster_bwd_ms = r0.elapsed_time(r1) / 5
    del rz_b, grad_b
    # the reference's visualisation loop (utils/inference_utils.py:243-283): every timestep from 128 orbit
    # cameras, clamped and converted to uint8 on the device (24 x 128 renders, one rasteriser call per timestep)
    from gvfdiffusion_b200 import synthetic as S
    views_ext = S.orbit_extrinsics(128)
    u8 = pipe.render_views(obj, delta, views_ext, hin["intr"])
    r0.record()
    pipe.render_views(obj, delta, views_ext, hin["intr"], out=u8)
    r1.record()
    torch.cuda.synchronize()
    views_ms = r0.elapsed_time(r1)
    del u8

    if rank != 0:
        return
    pk = peaks()
    ms_step = ms_total / args.steps
    value = world * T_FRAMES / (ms_step / 1e3)
    e2e_val = world * T_FRAMES / (ms_e2e / args.steps / 1e3)
    rec = timer.summary()
    roof_detail = {}
    for tag, (n, ms) in rec.items():
        fl = ATTN_FLOPS.get(tag)
        if fl:
            roof_detail[tag] = {"launches": n, "avg_ms": ms / n, "tflops": fl / (ms / n) / 1e9,
                                "frac_of_sustained": fl / (ms / n) / 1e9 / pk["tf_sustained"]}
    dom = roof_detail.get("attn_static", {})
    h2d = sum(v.numel() * 4 for v in hin["canon"].values()) + hin["noise"].numel() * 4 + hin["cond_images"].numel() * 4
    line = {
        "metric": "4D frames/sec (32-step DPM, 24f x 512^2, 16k Gaussians)", "value": value, "unit": "frames/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16 (fp32 accumulate/state)",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "objects_per_step": world, "nfe": NFE, "guidance": "1.0/1.0 (1 branch)",
                   "l2": "inputs larger than L2 (808 MB hoisted image K/V + 1.2 GB activations per step; no flush)",
                   "num_rendered": Rn,
                   "modulation": "timestep MLP + adaLN vectors of the run's 32 model times computed once PER OBJECT inside the "
                                 "timed region (4 batched launch pairs), not once per NFE; nothing is carried between objects",
                   "sampling": ("the 32 NFEs + solver updates of an object replayed as one CUDA graph" if getattr(pipe, "sampler_graph", False)
                                else "one CUDA graph per NFE"),
                   "object_prefetch": ("sample_gs (FPS) of object k+1 on a side stream during the sampling of object k; "
                                       "one prepare_object per step" if prefetch else "off")},
        "clocks": sampler.summary(),
        "e2e": {"value": e2e_val, "unit": "frames/s",
                "h2d_bytes_per_step": scatter_info["root_h2d_bytes_per_step"] if scatter_info else h2d,
                "d2h_bytes_per_step": scatter_info["root_d2h_bytes_per_step"] if scatter_info else out_host.numel() * 4,
                "max_abs_diff_vs_resident_rgba": e2e_max_diff,
                "note": (scatter_info["path"] if scatter_info else
                         "copies on a second stream, double buffered: upload of step k+1 / read-back of step k-1 overlap step k")},
        "gpu_launches": launch_estimate() * args.steps,
        "roofline": {"bound": "tensor", "kernel": "attn_fwd8_kernel (d=32, static cross-attention, kv 4096)",
                     "achieved": dom.get("tflops"), "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                     "frac": dom.get("frac_of_sustained"), "traffic": ncu_traffic("attn_fwd8_kernel static"),
                     "peak_source": pk["source"] + " sustained bf16",
                     "mufu_ceiling_tflops": 148 * 16 * 128 * (sampler.summary()["sm_mhz"] or 1965) * 1e6 / 1e12,
                     "note": ("head dim 32: every score costs one exponential against 128 tensor flops; with MUFU alone (16 / clk / SM) "
                              "the ceiling is mufu_ceiling_tflops (596 at 1965 MHz = 43 % of the measured cuBLAS peak).  v8 runs 1/4 "
                              "of the exponentials on the FMA pipe (ceiling 794) and issues tcgen05.mma from warp-uniform code; the "
                              "measured balance point of the loop is 12.3 clk per pair of scores and sub-partition "
                              "(tools/probe/pipe_probe.cu) = 770 TFLOP/s, the kernel reaches 59 % of that -- the rest is the per-block "
                              "TMEM load / store / mbarrier hand-off (~840 of 2130 clk per 64-key block, tools/attn_trace.py); ncu: "
                              "XU 68 %, issue 65 %, tensor pipe 21 % (profiles/r02_attn8_full_extract.csv)")},
        "roofline_detail": roof_detail,
        "stage_ms_eager": {"prepare_fps": stage_ms[3], "sample_32nfe": stage_ms[0], "vae_decode": stage_ms[1],
                           "raster_24f": stage_ms[2], "raster_24f_backward": raster_bwd_ms,
                           "raster_24x128_views_u8": views_ms},
        "roofline_raster": {"bound": "hbm", "kernel": "gvf_raster_forward (4 kernels, 24 frames)",
                            "achieved": (T_FRAMES * (112 * VOXELS * 8 + 16 * RES * RES) + 64 * Rn) / raster_ms / 1e6,
                            "peak": pk["hbm_gbs"], "unit": "GB/s",
                            "frac": (T_FRAMES * (112 * VOXELS * 8 + 16 * RES * RES) + 64 * Rn) / raster_ms / 1e6 / pk["hbm_gbs"],
                            "ms": raster_ms, "traffic": ncu_traffic("gvf_raster_forward 24 frames"),
                            "note": ("not HBM-bound at this depth complexity: the tile lists hold 460-790 splats in the scene's "
                                     "central tiles; sort_blend (75 % of the call) executes 237 M warp instructions (bitonic sort: "
                                     "310 M, before sub-tile culling: 671 M) at 64 % issue utilisation, two thirds of them in the "
                                     "blend loop, a fifth in the bucket sort; DRAM traffic is below the algorithmic bytes because "
                                     "key lists and splat records are re-read from L2 (profiles/r01_raster_bucket_full_extract.csv)")},
    }
    if scatter_info:
        line["scatter_gather"] = scatter_info
    if world == 1 and not args.no_gpu_reference:
        try:        # the reference's own execution restated on this GPU (stand-in; see tools/gpu_reference.py)
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import gpu_reference as GR
            gr, rgba_ref, _, lat_ref = GR.measure(dit, vae, pipe, canon_d, cond_d, noise_d, hin["ext"], hin["intr"],
                                                  steps=NFE, reps=2)
            dit.reset_conditioning()
            lat_p = pipe.sample(obj, cond_d, noise_d, steps=NFE)
            rgba_p = pipe.render(obj, pipe.decode(lat_p, obj), hin["ext"], hin["intr"])
            rel = lambda x, y: float((x - y).norm() / y.norm())
            gr["product_vs_standin"] = {"latent_rel_l2": rel(lat_p, lat_ref), "rgba_rel_l2": rel(rgba_p, rgba_ref),
                                        "rgba_max_abs": float((rgba_p - rgba_ref).abs().max())}
            gr["speedup_e2e"] = e2e_val / gr["value"]
            gr["speedup_resident"] = value / gr["value"]
            line["gpu_reference"] = gr
        except Exception as e:
            line["gpu_reference"] = {"error": repr(e)}
    try:        # derived figure only: never allowed to break the line
        line["roofline_dit"] = dit_roofline(ms_step, line["stage_ms_eager"], pk["tf_sustained"], NFE)
    except Exception as e:
        line["roofline_dit"] = {"error": str(e)}
    if not args.no_cpu_baseline and world == 1:
        try:
            line["cpu_baseline"] = cpu_baseline(samples=1)
        except Exception as e:   # the oracle is a checker; never let it take the GPU number down
            line["cpu_baseline"] = {"error": repr(e)}
    args._out.write(json.dumps(line) + "\n")
    args._out.flush()


# attention FLOPs per launch (4 * Nb * H * Lq * Lk * d), tagged by shape
ATTN_FLOPS = {
    "attn_static": 4 * T_FRAMES * 16 * N_LAT * N_STATIC * 32,
    "attn_image": 4 * T_FRAMES * 16 * N_LAT * L_IMG * 32,
    "attn_spatial": 4 * T_FRAMES * 16 * N_LAT * N_LAT * 32,
    "attn_temporal": 4 * N_LAT * 16 * T_FRAMES * T_FRAMES * 32,
    "attn_vae_self": 4 * T_FRAMES * 12 * N_LAT * N_LAT * 64,
    "attn_vae_dec": 4 * T_FRAMES * 12 * 8192 * N_LAT * 64,
}


def _tagged_attention(fn, timer):
    """CUDA events around every attention launch of the timed region, bucketed by shape."""
    def call(q, k, v, scale, out=None, q_shared=False, kv_shared=False):
        D, Lq = q.shape[-1], q.shape[-3]
        Lk = k.shape[-3]
        tag = ("attn_static" if kv_shared else "attn_vae_dec" if q_shared else "attn_vae_self" if D == 64 else
               "attn_temporal" if Lq <= 32 else "attn_image" if Lk == L_IMG else "attn_spatial")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(q, k, v, scale, out=out, q_shared=q_shared, kv_shared=kv_shared)
        e1.record()
        timer.records.setdefault(tag, []).append((e0, e1))
        return r
    return call



def dit_roofline(ms_step, stage_ms_eager, peak_tflops, nfe=32):
    """Tensor roofline of the whole sampler: algorithmic F_NFE = 3.394 TFLOP per NFE + F_once =
    0.465 TFLOP per object (time- and frame-invariant projections hoisted) over the sampler's share of the timed
    step = ms_step minus the eagerly timed non-sampler stages (FPS, VAE decode, render)."""
    other = sum(float(stage_ms_eager.get(k, 0.0)) for k in ("prepare_fps", "vae_decode", "raster_24f"))
    t_ms = max(float(ms_step) - other, 1e-6)
    flop = 3.394e12 * nfe + 0.465e12
    ach = flop / (t_ms * 1e-3) / 1e12
    return {"bound": "tensor", "kernel": "DiT sampler (32 NFE, graph replay)", "achieved": ach, "peak": peak_tflops,
            "unit": "TFLOP/s", "frac": ach / peak_tflops, "ms": t_ms, "algorithmic_tflop": flop / 1e12}

def ncu_traffic(kernel_prefix):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/r02_traffic.json)."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        for k, v in json.load(open(p)).items():
            if k.startswith(kernel_prefix):
                return v.get("dram_bytes")
    except Exception:
        pass
    return None


def launch_estimate():
    """Kernel launches of ours per object, counted from the engine structure and checked against the committed
    ncu launch list (profiles/r01_nfe_launch_list.csv: 471 launches for 2 NFE incl. 4 torch copies, 117 for
    decode + render): per NFE 1 (input) + 12 blocks x 19 (5 LayerNorm, 2 qkv, 4 attention, 4 out-proj, 2 q-proj,
    fc1, fc2) + 1 (final) + 2 (DPM); the modulation vectors of the 32 model times are 4 batched launch pairs per object
    (they were 2 launches per NFE until the table of dit_engine.precompute_modulation)."""
    per_nfe = 1 + 12 * 19 + 1 + 2
    hoist = 2 + 12 + 1 + 12 + 1 + 3 + 2 * ((NFE + 7) // 8)
    decode_render = 112 + 5
    return NFE * per_nfe + hoist + decode_render


if __name__ == "__main__":
    main()      # no destroy_process_group(): ranks leave at different times (rank 0 prints), process exit tears NCCL down
