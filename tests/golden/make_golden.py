"""Generates tests/golden/*.pt by running the REFERENCE's own Python (imported from
/root/reference with the stubs of _ref_import.py) on seeded inputs, CPU.

Run in the build container only:   python tests/golden/make_golden.py
The fixtures are committed; tests/test_oracle_golden.py checks the oracle against them on
any machine (the reference tree does not exist on the GPU box).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_import  # noqa: E402

_ref_import.install()

from model.dit import DiT  # noqa: E402
from model import dpmsolver as ref_dpm  # noqa: E402
from model.autoencoder import GSKLTemporalVariationalAutoEncoder  # noqa: E402
from utils.script_util import create_gaussian_diffusion  # noqa: E402

DIFF_CFG = dict(steps=1000, learn_sigma=False, sigma_small=False, use_kl=False, noise_schedule="cosine",
                predict_type="v", predict_xstart=False, rescale_timesteps=True, rescale_learned_sigmas=True)
TINY_DIT = dict(resolution=32, in_channels=16, model_channels=64, static_cond_channels=14,
                image_cond_channels=48, out_channels=16, num_blocks=2, num_heads=2, mlp_ratio=4,
                pe_mode="ape", qk_rms_norm=True, use_fp16=True, no_temporal_attn=False)
TINY_VAE = dict(depth=2, dim=96, queries_dim=96, output_dim=14, num_inputs=64, num_latents=32,
                latent_dim=16, heads=3, dim_head=-1, weight_tie_layers=False, decoder_ff=False,
                enable_flash_attn=False, num_timesteps=3)


def rerandomise_zero_layers(model, std=0.02, seed=123):
    """The reference zero-initialises adaLN / final layers (model/dit.py:414-427): re-draw them
    so the golden outputs are non-trivial."""
    g = torch.Generator().manual_seed(seed)
    for p in model.parameters():
        if p.abs().sum() == 0:
            p.data = torch.randn(p.shape, generator=g) * std


def gen_schedule():
    diffusion = create_gaussian_diffusion(**DIFF_CFG)
    ns = ref_dpm.NoiseScheduleVP("discrete", betas=torch.from_numpy(diffusion.betas))
    ts = torch.tensor([1.0, 0.999, 0.75, 0.5, 0.123456, 0.01, 0.002, 0.001])
    out = {"betas": torch.from_numpy(diffusion.betas), "total_N": ns.total_N,
           "log_alpha_array": ns.log_alpha_array, "t_array": ns.t_array, "ts": ts,
           "log_alpha": torch.stack([ns.marginal_log_mean_coeff(t[None]) for t in ts]),
           "lambda": torch.stack([ns.marginal_lambda(t[None]) for t in ts]),
           "std": torch.stack([ns.marginal_std(t[None]) for t in ts])}
    lam = torch.tensor([-5.0, -2.0, 0.0, 1.5, 4.0])
    out["inv_lambda_in"] = lam
    out["inv_lambda"] = ns.inverse_lambda(lam)
    return out, diffusion, ns


def gen_dit(ns):
    torch.manual_seed(0)
    m = DiT(**TINY_DIT).eval()
    rerandomise_zero_layers(m)
    B, T, N = 2, 3, 32
    g = torch.Generator().manual_seed(7)
    x = torch.randn(B, T, N, 16, generator=g)
    t = torch.tensor([812.25, 3.5])
    cond = {"cond_images": torch.randn(B, T, 10, 48, generator=g),
            "static_latent": torch.randn(B, 40, 14, generator=g),
            "deformation_position_xyz": torch.rand(B, N, 3, generator=g) - 0.5}
    out = {"cfg": TINY_DIT, "state_dict": {k: v.clone() for k, v in m.state_dict().items()},
           "x": x, "t": t, **cond}
    with torch.no_grad():
        out["y_fp32"] = m(x, t, **cond)
        with torch.autocast("cpu", dtype=torch.float16):
            out["y_autocast_fp16"] = m(x, t, **cond).float()

        # sampler goldens (B = 1 object)
        c1 = {k: v[:1] for k, v in cond.items()}
        u1 = dict(c1)
        u1["cond_images"] = torch.zeros_like(c1["cond_images"])
        noise = torch.randn(1, T, N, 16, generator=g)
        out["noise"] = noise
        for name, gs1, gs2 in (("g1", 1.0, 1.0), ("cfg", 2.0, 1.5)):
            fn = ref_dpm.model_wrapper(m, ns, model_type="v", model_kwargs={}, guidance_type="classifier-free",
                                       guidance_scale=gs1, guidance_scale2=gs2, condition=c1,
                                       unconditional_condition=u1)
            solver = ref_dpm.DPM_Solver(fn, ns, algorithm_type="dpmsolver++")
            for steps in (6, 12):
                out[f"sample_{name}_{steps}"] = solver.sample(x=noise, steps=steps, t_start=1.0, t_end=1 / 1000,
                                                              order=2, skip_type="time_uniform", method="multistep")
        fn = ref_dpm.model_wrapper(m, ns, model_type="v", model_kwargs={}, guidance_type="classifier-free",
                                   guidance_scale=1.0, guidance_scale2=1.0, condition=c1, unconditional_condition=u1)
        solver = ref_dpm.DPM_Solver(fn, ns, algorithm_type="dpmsolver++")
        out["sample_adaptive"] = solver.sample(x=noise, steps=0, t_start=1.0, t_end=1 / 1000, order=2,
                                               skip_type="time_uniform", method="adaptive")
        out["eps_g1_t0.37"] = fn(noise, torch.tensor([0.37]))
    return out


def gen_vae():
    torch.manual_seed(1)
    vae = GSKLTemporalVariationalAutoEncoder(**TINY_VAE).eval()
    rerandomise_zero_layers(vae, std=0.05)     # to_outputs and every bias are zero-initialised (:422-436)
    g = torch.Generator().manual_seed(11)
    z = torch.randn(2 * 3, 32, 16, generator=g)
    q = torch.randn(2, 50, 14, generator=g) * 0.3
    q[1, 40:] = 0.0
    q[1, 40:, 10] = 1.0           # pad rows as pad_static_gs writes them (train_vae.py:478-479)
    keep = ("layers.", "proj.", "gs_embedding.", "decoder_cross_attn.", "to_outputs.")   # decode-only weights
    out = {"cfg": TINY_VAE, "z": z, "queries": q,
           "state_dict": {k: v.clone() for k, v in vae.state_dict().items() if k.startswith(keep)}}
    with torch.no_grad():
        out["delta_fp32"] = vae.decode(z, q)
        with torch.autocast("cpu", dtype=torch.float16):
            out["delta_autocast_fp16"] = vae.decode(z, q).float()
        vae.chunk_size = 16       # exercises the chunked path (autoencoder.py:592-607)
        out["delta_fp32_chunked"] = vae.decode(z, q)
    return out


def _fps_stub(x, batch, ratio):
    """Stand-in for torch_cluster.fps (absent here): greedy farthest point sampling per batch entry from its FIRST point
    (torch_cluster defaults to a random start), squared distances (dx*dx + dy*dy) + dz*dz in fp32, ties -> lowest index --
    the deterministic rule of gvf_fps.  Returns global row indices, entry by entry."""
    out = []
    x = x.numpy().astype(np.float32)
    for b in range(int(batch.max()) + 1):
        rows = np.nonzero(batch.numpy() == b)[0]
        p = x[rows]
        K = int(round(float(ratio[b]) * len(rows)))
        mind = np.full(len(rows), 3.0e38, np.float32)
        cur, idx = 0, [0]
        for _ in range(1, K):
            d = p - p[cur]
            dist = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
            mind = np.minimum(mind, dist.astype(np.float32))
            cur = int(np.argmax(mind))
            idx.append(cur)
        out.append(torch.from_numpy(rows[np.array(idx)]))
    return torch.cat(out)


def gen_vae_encode():
    """The reference's GSKLTemporalVariationalAutoEncoder.encode (model/autoencoder.py:502-550) on the tiny config, CPU,
    fp32 and fp16 autocast, with torch_cluster.fps / pytorch3d.ops.knn_points replaced by deterministic stand-ins of their
    documented semantics (both are absent third parties).  The posterior noise is drawn under a fixed seed and stored."""
    import model.autoencoder as AE
    import pytorch3d.ops as P3
    AE.fps = _fps_stub
    P3.knn_points = _bruteforce_knn_points
    torch.manual_seed(2)
    cfg = dict(TINY_VAE, knn_k=4, beta=7.0)
    vae = GSKLTemporalVariationalAutoEncoder(**cfg).eval()
    rerandomise_zero_layers(vae, std=0.05)
    g = torch.Generator().manual_seed(17)
    B, T, N, L = 2, cfg["num_timesteps"], cfg["num_inputs"], cfg["num_latents"]
    static_pc = torch.rand(B, N, 3, generator=g) - 0.5
    delta_pc = torch.randn(B, T, N, 3, generator=g) * 0.05
    gs_list = []
    for Pn in (100, 80):
        gsx = torch.randn(Pn, 14, generator=g) * 0.3
        gsx[:, :3] = torch.rand(Pn, 3, generator=g) - 0.5
        gs_list.append(gsx)
    keep = ("cross_attend_blocks.", "input_embedding.", "mean_fc.", "logvar_fc.")
    out = {"cfg": cfg, "static_pc": static_pc, "delta_pc": delta_pc, "static_gs": gs_list,
           "state_dict": {k: v.clone() for k, v in vae.state_dict().items() if k.startswith(keep)}}
    with torch.no_grad():
        for name, ctx in (("fp32", None), ("autocast_fp16", torch.autocast("cpu", dtype=torch.float16))):
            torch.manual_seed(5)
            if ctx is None:
                kl, x, post, sgs = vae.encode(static_pc, delta_pc, gs_list)
            else:
                with ctx:
                    kl, x, post, sgs = vae.encode(static_pc, delta_pc, gs_list)
            out[name] = {"kl": kl.float(), "x": x.float(), "mean": post.mean.float(), "logvar": post.logvar.float(),
                         "sampled_static_gs": sgs}
        torch.manual_seed(5)
        out["noise"] = torch.randn(out["fp32"]["mean"].shape)
    return out


def gen_p_sample(diffusion):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1, 4, 8, 8, 8, generator=g)
    W = torch.randn(4, 4, generator=g) * 0.5
    model = lambda x, ts: torch.einsum("oc,bcdhw->bodhw", W, x) * torch.cos(ts / 1000.0).view(-1, 1, 1, 1, 1)
    outs = {}
    for tt in (999, 500, 1, 0):
        torch.manual_seed(100 + tt)
        o = diffusion.p_sample(model, x, torch.tensor([tt]))
        torch.manual_seed(100 + tt)
        noise = torch.randn_like(x)
        outs[tt] = {"sample": o["sample"], "pred_xstart": o["pred_xstart"], "noise": noise}
    return {"x": x, "W": W, "outs": outs}


def gen_respace():
    """space_timesteps / SpacedDiffusion of the reference (model/respace.py, utils/script_util.py) for the product-side
    twin gvfdiffusion_b200/model/{gaussian_diffusion,respace}.py: kept-timestep sets, respaced betas / timestep maps and
    one p_sample per prediction type on a respaced, rescaled process."""
    from model.respace import space_timesteps
    out = {"space": {}}
    for n, sec in ((1000, "100"), (1000, "ddim50"), (1000, "10,15,20"), (1000, "fast27"), (300, [10, 15, 20]), (1000, [1000]),
                   (1000, "7"), (997, "13,5")):
        out["space"][(n, str(sec))] = sorted(space_timesteps(n, sec))
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 4, 8, 8, 8, generator=g)
    W = torch.randn(4, 4, generator=g) * 0.5
    model = lambda x, ts: torch.einsum("oc,bcdhw->bodhw", W, x) * torch.cos(ts / 1000.0).view(-1, 1, 1, 1, 1)
    out["x"], out["W"], out["cases"] = x, W, []
    for ptype, sched, resp, small in (("v", "cosine", "50", False), ("eps", "linear", "ddim25", True),
                                      ("xstart", "cosine", "", False)):
        cfg = dict(DIFF_CFG, predict_type=ptype, noise_schedule=sched, timestep_respacing=resp, sigma_small=small)
        d = create_gaussian_diffusion(**cfg)
        case = {"cfg": cfg, "betas": torch.from_numpy(d.betas), "timestep_map": list(d.timestep_map), "outs": {}}
        for tt in (d.num_timesteps - 1, d.num_timesteps // 2, 0):
            t = torch.tensor([tt, max(tt - 1, 0)])
            for clip in (True, False):
                torch.manual_seed(7 + tt)
                o = d.p_sample(model, x, t, clip_denoised=clip)
                case["outs"][(tt, clip)] = {"sample": o["sample"], "pred_xstart": o["pred_xstart"]}
        out["cases"].append(case)
    return out


def gen_gaussian():
    """GaussianModel activations (+ delta).  The reference hard-codes .cuda() at construction:
    patch Tensor.cuda to identity for this CPU run."""
    import types
    for name in ("utils3d", "plyfile", "easydict"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["plyfile"].PlyData = sys.modules["plyfile"].PlyElement = object
    sys.modules["easydict"].EasyDict = dict
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        from representations.gaussian.gaussian_model import GaussianModel
        gm = GaussianModel(sh_degree=0, aabb=[-0.5, -0.5, -0.5, 1.0, 1.0, 1.0], mininum_kernel_size=0.0009,
                           scaling_bias=0.004, opacity_bias=0.1, scaling_activation="softplus", device="cpu")
        g = torch.Generator().manual_seed(5)
        P = 257
        gm._xyz = torch.rand(P, 3, generator=g)
        gm._features_dc = torch.randn(P, 1, 3, generator=g)
        gm._scaling = torch.randn(P, 3, generator=g) * 2
        gm._rotation = torch.randn(P, 4, generator=g) * 0.1
        gm._opacity = torch.randn(P, 1, generator=g) * 2
        delta = torch.randn(P, 14, generator=g) * 0.1
        out = {"raw": {k: getattr(gm, k).clone() for k in ("_xyz", "_features_dc", "_scaling", "_rotation", "_opacity")},
               "delta": delta,
               "scale_bias": gm.scale_bias.clone(), "opacity_bias": gm.opacity_bias.clone(),
               "plain": [gm.get_xyz, gm.get_scaling, gm.get_rotation, gm.get_features, gm.get_opacity],
               "with_delta": [gm.get_xyz_with_delta(delta[:, :3]), gm.get_scaling_with_delta(delta[:, 3:6]),
                              gm.get_rotation_with_delta(delta[:, 6:10]),
                              gm.get_features_with_delta(delta[:, 10:13].unsqueeze(1)),
                              gm.get_opacity_with_delta(delta[:, 13:])]}
    finally:
        torch.Tensor.cuda = orig
    return out


def gen_window_partition():
    """The reference's own calc_window_partition (sparse/attention/windowed_attn.py:20-58) on seeded sparse
    voxel sets: two batch entries of a 64^3 grid, window 8, unshifted and shifted by half a window (the two
    settings the swin blocks alternate, sparse_transformer.py).  Stored: the inputs, the window id of every
    voxel in the reference's forward order (sorted ids -- independent of argsort's tie order), seq_lens and
    seq_batch_indices."""
    import types
    from sparse.attention.windowed_attn import calc_window_partition
    g = torch.Generator().manual_seed(11)
    cases = []
    for (n_vox, res, window, shift) in [(1500, 64, 8, (0, 0, 0)), (1500, 64, 8, (4, 4, 4)), (300, 16, 4, (2, 2, 2)),
                                        (64, 8, 8, (0, 0, 0))]:
        coords = []
        for b in range(2):
            lin = torch.randperm(res ** 3, generator=g)[:n_vox]
            xyz = torch.stack([lin // (res * res), (lin // res) % res, lin % res], 1)
            coords.append(torch.cat([torch.full((n_vox, 1), b), xyz], 1))
        coords = torch.cat(coords).int()
        t = types.SimpleNamespace(coords=coords, device=coords.device)
        fwd, bwd, seq_lens, seq_batch = calc_window_partition(t, window, shift)
        assert torch.equal(bwd[fwd], torch.arange(fwd.shape[0]))
        cases.append({"coords": coords, "window": window, "shift": shift, "fwd": fwd, "bwd": bwd,
                      "seq_lens": seq_lens, "seq_batch_indices": seq_batch})
    return cases


def gen_serialization():
    """The reference's own calc_serialization (sparse/attention/serialized_attn.py:38-119) with `vox2seq.encode`
    provided by the reference's PyTorch twin of its extension (vox2seq/vox2seq/pytorch): forward / backward indices,
    sequence lengths and batch indices for the four serialisation modes, shifted sequences and shifted windows."""
    import importlib.util
    import types
    twin_dir = os.path.join(_ref_import.REF, "model", "sparse_voxel_diffusion", "vox2seq", "vox2seq", "pytorch")
    spec = importlib.util.spec_from_file_location("_vox2seq_twin", os.path.join(twin_dir, "__init__.py"),
                                                  submodule_search_locations=[twin_dir])
    twin = importlib.util.module_from_spec(spec)
    sys.modules["_vox2seq_twin"] = twin
    spec.loader.exec_module(twin)
    sys.modules["vox2seq"].encode = twin.encode
    from sparse.attention.serialized_attn import SerializeMode, calc_serialization
    g = torch.Generator().manual_seed(21)
    cases = []
    for (counts, res, window, mode, shift_seq, shift_win) in [
            ((700, 333), 64, 256, "Z_ORDER", 0, (0, 0, 0)), ((700, 333), 64, 256, "HILBERT", 0, (0, 0, 0)),
            ((700, 333), 64, 256, "Z_ORDER_TRANSPOSED", 128, (0, 0, 0)), ((700, 333), 64, 256, "HILBERT_TRANSPOSED", 64, (3, 5, 7)),
            ((100, 1025, 64), 32, 64, "HILBERT", 16, (1, 0, 2)), ((50,), 16, 64, "Z_ORDER", 0, (0, 0, 0))]:
        coords, layout, off = [], [], 0
        for b, n_vox in enumerate(counts):
            lin = torch.randperm(res ** 3, generator=g)[:n_vox]
            xyz = torch.stack([lin // (res * res), (lin // res) % res, lin % res], 1)
            coords.append(torch.cat([torch.full((n_vox, 1), b), xyz], 1))
            layout.append(slice(off, off + n_vox))
            off += n_vox
        coords = torch.cat(coords).int()
        t = types.SimpleNamespace(coords=coords, layout=layout, device=coords.device)
        fwd, bwd, seq_lens, seq_batch = calc_serialization(t, window, SerializeMode[mode], shift_seq, shift_win)
        cases.append({"coords": coords, "counts": counts, "window": window, "mode": mode, "shift_sequence": shift_seq,
                      "shift_window": shift_win, "fwd": fwd, "bwd": bwd, "seq_lens": list(seq_lens),
                      "seq_batch_indices": list(seq_batch)})
    return cases


def gen_flow_euler():
    """The reference's FlowEuler samplers (trellis/pipelines/samplers/flow_euler.py:11-199, TRELLIS stage in front of the
    path) on a toy velocity model, CPU fp32: plain, classifier-free guidance and guidance interval."""
    import importlib.util
    import types
    try:
        import easydict  # noqa: F401
    except ImportError:
        m = types.ModuleType("easydict")

        class EasyDict(dict):
            __getattr__ = dict.get

            def __setattr__(self, k, v):
                self[k] = v
        m.EasyDict = EasyDict
        sys.modules["easydict"] = m
    d = os.path.join(_ref_import.REF, "trellis", "pipelines", "samplers")
    pkg = types.ModuleType("refsamplers")
    pkg.__path__ = [d]
    sys.modules["refsamplers"] = pkg
    spec = importlib.util.spec_from_file_location("refsamplers.flow_euler", os.path.join(d, "flow_euler.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["refsamplers.flow_euler"] = mod
    spec.loader.exec_module(mod)
    g = torch.Generator().manual_seed(31)
    noise = torch.randn(2, 8, 16, generator=g)
    W = torch.randn(16, 16, generator=g) * 0.3
    cond, neg = torch.randn(2, 1, 16, generator=g), torch.zeros(2, 1, 16)
    model = lambda x, t, c: torch.tanh(x @ W) * torch.cos(t / 1000.0).view(-1, 1, 1) + 0.1 * c
    out = {"noise": noise, "W": W, "cond": cond, "neg_cond": neg, "cases": {}}
    r = mod.FlowEulerSampler(1e-5).sample(model, noise, cond, steps=12, rescale_t=1.0, verbose=False)
    out["cases"]["plain"] = {"args": dict(steps=12, rescale_t=1.0), "samples": r.samples, "pred_x_0": r.pred_x_0[-1]}
    r = mod.FlowEulerSampler(1e-5).sample(model, noise, cond, steps=7, rescale_t=3.0, verbose=False)
    out["cases"]["rescaled"] = {"args": dict(steps=7, rescale_t=3.0), "samples": r.samples, "pred_x_0": r.pred_x_0[-1]}
    r = mod.FlowEulerCfgSampler(1e-5).sample(model, noise, cond, neg, steps=9, rescale_t=1.0, cfg_strength=3.0, verbose=False)
    out["cases"]["cfg"] = {"args": dict(steps=9, rescale_t=1.0, cfg_strength=3.0), "samples": r.samples, "pred_x_0": r.pred_x_0[-1]}
    r = mod.FlowEulerGuidanceIntervalSampler(1e-5).sample(model, noise, cond, neg, steps=9, rescale_t=3.0, cfg_strength=7.5,
                                                            cfg_interval=(0.5, 0.95), verbose=False)
    out["cases"]["interval"] = {"args": dict(steps=9, rescale_t=3.0, cfg_strength=7.5, cfg_interval=(0.5, 0.95)),
                                "samples": r.samples, "pred_x_0": r.pred_x_0[-1]}
    return out


def gen_slat_flow():
    """The reference's own SLatFlowModel (trellis/models/structured_latent_flow.py:69-262: SparseResBlock3d down / up
    blocks around ModulatedSparseTransformerCrossBlocks with full sparse self-attention and dense-context cross-attention)
    on the CPU in fp32, seeded two-entry batch, plus two Euler steps of the reference's FlowEulerSampler over it.
    Stand-ins (third-party code that is absent or CUDA-only): spconv.pytorch.SparseConvTensor as a container,
    spconv.pytorch.SubMConv3d restated as the dense cross-correlation at the active sites (oracle.sparse_vae.subm_conv3d:
    the submanifold definition; spconv itself is absent, so that arithmetic stays unpinned), flash_attn's varlen entry
    points restated as plain softmax attention inside each cu_seqlens segment."""
    import types
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import sparse_vae as OSV

    class _SCT:
        def __init__(self, features, indices, spatial_shape, batch_size, grid=None, voxel_num=None, indice_dict=None, **kw):
            self._features, self.indices, self.spatial_shape, self.batch_size = features, indices, spatial_shape, batch_size
            self.grid, self.voxel_num, self.indice_dict = grid, voxel_num, indice_dict
            self.benchmark = self.benchmark_record = self.thrust_allocator = self._timer = None
            self.force_algo = self.int8_scale = None

        @property
        def features(self):
            return self._features

        def replace_feature(self, f):
            return _SCT(f, self.indices, self.spatial_shape, self.batch_size)

    class SubMConv3d(torch.nn.Module):
        def __init__(self, cin, cout, ks, dilation=1, bias=True, indice_key=None, algo=None):
            super().__init__()
            self.in_channels, self.out_channels, self.ks, self.dilation = cin, cout, ks, dilation
            self.weight = torch.nn.Parameter(torch.zeros(cout, ks, ks, ks, cin))
            self.bias = torch.nn.Parameter(torch.zeros(cout)) if bias else None

        def forward(self, x):
            grid = int(x.indices[:, 1:].max()) + 1
            return x.replace_feature(OSV.subm_conv3d(x.features, x.indices, self.weight, self.bias, x.batch_size, grid,
                                                     self.dilation))

    spp = sys.modules["spconv.pytorch"]
    saved_sct = spp.SparseConvTensor
    spp.SparseConvTensor, spp.SubMConv3d = _SCT, SubMConv3d
    sdpa = torch.nn.functional.scaled_dot_product_attention

    def _seg(q, k, v, cq, ck):
        out = torch.empty_like(q)
        for i in range(len(cq) - 1):
            a, b, c, d = int(cq[i]), int(cq[i + 1]), int(ck[i]), int(ck[i + 1])
            out[a:b] = sdpa(q[a:b].transpose(0, 1).float(), k[c:d].transpose(0, 1).float(),
                            v[c:d].transpose(0, 1).float()).transpose(0, 1).to(q.dtype)
        return out

    fa = types.ModuleType("flash_attn")
    fa.flash_attn_varlen_qkvpacked_func = lambda qkv, cu, maxlen: _seg(qkv[:, 0], qkv[:, 1], qkv[:, 2], cu, cu)
    fa.flash_attn_varlen_kvpacked_func = lambda q, kv, cq, ck, mq, mk: _seg(q, kv[:, 0], kv[:, 1], cq, ck)
    fa.flash_attn_varlen_func = lambda q, k, v, cq, ck, mq, mk: _seg(q, k, v, cq, ck)
    saved_fa = sys.modules.get("flash_attn")
    sys.modules["flash_attn"] = fa
    pkg = types.ModuleType("trellis")           # a bare package: trellis/__init__.py pulls in renderers / pipelines
    pkg.__path__ = [os.path.join(_ref_import.REF, "trellis")]
    sys.modules["trellis"] = pkg
    try:
        from trellis.models.structured_latent_flow import SLatFlowModel
        from trellis.modules import sparse as tsp
        cfg = dict(resolution=16, in_channels=8, model_channels=128, cond_channels=128, out_channels=8, num_blocks=2,
                   num_head_channels=64, mlp_ratio=4, patch_size=2, num_io_res_blocks=2, io_block_channels=[64],
                   pe_mode="ape", use_fp16=False, use_skip_connection=True, share_mod=False, qk_rms_norm=True,
                   qk_rms_norm_cross=False)
        torch.manual_seed(0)
        m = SLatFlowModel(**cfg).eval()
        g0 = torch.Generator().manual_seed(77)
        for mod in m.modules():                 # the stand-in convolutions: a seeded draw at Kaiming scale
            if isinstance(mod, SubMConv3d):
                mod.weight.data = torch.randn(mod.weight.shape, generator=g0) / (mod.in_channels * 27) ** 0.5
        rerandomise_zero_layers(m)
        for mod in m.modules():                 # norm affines / rms gammas / biases away from their 1 / 0 initial values
            if isinstance(mod, torch.nn.LayerNorm) and mod.elementwise_affine:
                mod.weight.data += 0.1 * torch.randn(mod.weight.shape, generator=g0)
                mod.bias.data += 0.1 * torch.randn(mod.bias.shape, generator=g0)
        for n_, p_ in m.named_parameters():
            if n_.endswith("gamma") or (n_.endswith(".bias") and "norm" not in n_):
                p_.data += 0.1 * torch.randn(p_.shape, generator=g0)
        for p_ in m.parameters():               # fp16-representable values: the fixture stores them as halves, exactly
            p_.data = p_.data.half().float()
        g = torch.Generator().manual_seed(5)
        coords = []
        for b, n in enumerate((260, 140)):
            lin = torch.randperm(16 ** 3, generator=g)[:n].sort().values
            coords.append(torch.stack([torch.full((n,), b), lin // 256, (lin // 16) % 16, lin % 16], 1))
        coords = torch.cat(coords).int()
        x = torch.randn(coords.shape[0], 8, generator=g)
        cond = torch.randn(2, 24, 128, generator=g)
        t = torch.tensor([650.0, 120.0])
        with torch.no_grad():
            out = m(tsp.SparseTensor(x, coords), t, cond).feats
            # hand-checked quirk: SparseDownsample's scatter_reduce(zeros, 'mean') counts the initial zero
            down = tsp.SparseDownsample(2)(tsp.SparseTensor(x, coords))
        import importlib.util
        try:
            import easydict  # noqa: F401
        except ImportError:
            ed = types.ModuleType("easydict")

            class EasyDict(dict):
                __getattr__ = dict.get

                def __setattr__(self, k, v):
                    self[k] = v
            ed.EasyDict = EasyDict
            sys.modules["easydict"] = ed
        d = os.path.join(_ref_import.REF, "trellis", "pipelines", "samplers")
        spk = types.ModuleType("refsamplers2")
        spk.__path__ = [d]
        sys.modules["refsamplers2"] = spk
        spec = importlib.util.spec_from_file_location("refsamplers2.flow_euler", os.path.join(d, "flow_euler.py"))
        fe = importlib.util.module_from_spec(spec)
        sys.modules["refsamplers2.flow_euler"] = fe
        spec.loader.exec_module(fe)
        with torch.no_grad():
            r = fe.FlowEulerSampler(1e-5).sample(m, tsp.SparseTensor(x, coords), cond=cond, steps=2, rescale_t=3.0,
                                                 verbose=False)
        return {"cfg": cfg, "state_dict": {k: v.half() for k, v in m.state_dict().items()}, "coords": coords, "x": x,
                "cond": cond, "t": t, "out": out, "down_feats": down.feats, "down_coords": down.coords.int(),
                "euler_samples": r.samples.feats, "euler_args": dict(steps=2, rescale_t=3.0)}
    finally:
        spp.SparseConvTensor = saved_sct
        sys.modules.pop("trellis", None)
        for k in [k for k in sys.modules if k.startswith("trellis.")]:
            sys.modules.pop(k)
        if saved_fa is not None:
            sys.modules["flash_attn"] = saved_fa
        else:
            sys.modules.pop("flash_attn", None)


def gen_slat_decoder_gs():
    """The reference's own SLatGaussianDecoder (trellis/models/structured_latent_vae/decoder_gs.py:10-122 over
    SparseTransformerBase, base.py:36-117: input_layer + APE -> swin SparseTransformerBlocks -> layer_norm -> out_layer ->
    to_representation) on the CPU in fp32.  Stand-ins: the SparseConvTensor container, flash_attn's packed-QKV calls as
    per-segment SDPA, and a plain attribute container for trellis.representations.Gaussian (its constructor needs CUDA and
    utils3d; the decoder only sets attributes on it)."""
    import types

    class _SCT:
        def __init__(self, features, indices, spatial_shape, batch_size, grid=None, voxel_num=None, indice_dict=None, **kw):
            self._features, self.indices, self.spatial_shape, self.batch_size = features, indices, spatial_shape, batch_size
            self.grid, self.voxel_num, self.indice_dict = grid, voxel_num, indice_dict
            self.benchmark = self.benchmark_record = self.thrust_allocator = self._timer = None
            self.force_algo = self.int8_scale = None

        @property
        def features(self):
            return self._features

        def replace_feature(self, f):
            return _SCT(f, self.indices, self.spatial_shape, self.batch_size)

    class Gaussian:
        def __init__(self, **kw):
            self.init_params = kw

    spp = sys.modules["spconv.pytorch"]
    saved_sct = spp.SparseConvTensor
    spp.SparseConvTensor = _SCT
    sdpa = torch.nn.functional.scaled_dot_product_attention

    def varlen(qkv, cu, maxlen):
        out = torch.empty_like(qkv[:, 0])
        for i in range(len(cu) - 1):
            a, b = int(cu[i]), int(cu[i + 1])
            q, k, v = (t.transpose(0, 1).float() for t in qkv[a:b].unbind(1))
            out[a:b] = sdpa(q, k, v).transpose(0, 1).to(qkv.dtype)
        return out

    def packed(qkv):
        q, k, v = (t.transpose(1, 2).float() for t in qkv.unbind(2))
        return sdpa(q, k, v).transpose(1, 2).to(qkv.dtype)

    fa = types.ModuleType("flash_attn")
    fa.flash_attn_varlen_qkvpacked_func, fa.flash_attn_qkvpacked_func = varlen, packed
    saved_fa = sys.modules.get("flash_attn")
    sys.modules["flash_attn"] = fa

    def bare(name, path):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
    T = os.path.join(_ref_import.REF, "trellis")
    bare("trellis", T)
    bare("trellis.models", os.path.join(T, "models"))
    bare("trellis.models.structured_latent_vae", os.path.join(T, "models", "structured_latent_vae"))
    bare("trellis.utils", os.path.join(T, "utils"))
    rep = types.ModuleType("trellis.representations")
    rep.Gaussian = Gaussian
    sys.modules["trellis.representations"] = rep
    try:
        from trellis.models.structured_latent_vae.decoder_gs import SLatGaussianDecoder
        from trellis.modules import sparse as tsp
        rep_cfg = {"lr": {"_xyz": 1.0, "_features_dc": 1.0, "_opacity": 1.0, "_scaling": 1.0, "_rotation": 0.1},
                   "perturb_offset": True, "voxel_size": 1.5, "num_gaussians": 4, "2d_filter_kernel_size": 0.1,
                   "3d_filter_kernel_size": 9e-4, "scaling_bias": 4e-3, "opacity_bias": 0.1, "scaling_activation": "softplus"}
        cfg = dict(resolution=16, model_channels=128, latent_channels=8, num_blocks=2, num_head_channels=64, mlp_ratio=4,
                   attn_mode="swin", window_size=8, pe_mode="ape", use_fp16=False, qk_rms_norm=False,
                   representation_config=rep_cfg)
        torch.manual_seed(0)
        m = SLatGaussianDecoder(**cfg).eval()
        rerandomise_zero_layers(m, std=0.05)
        g = torch.Generator().manual_seed(15)
        coords = []
        for b, n in enumerate((170, 110)):
            lin = torch.randperm(16 ** 3, generator=g)[:n].sort().values
            coords.append(torch.stack([torch.full((n,), b), lin // 256, (lin // 16) % 16, lin % 16], 1))
        coords = torch.cat(coords).int()
        latent = torch.randn(coords.shape[0], 8, generator=g)
        rows = {}
        def _keep(mod, i, o):                   # a forward hook that returns a value would replace the output
            rows["feats"] = o.feats.detach().clone()
        hook = m.out_layer.register_forward_hook(_keep)
        with torch.no_grad():
            reps = m(tsp.SparseTensor(latent, coords))
        hook.remove()
        names = ("_xyz", "_features_dc", "_scaling", "_rotation", "_opacity")
        return {"cfg": cfg, "state_dict": {k: v.clone() for k, v in m.state_dict().items()}, "coords": coords, "latent": latent,
                "rows": rows["feats"], "reps": [{n: getattr(r, n).clone() for n in names} for r in reps],
                "init_params": [r.init_params for r in reps]}
    finally:
        spp.SparseConvTensor = saved_sct
        for k in [k for k in sys.modules if k == "trellis" or k.startswith("trellis.")]:
            sys.modules.pop(k)
        if saved_fa is not None:
            sys.modules["flash_attn"] = saved_fa
        else:
            sys.modules.pop("flash_attn", None)


def gen_sparse_structure_flow():
    """The reference's own SparseStructureFlowModel (trellis/models/sparse_structure_flow.py:55-200: the dense DiT over the
    16^3 occupancy latent that precedes the structured-latent stage) on the CPU in fp32 (attention backend sdpa), for patch
    sizes 1 (shipped) and 2."""
    import types
    pkg = types.ModuleType("trellis")
    pkg.__path__ = [os.path.join(_ref_import.REF, "trellis")]
    sys.modules["trellis"] = pkg
    try:
        from trellis.models.sparse_structure_flow import SparseStructureFlowModel
        out = {}
        for name, ps in (("p1", 1), ("p2", 2)):
            cfg = dict(resolution=8, in_channels=8, model_channels=128, cond_channels=128, out_channels=8, num_blocks=2,
                       num_head_channels=64, mlp_ratio=4, patch_size=ps, pe_mode="ape", use_fp16=False, share_mod=False,
                       qk_rms_norm=True, qk_rms_norm_cross=False)
            torch.manual_seed(3 + ps)
            m = SparseStructureFlowModel(**cfg).eval()
            rerandomise_zero_layers(m)
            g0 = torch.Generator().manual_seed(21)
            for n_, p_ in m.named_parameters():
                if n_.endswith("gamma") or n_.endswith(".bias") or ("norm" in n_ and n_.endswith(".weight")):
                    p_.data += 0.1 * torch.randn(p_.shape, generator=g0)
            for p_ in m.parameters():
                p_.data = p_.data.half().float()
            g = torch.Generator().manual_seed(8)
            x = torch.randn(2, 8, 8, 8, 8, generator=g)
            cond = torch.randn(2, 20, 128, generator=g)
            t = torch.tensor([700.0, 55.0])
            with torch.no_grad():
                y = m(x, t, cond)
            out[name] = {"cfg": cfg, "state_dict": {k: v.half() for k, v in m.state_dict().items() if k != "pos_emb"},
                         "pos_emb": m.pos_emb.clone(), "x": x, "cond": cond, "t": t, "out": y}
        return out
    finally:
        for k in [k for k in sys.modules if k == "trellis" or k.startswith("trellis.")]:
            sys.modules.pop(k)


def gen_sparse_structure_decoder():
    """The reference's own SparseStructureDecoder (trellis/models/sparse_structure_vae.py:209-306: dense Conv3d ResNet,
    ChannelLayerNorm32, pixel-shuffle upsampling; occupancy latent -> occupancy logits) on the CPU in fp32."""
    import types
    pkg = types.ModuleType("trellis")
    pkg.__path__ = [os.path.join(_ref_import.REF, "trellis")]
    sys.modules["trellis"] = pkg
    try:
        from trellis.models.sparse_structure_vae import SparseStructureDecoder
        cfg = dict(out_channels=1, latent_channels=8, num_res_blocks=1, channels=[64, 32, 32], num_res_blocks_middle=1,
                   norm_type="layer", use_fp16=False)
        torch.manual_seed(9)
        m = SparseStructureDecoder(**cfg).eval()
        rerandomise_zero_layers(m, std=0.05)
        g0 = torch.Generator().manual_seed(4)
        for n_, p_ in m.named_parameters():
            if "norm" in n_ or n_.endswith(".bias") or n_.startswith("out_layer.0"):
                p_.data += 0.1 * torch.randn(p_.shape, generator=g0)
        for p_ in m.parameters():
            p_.data = p_.data.half().float()
        g = torch.Generator().manual_seed(6)
        z = torch.randn(2, 8, 4, 4, 4, generator=g)
        with torch.no_grad():
            y = m(z)
        return {"cfg": cfg, "state_dict": {k: v.half() for k, v in m.state_dict().items()}, "z": z, "out": y}
    finally:
        for k in [k for k in sys.modules if k == "trellis" or k.startswith("trellis.")]:
            sys.modules.pop(k)


def _bruteforce_knn_points(p1, p2, lengths1=None, lengths2=None, K=1):
    """Stand-in for pytorch3d.ops.knn_points (absent here) with its documented semantics: exact squared
    distances ((dx*dx + dy*dy) + dz*dz in fp32), ascending, ties -> lowest index, padded rows zero."""
    B, P1, _ = p1.shape
    d = p1[:, :, None, :] - p2[:, None, :, :]
    d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
    if lengths2 is not None:
        d2 = d2.masked_fill(torch.arange(p2.shape[1])[None, None, :] >= lengths2[:, None, None], float("inf"))
    order = torch.sort(d2, dim=-1, stable=True)
    dists, idx = order.values[..., :K].clone(), order.indices[..., :K].clone()
    if lengths1 is not None:
        pad = torch.arange(P1)[None, :] >= lengths1[:, None]
        dists[pad] = 0
        idx[pad] = 0
    return dists, idx, None


def gen_losses():
    """Pixel and interpolation losses of the training step, from the reference's own code:
    utils.loss_util.ssim (imported; `lpips` stubbed, unused) with autograd gradients, and
    train_vae.compute_interpolation_loss_delta_interp -- the function's source is read from
    /root/reference/train_vae.py and executed as is (the module itself needs accelerate / imageio /
    pytorch3d); only pytorch3d.ops.knn_points is replaced by the brute-force stand-in above."""
    import ast
    import types
    import torch.nn.functional as F
    _ref_import._stub("lpips", LPIPS=None)
    from utils.loss_util import ssim
    g = torch.Generator().manual_seed(5)
    out = {"ssim": [], "interp": []}
    for shape in [(2, 3, 40, 52), (1, 3, 70, 33), (3, 1, 16, 16)]:
        gt = torch.rand(shape, generator=g)
        pred = (gt + 0.15 * torch.randn(shape, generator=g)).clamp(0, 1).requires_grad_(True)
        val = ssim(pred, gt)
        l1 = torch.abs(pred - gt).mean()
        (gs,) = torch.autograd.grad(val, pred, retain_graph=True)
        (gl,) = torch.autograd.grad(l1, pred)
        per = ssim(pred, gt, size_average=False)
        out["ssim"].append({"pred": pred.detach(), "gt": gt, "ssim": val.detach(), "l1": l1.detach(),
                            "grad_ssim": gs, "grad_l1": gl, "ssim_per_batch": per.detach()})
    src = open(os.path.join(_ref_import.REF, "train_vae.py")).read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef)
          and n.name == "compute_interpolation_loss_delta_interp"][0]
    ns = {"th": torch, "F": F, "pytorch3d": types.SimpleNamespace(ops=types.SimpleNamespace(knn_points=_bruteforce_knn_points))}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "train_vae.py", "exec"), ns)
    ref_fn = ns["compute_interpolation_loss_delta_interp"]
    for (B, T, sizes, n_pts, k, adaptive) in [(2, 3, (150, 97), 64, 4, True), (1, 4, (200,), 33, 8, True),
                                              (2, 2, (50, 50), 40, 4, False)]:
        static_gs = [torch.cat([torch.rand(n, 3, generator=g) - 0.5, torch.rand(n, 11, generator=g)], 1) for n in sizes]
        micro_static = torch.rand(B, n_pts, 3, generator=g) - 0.5
        micro_moving = micro_static[:, None] + 0.05 * torch.randn(B, T, n_pts, 3, generator=g)
        output = 0.05 * torch.randn(B, T, max(sizes), 14, generator=g)
        output.requires_grad_(True)
        loss, _, est = ref_fn(static_gs, micro_static, micro_moving, output, B, knn_k=k, adaptive_radius=adaptive)
        (go,) = torch.autograd.grad(loss, output)
        padded = torch.stack([F.pad(s[:, :3], (0, 0, 0, max(sizes) - s.shape[0])) for s in static_gs])
        kd, ki, _ = _bruteforce_knn_points(padded, micro_static, lengths1=torch.tensor(sizes), K=k)
        out["interp"].append({"static_gs": static_gs, "micro_static": micro_static, "micro_moving": micro_moving,
                              "output": output.detach(), "knn_k": k, "adaptive": adaptive, "loss": loss.detach(),
                              "estimated": est, "grad_output": go, "knn_dists": kd, "knn_idx": ki})
    return out


def gen_sparse_vae():
    """The reference's own SparseTransformerVAE (model/sparse_voxel_diffusion/sparse_transformer_vae.py) -- swin
    SparseTransformerBlocks over sparse.SparseTensor -- run on the CPU in fp32 on a seeded two-entry batch.
    Stand-ins needed for that: a container for spconv.pytorch.SparseConvTensor (the reference only stores tensors in
    it) and flash_attn's two packed-QKV entry points restated as plain scaled-dot-product attention inside each
    cu_seqlens segment (flash-attn's documented semantics; the CUDA wheel cannot run on the CPU)."""
    import types

    class _SCT:
        def __init__(self, features, indices, spatial_shape, batch_size, grid=None, voxel_num=None, indice_dict=None, **kw):
            self._features, self.indices, self.spatial_shape, self.batch_size = features, indices, spatial_shape, batch_size
            self.grid, self.voxel_num, self.indice_dict = grid, voxel_num, indice_dict
            self.benchmark = self.benchmark_record = self.thrust_allocator = self._timer = None
            self.force_algo = self.int8_scale = None

        @property
        def features(self):
            return self._features

        def replace_feature(self, f):
            return _SCT(f, self.indices, self.spatial_shape, self.batch_size)

    sys.modules["spconv.pytorch"].SparseConvTensor = _SCT
    sdpa = torch.nn.functional.scaled_dot_product_attention

    def varlen(qkv, cu, maxlen):
        out = torch.empty_like(qkv[:, 0])
        for i in range(len(cu) - 1):
            s, e = int(cu[i]), int(cu[i + 1])
            q, k, v = (t.transpose(0, 1).float() for t in qkv[s:e].unbind(1))
            out[s:e] = sdpa(q, k, v).transpose(0, 1).to(qkv.dtype)
        return out

    def packed(qkv):
        q, k, v = (t.transpose(1, 2).float() for t in qkv.unbind(2))
        return sdpa(q, k, v).transpose(1, 2).to(qkv.dtype)

    fa = types.ModuleType("flash_attn")
    fa.flash_attn_varlen_qkvpacked_func, fa.flash_attn_qkvpacked_func = varlen, packed
    saved = sys.modules.get("flash_attn")
    sys.modules["flash_attn"] = fa
    try:
        from model.sparse_voxel_diffusion.sparse_transformer_vae import SparseTransformerVAE
        import sparse as sp
        # use_old_attn_impl: false is what configs/vae.yml:30 and configs/diffusion.yml:57 ship ([3][H][d] qkv channels);
        # the class default True ([H][3][d], sparse/attention/modules.py:161-164) is run on the same weights as well
        cfg = dict(resolution=16, in_channels=16, model_channels=128, out_channels=24, latent_channels=8, num_blocks=2,
                   window_size=8, num_head_channels=64, attn_mode="swin", pe_mode="ape", use_fp16=False, norm_output=True,
                   use_old_attn_impl=False)
        torch.manual_seed(0)
        m = SparseTransformerVAE(**cfg).eval()
        rerandomise_zero_layers(m)
        m_old = SparseTransformerVAE(**dict(cfg, use_old_attn_impl=True)).eval()
        m_old.load_state_dict(m.state_dict())
        g = torch.Generator().manual_seed(1)
        coords = []
        for b, n in enumerate((150, 90)):
            lin = torch.randperm(16 ** 3, generator=g)[:n].sort().values
            coords.append(torch.stack([torch.full((n,), b), lin // 256, (lin // 16) % 16, lin % 16], 1))
        coords = torch.cat(coords).int()
        latent = torch.randn(coords.shape[0], 8, generator=g)
        feats = torch.randn(coords.shape[0], 16, generator=g)
        with torch.no_grad():
            dec = m.decode(sp.SparseTensor(latent, coords)).feats
            _, mean, logvar = m.encode(sp.SparseTensor(feats, coords), sample_posterior=False, return_raw=True)
            dec_old = m_old.decode(sp.SparseTensor(latent, coords)).feats
        return {"cfg": cfg, "state_dict": {k: v.clone() for k, v in m.state_dict().items()}, "coords": coords,
                "latent": latent, "feats": feats, "decode": dec, "mean": mean, "logvar": logvar,
                "decode_old_attn_impl": dec_old}
    finally:
        if saved is not None:
            sys.modules["flash_attn"] = saved
        else:
            sys.modules.pop("flash_attn", None)


def gen_render_call():
    """What the reference hands to the third-party rasteriser: the reference's own GaussianRenderer.render
    (renderers/gaussian_render.py:269-369 -> render() :85-238) is run on the CPU with a RECORDING stand-in for
    diff_gaussian_rasterization (the module is absent here and on the GPU box); the fixture holds the
    GaussianRasterizationSettings fields and the tensors of the GaussianRasterizer call for a seeded GaussianModel,
    delta and orbit camera.  `device="cuda"` literals of the two functions are dropped through a proxy of the torch
    module inside that one reference module."""
    import types
    for name in ("utils3d", "plyfile", "easydict"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["plyfile"].PlyData = sys.modules["plyfile"].PlyElement = object

    class _EDict(dict):
        def __init__(self, d=None, **kw):
            super().__init__()
            for k, v in dict(d or {}, **kw).items():
                self[k] = v
        __getattr__ = dict.__getitem__
        __setattr__ = dict.__setitem__

    sys.modules["easydict"].EasyDict = _EDict
    rec = {}

    class Settings:
        def __init__(self, **kw):
            self.__dict__.update(kw)

    class Rasterizer:
        def __init__(self, raster_settings):
            rec["settings"] = raster_settings

        def __call__(self, **kw):
            rec["call"] = kw
            s = rec["settings"]
            return torch.zeros(3, s.image_height, s.image_width), torch.ones(kw["means3D"].shape[0], dtype=torch.int32)

    sys.modules["diff_gaussian_rasterization"] = types.SimpleNamespace(GaussianRasterizationSettings=Settings,
                                                                       GaussianRasterizer=Rasterizer)

    class _TorchProxy:
        def __getattr__(self, name):
            f = getattr(torch, name)
            if name in ("zeros", "tensor", "zeros_like", "ones"):
                def g(*a, **k):
                    k.pop("device", None)
                    return f(*a, **k)
                return g
            return f

    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        from representations.gaussian.gaussian_model import GaussianModel
        import renderers.gaussian_render as GR
        GR.torch = _TorchProxy()
        gm = GaussianModel(sh_degree=0, aabb=[-0.5, -0.5, -0.5, 1.0, 1.0, 1.0], mininum_kernel_size=0.0009,
                           scaling_bias=0.004, opacity_bias=0.1, scaling_activation="softplus", device="cpu")
        g = torch.Generator().manual_seed(21)
        P = 96
        gm._xyz = torch.rand(P, 3, generator=g)
        gm._features_dc = torch.randn(P, 1, 3, generator=g)
        gm._scaling = torch.randn(P, 3, generator=g)
        gm._rotation = torch.randn(P, 4, generator=g) * 0.1
        gm._opacity = torch.randn(P, 1, generator=g)
        delta = torch.randn(P, 14, generator=g) * 0.05
        sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
        from gvfdiffusion_b200 import synthetic as S
        ext = S.orbit_extrinsics(24)[5]
        intr = S.intrinsics()
        r = GR.GaussianRenderer({"resolution": 64, "near": 0.8, "far": 1.6, "ssaa": 1, "bg_color": (1.0, 1.0, 1.0)})
        r.pipe.use_mip_gaussian = True
        r.pipe.kernel_size = 0.1
        out = {"raw": {k: getattr(gm, k).clone() for k in ("_xyz", "_features_dc", "_scaling", "_rotation", "_opacity")},
               "delta": delta, "extrinsics": ext, "intrinsics": intr, "near": 0.8, "far": 1.6, "resolution": 64}
        for tag, d in (("with_delta", delta), ("no_delta", None)):
            rec.clear()
            ret = r.render(gm, ext, intr, delta_pc=d)
            assert set(ret.keys()) == {"rgb"}
            s_, c_ = rec["settings"], rec["call"]
            out[tag] = {"settings": {k: (v.clone() if torch.is_tensor(v) else v) for k, v in s_.__dict__.items()},
                        "call": {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in c_.items()}}
            out[tag]["settings"]["tanfovx"] = float(s_.tanfovx)
            out[tag]["settings"]["tanfovy"] = float(s_.tanfovy)
    finally:
        torch.Tensor.cuda = orig
        sys.modules.pop("diff_gaussian_rasterization", None)
    return out


def gen_to_representation():
    """The reference's own SparseVAE.to_representation / _build_perturbation / _calc_layout
    (model/sparse_voxel_diffusion/sparse_vae.py:104-180,202-227) with the MipGS block of configs/vae.yml, on a
    seeded two-entry batch of sparse voxels.  The object is created without __init__ (which wants backbones and
    renderers); GaussianModel is the reference's, constructed on the CPU."""
    import functools
    import types
    for name in ("utils3d", "utils3d.torch", "plyfile", "easydict"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["utils3d"].torch = sys.modules["utils3d.torch"]
    sys.modules["plyfile"].PlyData = sys.modules["plyfile"].PlyElement = object

    class _EDict(dict):
        def __init__(self, d=None):
            super().__init__()
            for k, v in (d or {}).items():
                self[k] = _EDict(v) if isinstance(v, dict) else v
        __getattr__ = dict.__getitem__
        __setattr__ = dict.__setitem__

    sys.modules["easydict"].EasyDict = _EDict
    _ref_import._stub("lpips", LPIPS=None)
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        from model.sparse_voxel_diffusion import sparse_vae as SV
        from representations.gaussian.gaussian_model import GaussianModel
        SV.GaussianModel = functools.partial(GaussianModel, device="cpu")
        cfg = dict(SV._DEFAULT_MIPGS_CFG)
        cfg.update({"lr": {"_xyz": 1.0, "_features_dc": 1.0, "_opacity": 1.0, "_scaling": 1.0, "_rotation": 0.1},
                    "perturb_offset": True, "reg_mode": "soft_invoxel", "voxel_size": 1.5, "num_gaussians": 8,
                    "2d_filter_kernel_size": 0.1, "3d_filter_kernel_size": 0.0009, "scaling_bias": 0.004,
                    "opacity_bias": 0.1, "scaling_activation": "softplus"})
        sv = object.__new__(SV.SparseVAE)
        sv.resolution = 64
        sv.rep_config = {"MipGS": cfg}
        sv._calc_layout(sv.rep_config)
        bb = types.SimpleNamespace(device="cpu")
        sv.backbones = {"vae": bb}
        bb.MipGS_perturbation = sv._build_perturbation(8, "soft_invoxel")
        g = torch.Generator().manual_seed(11)
        counts = [37, 20]
        coords = torch.cat([torch.cat([torch.full((n, 1), b), torch.randint(0, 64, (n, 3), generator=g)], dim=1)
                            for b, n in enumerate(counts)]).int()
        feats = torch.randn(sum(counts), 112, generator=g) * 1.5
        x = types.SimpleNamespace(shape=torch.Size([2, 112]), coords=coords, feats=feats,
                                  layout=[slice(0, 37), slice(37, 57)])
        reps = sv.to_representation(x)["MipGS"]
        names = ("_xyz", "_features_dc", "_scaling", "_rotation", "_opacity")
        out = {"cfg": {k: (dict(v) if isinstance(v, dict) else v) for k, v in cfg.items()}, "resolution": 64,
               "coords": coords, "feats": feats, "counts": counts, "perturbation": bb.MipGS_perturbation.clone(),
               "layout": {k: tuple(v["range"]) for k, v in sv.layouts["MipGS"].items()},
               "reps": [{n: getattr(r, n).clone() for n in names} for r in reps],
               "activated": [{"xyz": r.get_xyz.clone(), "scaling": r.get_scaling.clone(),
                              "opacity": r.get_opacity.clone(), "rotation": r.get_rotation.clone()} for r in reps]}
    finally:
        torch.Tensor.cuda = orig
    return out


def gen_lpips():
    """The reference's own LPIPS(net_type='vgg') class (utils/lpips/lpips.py) with its two downloads replaced by seeded
    random weights: torchvision's vgg16 un-pretrained, filled from gvfdiffusion_b200.utils.lpips.LPIPS(seed=3) -- whose
    initialisation is a pure function of the seed -- and the linear heads from the same module.  Only the seed, the input
    seeds and the resulting loss / input-gradient checksums are stored (the 59 MB of weights are regenerated by the test)."""
    import torchvision
    import importlib.util
    spec = importlib.util.spec_from_file_location(          # the one product file this generator reads: the seeded weight
        "_gvf_lpips", os.path.join(HERE, "..", "..", "gvfdiffusion_b200", "utils", "lpips", "lpips.py"))    # initialiser
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ours = mod.LPIPS(seed=3).eval()
    import utils.lpips.networks as RN
    import utils.lpips.lpips as RL
    _vgg16 = torchvision.models.vgg16
    RN.models.vgg16 = lambda *a, **k: _vgg16(weights=None)            # no ImageNet download
    lin_sd = {f"{i}.1.weight": l[1].weight.detach().clone() for i, l in enumerate(ours.lin)}
    RL.get_state_dict = lambda net_type="vgg", version="0.1": lin_sd
    ref = RL.LPIPS(net_type="vgg").eval()
    ref.net.layers.load_state_dict(ours.layers.state_dict())
    g = torch.Generator().manual_seed(11)
    x = (torch.rand(2, 3, 96, 96, generator=g) * 2 - 1).requires_grad_(True)
    y = torch.rand(2, 3, 96, 96, generator=g) * 2 - 1
    loss = ref(x, y)
    loss.backward()
    return {"seed": 3, "input_seed": 11, "shape": (2, 3, 96, 96), "loss": float(loss), "grad_abs_sum": float(x.grad.abs().sum()),
            "grad_probe": x.grad[0, :, ::16, ::16].clone()}


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "lpips":
        torch.save(gen_lpips(), os.path.join(HERE, "lpips.pt"))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "render_call":
        torch.save(gen_render_call(), os.path.join(HERE, "render_call.pt"))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "sparse_vae":
        torch.save(gen_sparse_vae(), os.path.join(HERE, "sparse_vae_tiny.pt"))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "to_representation":
        torch.save(gen_to_representation(), os.path.join(HERE, "to_representation.pt"))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "losses":
        torch.save(gen_losses(), os.path.join(HERE, "losses.pt"))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "vae_encode":
        torch.save(gen_vae_encode(), os.path.join(HERE, "vae_encode_tiny.pt"))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "slat_decoder_gs":
        torch.save(gen_slat_decoder_gs(), os.path.join(HERE, "slat_decoder_gs_tiny.pt"))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "sparse_structure_flow":
        torch.save(gen_sparse_structure_flow(), os.path.join(HERE, "sparse_structure_flow_tiny.pt"))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "sparse_structure_decoder":
        torch.save(gen_sparse_structure_decoder(), os.path.join(HERE, "sparse_structure_decoder_tiny.pt"))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "slat_flow":
        torch.save(gen_slat_flow(), os.path.join(HERE, "slat_flow_tiny.pt"))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "flow_euler":
        torch.save(gen_flow_euler(), os.path.join(HERE, "flow_euler.pt"))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "serialization":
        torch.save(gen_serialization(), os.path.join(HERE, "serialization.pt"))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "respace":
        torch.save(gen_respace(), os.path.join(HERE, "respace.pt"))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "window":
        torch.save(gen_window_partition(), os.path.join(HERE, "window_partition.pt"))
        return
    sched, diffusion, ns = gen_schedule()
    torch.save(sched, os.path.join(HERE, "schedule.pt"))
    torch.save(gen_dit(ns), os.path.join(HERE, "dit_tiny.pt"))
    torch.save(gen_vae(), os.path.join(HERE, "vae_tiny.pt"))
    torch.save(gen_vae_encode(), os.path.join(HERE, "vae_encode_tiny.pt"))
    torch.save(gen_p_sample(diffusion), os.path.join(HERE, "p_sample.pt"))
    torch.save(gen_gaussian(), os.path.join(HERE, "gaussian.pt"))
    torch.save(gen_respace(), os.path.join(HERE, "respace.pt"))
    torch.save(gen_serialization(), os.path.join(HERE, "serialization.pt"))
    torch.save(gen_flow_euler(), os.path.join(HERE, "flow_euler.pt"))
    torch.save(gen_slat_flow(), os.path.join(HERE, "slat_flow_tiny.pt"))
    torch.save(gen_slat_decoder_gs(), os.path.join(HERE, "slat_decoder_gs_tiny.pt"))
    torch.save(gen_sparse_structure_flow(), os.path.join(HERE, "sparse_structure_flow_tiny.pt"))
    torch.save(gen_sparse_structure_decoder(), os.path.join(HERE, "sparse_structure_decoder_tiny.pt"))
    torch.save(gen_window_partition(), os.path.join(HERE, "window_partition.pt"))
    torch.save(gen_losses(), os.path.join(HERE, "losses.pt"))
    torch.save(gen_to_representation(), os.path.join(HERE, "to_representation.pt"))
    torch.save(gen_sparse_vae(), os.path.join(HERE, "sparse_vae_tiny.pt"))
    torch.save(gen_render_call(), os.path.join(HERE, "render_call.pt"))
    torch.save(gen_lpips(), os.path.join(HERE, "lpips.pt"))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".pt"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
