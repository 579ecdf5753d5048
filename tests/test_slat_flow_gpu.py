"""TRELLIS structured-latent flow model on the device (reference
trellis/models/structured_latent_flow.py:69-262) against oracle/slat_flow.py -- the fp32 CPU restatement pinned to the
reference's own class by tests/golden/slat_flow_tiny.pt -- plus the two resampling kernels and the sampler over sparse
samples.  Tolerances are relative L2 of fp16 execution against the fp32 oracle; the measured errors are printed."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
G = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "slat_flow_tiny.pt"), weights_only=False)


def _rel(a, b):
    return float((a.float().cpu() - b).norm() / b.norm())


def _model(cfg, sd):
    from gvfdiffusion_b200.trellis.models import SLatFlowModel
    return SLatFlowModel(**cfg, device=DEV).load_state_dict(sd)


def test_pool_mean_and_gather_concat_kernels():
    from gvfdiffusion_b200 import ops
    from gvfdiffusion_b200.sparse.basic import SparseTensor
    from gvfdiffusion_b200.sparse.spatial import SparseDownsample, SparseUpsample
    from oracle import slat_flow as O
    g = torch.Generator().manual_seed(3)
    coords = G["coords"]
    x = torch.randn(coords.shape[0], 64, generator=g).half()
    f_ref, c_ref, idx_ref = O.downsample(x.float(), coords)
    st = SparseTensor(x.to(DEV), coords.to(DEV))
    down = SparseDownsample(2)(st)
    assert torch.equal(down.coords.cpu(), c_ref)
    assert float((down.feats.float().cpu() - f_ref).abs().max()) < 2e-3           # one fp16 rounding of an fp32 sum
    up = SparseUpsample(2)(down)
    assert torch.equal(up.coords.cpu(), coords) and torch.equal(up.feats.cpu(), down.feats.cpu()[idx_ref])
    b = torch.randn(coords.shape[0], 24, generator=g).half().to(DEV)
    cat = ops.gather_concat(a=down.feats, idx=down.get_spatial_cache("upsample_(2, 2, 2)_idx"), b=b)
    assert torch.equal(cat.cpu(), torch.cat([down.feats.cpu()[idx_ref], b.cpu()], 1))
    assert torch.equal(ops.gather_concat(a=x.to(DEV), b=b).cpu(), torch.cat([x, b.cpu()], 1))


def test_ln_silu_kernel():
    from gvfdiffusion_b200 import ops
    g = torch.Generator().manual_seed(4)
    for C in (32, 64, 128, 256, 512, 1024, 2048):
        x = torch.randn(77, C, generator=g).half()
        w, b = torch.randn(C, generator=g), torch.randn(C, generator=g)
        ref = torch.nn.functional.silu(torch.nn.functional.layer_norm(x.float(), (C,), w, b, 1e-6))
        got = ops.ln_mod_act(x.to(DEV), w=w.to(DEV), b=b.to(DEV), act=1)
        assert _rel(got, ref) < 1e-3, C
        sc, sh = (torch.randn(1, C, generator=g) * 0.3).half(), (torch.randn(1, C, generator=g) * 0.3).half()
        mod = torch.cat([sc, sh], 1).to(DEV)
        ref = torch.nn.functional.silu(torch.nn.functional.layer_norm(x.float(), (C,), None, None, 1e-6) * (1 + sc.float()) + sh.float())
        got = ops.ln_mod_act(x.to(DEV), scale=mod[0, :C], shift=mod[0, C:], mod_stride=2 * C, act=1)
        assert _rel(got, ref) < 1e-3, C


def test_conv_residual_epilogue_matches_two_kernel_path():
    from gvfdiffusion_b200 import ops
    from gvfdiffusion_b200.sparse.basic import SparseTensor
    from gvfdiffusion_b200.sparse.conv import SparseConv3d
    g = torch.Generator().manual_seed(6)
    coords = G["coords"].to(DEV)
    x = torch.randn(coords.shape[0], 64, generator=g).half().to(DEV)
    conv = SparseConv3d(64, 128, 3, device=DEV)
    conv.weight = (torch.randn(128, 27 * 64, generator=g) / 40).half().to(DEV)
    conv.bias = torch.randn(128, generator=g).to(DEV)
    st = SparseTensor(x, coords)
    base = torch.randn(coords.shape[0], 128, generator=g).half().to(DEV)
    y = conv(st).feats
    out = base.clone()
    ops.sparse_conv_gemm(x, conv.neighbor_map(st), conv.weight, conv.bias, out=out, residual=True)
    assert torch.equal(out, (y.float() + base.float()).half())


def test_upsampled_conv_as_tap_products_plus_gather_sum():
    """conv(upsample(a)) == gather-sum over the taps of the coarse rows' per-tap products (fp32 until the single rounding)."""
    from gvfdiffusion_b200 import ops
    from gvfdiffusion_b200.sparse.basic import SparseTensor
    from gvfdiffusion_b200.sparse.conv import SparseConv3d
    from gvfdiffusion_b200.sparse.spatial import downsample_plan
    from oracle import sparse_vae as OSV
    g = torch.Generator().manual_seed(8)
    coords = G["coords"]
    st = SparseTensor(torch.zeros(coords.shape[0], 8, device=DEV), coords.to(DEV))
    plan = downsample_plan(st, 2)
    cells, Cin, Cout = plan["coords"].shape[0], 128, 64
    a = torch.randn(cells, Cin, generator=g).half()
    w = (torch.randn(Cout, 3, 3, 3, Cin, generator=g) / (27 * Cin) ** 0.5).half()
    bias = torch.randn(Cout, generator=g)
    idx = plan["idx"].cpu().long()
    ref = OSV.subm_conv3d(a.float()[idx], coords, w.float(), bias, 2, 16)
    tap_w = w.reshape(Cout, 27, Cin).permute(1, 0, 2).reshape(27 * Cout, Cin).contiguous().to(DEV)
    P = ops.gemm(a.to(DEV), tap_w, None, ops.EPI_F32)
    conv = SparseConv3d(Cin, Cout, 3, device=DEV)
    got = ops.sparse_tap_gather_sum(P, conv.neighbor_map(st), plan["idx"], bias.to(DEV))
    assert _rel(got, ref) < 1e-3
    # idx = None: the same formulation for a tensor that is not upsampled
    x = torch.randn(coords.shape[0], Cin, generator=g).half()
    P = ops.gemm(x.to(DEV), tap_w, None, ops.EPI_F32)
    got = ops.sparse_tap_gather_sum(P, conv.neighbor_map(st), None, bias.to(DEV))
    assert _rel(got, OSV.subm_conv3d(x.float(), coords, w.float(), bias, 2, 16)) < 1e-3


def test_conv_modes_are_bit_identical():
    from gvfdiffusion_b200.sparse.basic import SparseTensor
    outs = []
    for mode in ("fused", "im2col"):
        m = _model(G["cfg"], G["state_dict"])
        for blk in m.input_blocks + m.out_blocks:
            blk.conv_mode = mode
        outs.append(m(SparseTensor(G["x"].to(DEV), G["coords"].to(DEV)), G["t"].to(DEV), G["cond"].to(DEV)).feats)
    assert torch.equal(outs[0], outs[1])


def test_forward_tiny_golden():
    from gvfdiffusion_b200.sparse.basic import SparseTensor
    m = _model(G["cfg"], G["state_dict"])
    out = m(SparseTensor(G["x"].to(DEV), G["coords"].to(DEV)), G["t"].to(DEV), G["cond"].to(DEV))
    err = _rel(out.feats, G["out"])
    print(f"\nslat flow tiny vs the reference's own class (fp32): rel L2 {err:.2e}")
    assert out.feats.shape == G["out"].shape and err < 5e-3, err
    # a second call (conditioning K / V and neighbour maps from their caches) gives the same bits
    out2 = m(SparseTensor(G["x"].to(DEV), G["coords"].to(DEV)), G["t"].to(DEV), G["cond"].to(DEV))
    assert torch.equal(out.feats, out2.feats)


def test_graph_replay_is_bit_identical_to_eager():
    from gvfdiffusion_b200.sparse.basic import SparseTensor
    m = _model(G["cfg"], G["state_dict"])
    coords, cond = G["coords"].to(DEV), G["cond"].to(DEV)
    g = torch.Generator().manual_seed(9)
    m.use_graphs = True
    for i in range(3):                                                   # capture, then two replays with new inputs
        x = torch.randn(coords.shape[0], 8, generator=g).to(DEV)
        t = torch.tensor([100.0 + 300 * i, 900.0 - 200 * i], device=DEV)
        got = m(SparseTensor(x, coords), t, cond).feats
        want = m.forward(SparseTensor(x, coords), t, cond).feats
        assert torch.equal(got, want), i
    # another conditioning tensor gets its own K / V and graph
    cond2 = (cond * 0.5).contiguous()
    x = torch.randn(coords.shape[0], 8, generator=g).to(DEV)
    t = torch.tensor([500.0, 500.0], device=DEV)
    assert torch.equal(m(SparseTensor(x, coords), t, cond2).feats, m.forward(SparseTensor(x, coords), t, cond2).feats)
    assert torch.equal(m(SparseTensor(x, coords), t, cond).feats, m.forward(SparseTensor(x, coords), t, cond).feats)


def test_sampler_over_sparse_samples_matches_reference():
    from gvfdiffusion_b200.sparse.basic import SparseTensor
    from gvfdiffusion_b200.trellis.pipelines.samplers import FlowEulerSampler
    m = _model(G["cfg"], G["state_dict"])
    noise = SparseTensor(G["x"].to(DEV), G["coords"].to(DEV))
    r = FlowEulerSampler(1e-5).sample(m, noise, cond=G["cond"].to(DEV), verbose=False, **G["euler_args"])
    err = _rel(r.samples.feats, G["euler_samples"])
    print(f"\nslat flow, 2 Euler steps vs the reference's sampler + model: rel L2 {err:.2e}")
    assert err < 5e-3, err


def test_forward_shipped_widths():
    """The widths of the shipped checkpoint (model 1024, 16 heads x 64, io 128, cond 1024, q / k RMS-norm) at depth 2 on
    ~3000 voxels of a 32^3 grid, one entry, against the oracle."""
    from gvfdiffusion_b200.sparse.basic import SparseTensor
    from oracle import slat_flow as O
    cfg = dict(resolution=32, in_channels=8, model_channels=1024, cond_channels=1024, out_channels=8, num_blocks=2,
               num_heads=16, mlp_ratio=4, patch_size=2, num_io_res_blocks=2, io_block_channels=[128], pe_mode="ape",
               use_fp16=True, qk_rms_norm=True)
    g = torch.Generator().manual_seed(11)
    ref_sd = {k: v for k, v in G["state_dict"].items()}
    shapes = {}
    C, io = 1024, 128

    def res(p, cin, cout):
        shapes.update({p + "norm1.weight": (cin,), p + "norm1.bias": (cin,), p + "conv1.conv.weight": (cout, 3, 3, 3, cin),
                       p + "conv1.conv.bias": (cout,), p + "conv2.conv.weight": (cout, 3, 3, 3, cout), p + "conv2.conv.bias": (cout,),
                       p + "emb_layers.1.weight": (2 * cout, C), p + "emb_layers.1.bias": (2 * cout,)})
        if cin != cout:
            shapes.update({p + "skip_connection.weight": (cout, cin), p + "skip_connection.bias": (cout,)})
    res("input_blocks.0.", io, io)
    res("input_blocks.1.", io, C)
    res("out_blocks.0.", 2 * C, io)
    res("out_blocks.1.", 2 * io, io)
    shapes.update({"t_embedder.mlp.0.weight": (C, 256), "t_embedder.mlp.0.bias": (C,), "t_embedder.mlp.2.weight": (C, C),
                   "t_embedder.mlp.2.bias": (C,), "input_layer.weight": (io, 8), "input_layer.bias": (io,),
                   "out_layer.weight": (8, io), "out_layer.bias": (8,)})
    for i in range(2):
        p = f"blocks.{i}."
        shapes.update({p + "norm2.weight": (C,), p + "norm2.bias": (C,), p + "self_attn.to_qkv.weight": (3 * C, C),
                       p + "self_attn.to_qkv.bias": (3 * C,), p + "self_attn.q_rms_norm.gamma": (16, 64),
                       p + "self_attn.k_rms_norm.gamma": (16, 64), p + "self_attn.to_out.weight": (C, C),
                       p + "self_attn.to_out.bias": (C,), p + "cross_attn.to_q.weight": (C, C), p + "cross_attn.to_q.bias": (C,),
                       p + "cross_attn.to_kv.weight": (2 * C, C), p + "cross_attn.to_kv.bias": (2 * C,),
                       p + "cross_attn.to_out.weight": (C, C), p + "cross_attn.to_out.bias": (C,),
                       p + "mlp.mlp.0.weight": (4 * C, C), p + "mlp.mlp.0.bias": (4 * C,), p + "mlp.mlp.2.weight": (C, 4 * C),
                       p + "mlp.mlp.2.bias": (C,), p + "adaLN_modulation.1.weight": (6 * C, C), p + "adaLN_modulation.1.bias": (6 * C,)})
    assert {k.split(".", 2)[-1] for k in ref_sd if k.startswith("blocks.0.")} == {k.split(".", 2)[-1] for k in shapes if k.startswith("blocks.0.")}
    sd = {}
    for k, shp in shapes.items():
        if k.endswith("norm1.weight") or k.endswith("norm2.weight") or k.endswith("gamma"):
            v = 1 + 0.1 * torch.randn(shp, generator=g)
        elif k.endswith("bias"):
            v = 0.05 * torch.randn(shp, generator=g)
        elif "adaLN" in k or "emb_layers" in k:
            v = 0.02 * torch.randn(shp, generator=g)
        else:
            fan_in = 1
            for s in shp[1:]:
                fan_in *= s
            v = torch.randn(shp, generator=g) / fan_in ** 0.5
        sd[k] = v.half().float()
    n = 3000
    lin = torch.randperm(32 ** 3, generator=g)[:n].sort().values
    coords = torch.stack([torch.zeros(n, dtype=torch.long), lin // 1024, (lin // 32) % 32, lin % 32], 1).int()
    x = torch.randn(n, 8, generator=g)
    cond = torch.randn(1, 1374, 1024, generator=g)
    t = torch.tensor([437.0])
    ref = O.slat_flow_forward(sd, cfg, x, coords, t, cond)
    m = _model(cfg, sd)
    out = m(SparseTensor(x.to(DEV), coords.to(DEV)), t.to(DEV), cond.to(DEV))
    err = _rel(out.feats, ref)
    print(f"\nslat flow at the shipped widths (depth 2, {n} voxels) vs the fp32 oracle: rel L2 {err:.2e}")
    assert err < 5e-3, err


def test_gaussian_decoder_matches_reference_class():
    """SLatGaussianDecoder (decoder_gs.py:10-122) on the static-VAE engine against the reference's own class (CPU fp32)."""
    from gvfdiffusion_b200.sparse.basic import SparseTensor
    from gvfdiffusion_b200.trellis.models import SLatGaussianDecoder
    GD = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "slat_decoder_gs_tiny.pt"), weights_only=False)
    m = SLatGaussianDecoder(**GD["cfg"], device=DEV).load_state_dict(GD["state_dict"])
    assert float((m.offset_perturbation.cpu() - GD["state_dict"]["offset_perturbation"]).abs().max()) < 1e-6
    x = SparseTensor(GD["latent"].to(DEV), GD["coords"].to(DEV))
    rows = m.decode_rows(x)
    err = _rel(rows.feats, GD["rows"])
    print(f"\nSLatGaussianDecoder rows vs the reference's own class (fp32): rel L2 {err:.2e}")
    assert err < 3e-3, err
    reps = m(x)
    assert len(reps) == len(GD["reps"])
    for rep, want, ip in zip(reps, GD["reps"], GD["init_params"]):
        assert rep.mininum_kernel_size == ip["mininum_kernel_size"] and rep.scaling_activation_type == ip["scaling_activation"]
        for name, w in want.items():
            got = getattr(rep, name).float().cpu().reshape(w.shape)
            assert float((got - w).abs().max()) < 5e-3 * max(1.0, float(w.abs().max())), name
    # to_representation alone, on the reference's rows: exact arithmetic of the kernel vs the reference's torch expressions
    reps = m.to_representation(SparseTensor(GD["rows"].to(DEV), GD["coords"].to(DEV)))
    for rep, want in zip(reps, GD["reps"]):
        for name, w in want.items():
            assert float((getattr(rep, name).cpu().reshape(w.shape) - w).abs().max()) < 1e-6, name


def test_pipeline_sample_slat_and_decode_slat():
    """trellis_image_to_3d.py:197-256 on the tiny models: guidance-interval sampling over the flow model (graph replay on),
    de-normalisation, Gaussian decoding; against the same chain composed by hand from the oracle-checked pieces."""
    from gvfdiffusion_b200.sparse.basic import SparseTensor
    from gvfdiffusion_b200.trellis.models import SLatGaussianDecoder
    from gvfdiffusion_b200.trellis.pipelines.samplers import FlowEulerGuidanceIntervalSampler
    from gvfdiffusion_b200.trellis.pipelines.trellis_image_to_3d import TrellisImageTo3DPipeline
    GD = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "slat_decoder_gs_tiny.pt"), weights_only=False)
    flow = _model(G["cfg"], G["state_dict"])
    flow.use_graphs = True
    dec = SLatGaussianDecoder(**GD["cfg"], device=DEV).load_state_dict(GD["state_dict"])
    args = {"slat_sampler": {"name": "FlowEulerGuidanceIntervalSampler", "args": {"sigma_min": 1e-5},
                             "params": {"steps": 6, "cfg_strength": 3.0, "cfg_interval": [0.5, 1.0], "rescale_t": 3.0}},
            "slat_normalization": {"mean": [0.1 * i for i in range(8)], "std": [1.0 + 0.05 * i for i in range(8)]}}
    pipe = TrellisImageTo3DPipeline.from_args(args, {"slat_flow_model": flow, "slat_decoder_gs": dec}, device=DEV)
    cond = {"cond": G["cond"].to(DEV), "neg_cond": torch.zeros_like(G["cond"]).to(DEV)}
    coords = G["coords"].to(DEV)
    slat = pipe.sample_slat(cond, coords, noise=G["x"])
    # by hand, eager model calls
    flow2 = _model(G["cfg"], G["state_dict"])
    r = FlowEulerGuidanceIntervalSampler(1e-5).sample(flow2, SparseTensor(G["x"].to(DEV), coords), verbose=False, **cond,
                                                      **args["slat_sampler"]["params"])
    want = r.samples.feats * torch.tensor(args["slat_normalization"]["std"], device=DEV) + torch.tensor(args["slat_normalization"]["mean"], device=DEV)
    assert torch.allclose(slat.feats, want, rtol=0, atol=1e-6)
    out = pipe.decode_slat(slat, ["gaussian"])["gaussian"]
    assert len(out) == 2 and out[0]._xyz.shape == (int((coords[:, 0] == 0).sum()) * 4, 3)
    assert all(torch.isfinite(getattr(g_, n)).all() for g_ in out for n in ("_xyz", "_scaling", "_rotation", "_opacity"))
    with pytest.raises(NotImplementedError):
        pipe.decode_slat(slat, ["mesh"])


@pytest.mark.parametrize("case", ["p1", "p2"])
def test_sparse_structure_flow_matches_reference_class(case):
    """SparseStructureFlowModel (sparse_structure_flow.py:55-200) against the reference's own class (CPU fp32), patch sizes
    1 (shipped) and 2; eager and graph replay give the same bits."""
    from gvfdiffusion_b200.trellis.models import SparseStructureFlowModel
    GS = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sparse_structure_flow_tiny.pt"), weights_only=False)
    c = GS[case]
    m = SparseStructureFlowModel(**c["cfg"], device=DEV).load_state_dict(c["state_dict"])
    assert float((m.pos_emb.cpu() - c["pos_emb"]).abs().max()) < 1e-5
    x, t, cond = c["x"].to(DEV), c["t"].to(DEV), c["cond"].to(DEV)
    out = m(x, t, cond)
    err = _rel(out, c["out"])
    print(f"\nsparse-structure flow ({case}) vs the reference's own class (fp32): rel L2 {err:.2e}")
    assert out.shape == c["out"].shape and err < 5e-3, err
    m.use_graphs = True
    assert torch.equal(m(x, t, cond), out) and torch.equal(m(x, t, cond), out)


def test_pipeline_sample_sparse_structure_and_run_from_cond():
    """trellis_image_to_3d.py:165-195,279-284: occupancy-latent sampling over SparseStructureFlowModel, a caller-supplied
    occupancy decoder (here: nearest upsampling of one latent channel), argwhere -> coords -> sample_slat -> decode_slat."""
    from gvfdiffusion_b200.trellis.models import SLatGaussianDecoder, SparseStructureFlowModel
    from gvfdiffusion_b200.trellis.pipelines.trellis_image_to_3d import TrellisImageTo3DPipeline
    here = os.path.dirname(os.path.abspath(__file__))
    GS = torch.load(os.path.join(here, "golden", "sparse_structure_flow_tiny.pt"), weights_only=False)["p1"]
    GD = torch.load(os.path.join(here, "golden", "slat_decoder_gs_tiny.pt"), weights_only=False)
    ss = SparseStructureFlowModel(**GS["cfg"], device=DEV).load_state_dict(GS["state_dict"])
    flow = _model(G["cfg"], G["state_dict"])
    dec = SLatGaussianDecoder(**GD["cfg"], device=DEV).load_state_dict(GD["state_dict"])
    occ = lambda z: torch.nn.functional.interpolate(z[:, :1], scale_factor=2, mode="nearest") - 0.8      # 8^3 -> 16^3 logits
    euler = {"name": "FlowEulerGuidanceIntervalSampler", "args": {"sigma_min": 1e-5},
             "params": {"steps": 4, "cfg_strength": 5.0, "cfg_interval": [0.5, 1.0], "rescale_t": 3.0}}
    args = {"sparse_structure_sampler": euler, "slat_sampler": euler, "slat_normalization": {"mean": [0.0] * 8, "std": [1.0] * 8}}
    pipe = TrellisImageTo3DPipeline.from_args(args, {"sparse_structure_flow_model": ss, "sparse_structure_decoder": occ,
                                                     "slat_flow_model": flow, "slat_decoder_gs": dec}, device=DEV)
    cond = {"cond": G["cond"].to(DEV), "neg_cond": torch.zeros_like(G["cond"]).to(DEV)}
    torch.manual_seed(0)
    coords = pipe.sample_sparse_structure(cond, num_samples=2)
    assert coords.dtype == torch.int32 and coords.shape[1] == 4 and coords.shape[0] > 0
    assert int(coords[:, 0].max()) <= 1 and int(coords[:, 1:].max()) < 16
    assert torch.equal(coords[:, 0], coords[:, 0].sort().values)            # rows grouped by batch entry (argwhere order)
    torch.manual_seed(0)
    out = pipe.run_from_cond(cond, num_samples=2)["gaussian"]
    assert len(out) == int(coords[:, 0].max()) + 1
    assert sum(g_._xyz.shape[0] for g_ in out) == coords.shape[0] * 4


def test_graph_and_workspace_caches_stay_bounded_over_a_stream_of_objects():
    """Every object has its own token count: the per-count workspaces live in a small LRU, evicting one drops the graphs
    captured against it, and a revisited object is simply captured again -- results equal the eager path throughout."""
    from gvfdiffusion_b200.sparse.basic import SparseTensor
    m = _model(G["cfg"], G["state_dict"])
    m.use_graphs = True
    cond = G["cond"].to(DEV)
    g = torch.Generator().manual_seed(12)
    keep = []
    for i in range(7):
        n0, n1 = 260 - 17 * i, 140 - 9 * i                          # a different voxel set (and coarse token count) per object
        rows = torch.cat([torch.arange(n0), 260 + torch.arange(n1)])
        coords = G["coords"][rows].to(DEV)
        x = torch.randn(coords.shape[0], 8, generator=g).to(DEV)
        t = torch.tensor([300.0 + 50 * i, 700.0 - 50 * i], device=DEV)
        keep.append((coords, x, t))
        for _ in range(2):                                           # capture, then replay
            assert torch.equal(m(SparseTensor(x, coords), t, cond).feats, m.forward(SparseTensor(x, coords), t, cond).feats), i
        assert len(m._ws) <= m._MAX_WORKSPACES
        assert sum(len(e["graphs"]) for e in m._kv_cache.values()) <= m._MAX_WORKSPACES
    coords, x, t = keep[0]                                           # its workspace and graph are long gone
    assert torch.equal(m(SparseTensor(x, coords), t, cond).feats, m.forward(SparseTensor(x, coords), t, cond).feats)


def test_sparse_structure_decoder_matches_reference_class():
    """SparseStructureDecoder (sparse_structure_vae.py:209-306) on the voxel-side operators (dense grid = fully active
    sparse grid) against the reference's own class (CPU fp32): logits and the occupancy they decide."""
    from gvfdiffusion_b200.trellis.models import SparseStructureDecoder
    c = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sparse_structure_decoder_tiny.pt"), weights_only=False)
    m = SparseStructureDecoder(**c["cfg"], device=DEV).load_state_dict(c["state_dict"])
    out = m(c["z"].to(DEV))
    err = _rel(out, c["out"])
    agree = float(((out.cpu() > 0) == (c["out"] > 0)).float().mean())
    print(f"\nsparse-structure decoder vs the reference's own class (fp32): rel L2 {err:.2e}, occupancy agreement {agree:.4f}")
    assert out.shape == c["out"].shape and err < 5e-3, err
    assert agree > 0.995
    assert torch.equal(m(c["z"].to(DEV)), out)                       # cached grid / neighbour map: same bits
