"""Structured-latent flow model of the TRELLIS stage (reference trellis/models/structured_latent_flow.py,
sampled by trellis/pipelines/trellis_image_to_3d.py:223-256) at the widths of the shipped checkpoint
(slat_flow_img_dit_L_64l8p2_fp16: 64^3 grid, in / out 8, model 1024, 24 blocks, 16 heads, patch 2, io [128], cond 1024,
q / k RMS-norm, fp16 torso) on a synthetic shell of ~20 k active voxels with 1374 image tokens.

Times ONE model call (what the Euler sampler repeats 25 x 2 times with guidance) with CUDA events, next to a GPU stand-in
of the reference's execution in plain torch: fp16 nn.functional.linear (cuBLAS), flash_attn 2.8.3 varlen for the sparse
self-attention and the cross-attention (K / V projections recomputed every call, as the reference does), torch LayerNorm /
SiLU / scatter_reduce / index gathers, and -- spconv being absent from this image -- the submanifold convolution as an
index-gather im2col + cuBLAS matmul (stated in the output).  Prints one JSON line.

    python tools/slat_flow_bench.py [--blocks 24] [--steps 10] [--no-standin]
"""
import argparse
import json
import math
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
C, IO, HEADS, L, CC = 1024, 128, 16, 1374, 1024


def shell_coords(res=64, radius=22.0, half_thickness=1.6):
    ax = torch.arange(res, dtype=torch.float32) - (res - 1) / 2
    x, y, z = torch.meshgrid(ax, ax, ax, indexing="ij")
    keep = ((x * x + y * y + z * z).sqrt() - radius).abs() < half_thickness
    idx = keep.nonzero().int()
    return torch.cat([torch.zeros(idx.shape[0], 1, dtype=torch.int32), idx], 1)


def state_dict(blocks, seed=0):
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def lin(name, o, i, std=None):
        sd[name + ".weight"] = torch.randn(o, i, generator=g) * (std or 1.0 / math.sqrt(i))
        sd[name + ".bias"] = torch.randn(o, generator=g) * 0.02

    def res(p, cin, cout):
        sd[p + "norm1.weight"], sd[p + "norm1.bias"] = 1 + 0.1 * torch.randn(cin, generator=g), 0.1 * torch.randn(cin, generator=g)
        for nm, ci in (("conv1", cin), ("conv2", cout)):
            sd[p + nm + ".conv.weight"] = torch.randn(cout, 3, 3, 3, ci, generator=g) / math.sqrt(27 * ci)
            sd[p + nm + ".conv.bias"] = torch.randn(cout, generator=g) * 0.02
        lin(p + "emb_layers.1", 2 * cout, C, 0.02)
        if cin != cout:
            lin(p + "skip_connection", cout, cin)
    res("input_blocks.0.", IO, IO)
    res("input_blocks.1.", IO, C)
    res("out_blocks.0.", 2 * C, IO)
    res("out_blocks.1.", 2 * IO, IO)
    lin("t_embedder.mlp.0", C, 256, 0.02)
    lin("t_embedder.mlp.2", C, C, 0.02)
    lin("input_layer", IO, 8)
    lin("out_layer", 8, IO)
    for i in range(blocks):
        p = f"blocks.{i}."
        sd[p + "norm2.weight"], sd[p + "norm2.bias"] = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
        lin(p + "self_attn.to_qkv", 3 * C, C)
        sd[p + "self_attn.q_rms_norm.gamma"] = 1 + 0.1 * torch.randn(HEADS, C // HEADS, generator=g)
        sd[p + "self_attn.k_rms_norm.gamma"] = 1 + 0.1 * torch.randn(HEADS, C // HEADS, generator=g)
        lin(p + "self_attn.to_out", C, C)
        lin(p + "cross_attn.to_q", C, C)
        lin(p + "cross_attn.to_kv", 2 * C, CC)
        lin(p + "cross_attn.to_out", C, C)
        lin(p + "mlp.mlp.0", 4 * C, C)
        lin(p + "mlp.mlp.2", C, 4 * C)
        lin(p + "adaLN_modulation.1", 6 * C, C, 0.02)
    return sd


class StandIn:
    """The reference's forward, op for op, in plain torch fp16 on the GPU (see the module docstring)."""

    def __init__(self, sd, blocks, dev):
        self.sd = {k: (v.to(dev).half() if ("t_embedder" not in k and "out_layer" not in k) else v.to(dev)) for k, v in sd.items()}
        self.blocks, self.dev = blocks, dev
        self.plans = {}

    def _nbr(self, coords):
        key = (coords.data_ptr(), coords.shape[0])
        if key not in self.plans:
            from gvfdiffusion_b200 import ops
            grid = int(coords[:, 1:].max()) + 1
            self.plans[key] = ops.sparse_neighbor_map(coords, 1, grid, 3, 1).long()       # spconv caches its indice pairs too
        return self.plans[key]

    def conv(self, p, x, coords):
        nbr = self._nbr(coords)
        w = self.sd[p + "conv.weight"].reshape(self.sd[p + "conv.weight"].shape[0], -1)
        xp = torch.cat([x, x.new_zeros(1, x.shape[1])], 0)
        cols = xp[nbr.clamp_min(-1)].reshape(x.shape[0], -1)                               # -1 -> the appended zero row
        return F.linear(cols, w, self.sd[p + "conv.bias"])

    def res(self, p, x, coords, emb):
        sd = self.sd
        emb_out = F.linear(F.silu(emb), sd[p + "emb_layers.1.weight"], sd[p + "emb_layers.1.bias"])
        scale, shift = emb_out.chunk(2, dim=1)
        h = F.layer_norm(x.float(), x.shape[-1:], sd[p + "norm1.weight"].float(), sd[p + "norm1.bias"].float(), 1e-6).half()
        h = self.conv(p + "conv1.", F.silu(h), coords)
        h = F.layer_norm(h.float(), h.shape[-1:], None, None, 1e-6).half() * (1 + scale) + shift
        h = self.conv(p + "conv2.", F.silu(h), coords)
        if p + "skip_connection.weight" in sd:
            x = F.linear(x, sd[p + "skip_connection.weight"], sd[p + "skip_connection.bias"])
        return h + x

    def block(self, p, x, emb, cond):
        from flash_attn import flash_attn_varlen_kvpacked_func, flash_attn_varlen_qkvpacked_func
        sd, n = self.sd, x.shape[0]
        d = C // HEADS
        mod = F.linear(F.silu(emb), sd[p + "adaLN_modulation.1.weight"], sd[p + "adaLN_modulation.1.bias"])
        s1, c1, g1, s2, c2, g2 = mod.chunk(6, dim=1)
        ln = lambda t, w=None, b=None: F.layer_norm(t.float(), t.shape[-1:], w, b, 1e-6).half()
        cu = torch.tensor([0, n], dtype=torch.int32, device=x.device)
        h = ln(x) * (1 + c1) + s1
        qkv = F.linear(h, sd[p + "self_attn.to_qkv.weight"], sd[p + "self_attn.to_qkv.bias"]).reshape(n, 3, HEADS, d)
        q, k, v = qkv.unbind(1)
        rms = lambda t, gm: (F.normalize(t.float(), dim=-1) * gm.float() * d ** 0.5).half()
        qkv = torch.stack([rms(q, sd[p + "self_attn.q_rms_norm.gamma"]), rms(k, sd[p + "self_attn.k_rms_norm.gamma"]), v], 1)
        h = flash_attn_varlen_qkvpacked_func(qkv, cu, n).reshape(n, C)
        x = x + F.linear(h, sd[p + "self_attn.to_out.weight"], sd[p + "self_attn.to_out.bias"]) * g1
        h = ln(x, sd[p + "norm2.weight"].float(), sd[p + "norm2.bias"].float())
        q = F.linear(h, sd[p + "cross_attn.to_q.weight"], sd[p + "cross_attn.to_q.bias"]).reshape(n, HEADS, d)
        kv = F.linear(cond[0], sd[p + "cross_attn.to_kv.weight"], sd[p + "cross_attn.to_kv.bias"]).reshape(-1, 2, HEADS, d)
        cuk = torch.tensor([0, kv.shape[0]], dtype=torch.int32, device=x.device)
        h = flash_attn_varlen_kvpacked_func(q, kv, cu, cuk, n, kv.shape[0]).reshape(n, C)
        x = x + F.linear(h, sd[p + "cross_attn.to_out.weight"], sd[p + "cross_attn.to_out.bias"])
        h = ln(x) * (1 + c2) + s2
        h = F.gelu(F.linear(h, sd[p + "mlp.mlp.0.weight"], sd[p + "mlp.mlp.0.bias"]), approximate="tanh")
        return x + F.linear(h, sd[p + "mlp.mlp.2.weight"], sd[p + "mlp.mlp.2.bias"]) * g2

    @torch.no_grad()
    def __call__(self, x, coords, plan, t, cond, pos):
        sd = self.sd
        half = 128
        freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32, device=x.device) / half)
        a = t[:, None].float() * freqs[None]
        emb = F.linear(F.silu(F.linear(torch.cat([a.cos(), a.sin()], -1), sd["t_embedder.mlp.0.weight"], sd["t_embedder.mlp.0.bias"])),
                       sd["t_embedder.mlp.2.weight"], sd["t_embedder.mlp.2.bias"]).half()
        cond = cond.half()
        h = F.linear(x, sd["input_layer.weight"].float(), sd["input_layer.bias"].float()).half()
        h = self.res("input_blocks.0.", h, coords, emb)
        s0 = h
        idx = plan["idx"].long()
        cells = plan["coords"].shape[0]
        h = torch.scatter_reduce(torch.zeros(cells, h.shape[1], device=h.device, dtype=h.dtype), 0,
                                 idx[:, None].expand(-1, h.shape[1]), h, reduce="mean")
        h = self.res("input_blocks.1.", h, plan["coords"], emb)
        s1 = h
        h = h + pos.half()
        for i in range(self.blocks):
            h = self.block(f"blocks.{i}.", h, emb, cond)
        h = torch.cat([h, s1], 1)[idx]
        h = self.res("out_blocks.0.", h, coords, emb)
        h = self.res("out_blocks.1.", torch.cat([h, s0], 1), coords, emb)
        h = F.layer_norm(h.float(), h.shape[-1:])
        return F.linear(h, sd["out_layer.weight"], sd["out_layer.bias"])


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=24)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--no-standin", action="store_true")
    ap.add_argument("--stage", action="store_true", help="time sample_slat (25 steps, guidance interval) + decode_slat")
    ap.add_argument("--conv-ab", action="store_true", help="time the ResBlock convolutions as gather-fused GEMM vs im2col + GEMM")
    ap.add_argument("--profile-one", action="store_true", help="one eager call between cudaProfilerStart / Stop (for "
                    "ncu --profile-from-start off), no timing")
    a = ap.parse_args()
    from gvfdiffusion_b200 import ops
    from gvfdiffusion_b200.sparse.basic import SparseTensor
    from gvfdiffusion_b200.sparse.spatial import downsample_plan
    from gvfdiffusion_b200.trellis.models import SLatFlowModel
    dev = torch.device("cuda:0")
    cfg = dict(resolution=64, in_channels=8, model_channels=C, cond_channels=CC, out_channels=8, num_blocks=a.blocks,
               num_heads=HEADS, mlp_ratio=4, patch_size=2, num_io_res_blocks=2, io_block_channels=[IO], pe_mode="ape",
               use_fp16=True, qk_rms_norm=True)
    sd = state_dict(a.blocks)
    model = SLatFlowModel(**cfg, device=dev).load_state_dict(sd)
    coords = shell_coords().to(dev)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(coords.shape[0], 8, generator=g).to(dev)
    cond = torch.randn(1, L, CC, generator=g).to(dev)
    t = torch.tensor([500.0], device=dev)
    st = SparseTensor(x, coords)
    out = model(st, t, cond).feats
    plan = downsample_plan(st, 2)
    n, nc = coords.shape[0], plan["coords"].shape[0]
    if a.stage:
        from gvfdiffusion_b200.trellis.models import SLatGaussianDecoder
        from gvfdiffusion_b200.trellis.pipelines.trellis_image_to_3d import TrellisImageTo3DPipeline
        DC, DH, DB, NG = 768, 12, 12, 8
        dsd = {}
        gg = torch.Generator().manual_seed(3)
        for nm, (o, i) in {"input_layer": (DC, 8), "out_layer": (14 * NG, DC)}.items():
            dsd[nm + ".weight"], dsd[nm + ".bias"] = torch.randn(o, i, generator=gg) / math.sqrt(i), torch.zeros(o)
        for b in range(DB):
            for nm, (o, i) in {"attn.to_qkv": (3 * DC, DC), "attn.to_out": (DC, DC), "mlp.mlp.0": (4 * DC, DC), "mlp.mlp.2": (DC, 4 * DC)}.items():
                dsd[f"blocks.{b}.{nm}.weight"], dsd[f"blocks.{b}.{nm}.bias"] = torch.randn(o, i, generator=gg) / math.sqrt(i), torch.zeros(o)
        rep_cfg = {"lr": {"_xyz": 1.0, "_features_dc": 1.0, "_opacity": 1.0, "_scaling": 1.0, "_rotation": 0.1},
                   "perturb_offset": True, "voxel_size": 1.5, "num_gaussians": NG, "2d_filter_kernel_size": 0.1,
                   "3d_filter_kernel_size": 9e-4, "scaling_bias": 4e-3, "opacity_bias": 0.1, "scaling_activation": "softplus"}
        dec = SLatGaussianDecoder(resolution=64, model_channels=DC, latent_channels=8, num_blocks=DB, num_heads=DH, use_fp16=True,
                                  representation_config=rep_cfg, device=dev).load_state_dict(dsd)
        model.use_graphs = True
        args = {"slat_sampler": {"name": "FlowEulerGuidanceIntervalSampler", "args": {"sigma_min": 1e-5},
                                 "params": {"steps": 25, "cfg_strength": 3.0, "cfg_interval": [0.5, 1.0], "rescale_t": 3.0}},
                "slat_normalization": {"mean": [0.0] * 8, "std": [1.0] * 8}}
        # the sparse-structure half at the shipped sizes: dense DiT over the 16^3 x 8 occupancy latent (4096 tokens x 1024,
        # 24 blocks), 25 steps with guidance 7.5 on [0.5, 1], then the Conv3d decoder [512, 128, 32] to 64^3 logits
        from gvfdiffusion_b200.trellis.models import SparseStructureDecoder, SparseStructureFlowModel
        ssd = {k: v for k, v in sd.items() if k.startswith("blocks.") or k.startswith("t_embedder.")}
        ssd["input_layer.weight"], ssd["input_layer.bias"] = torch.randn(C, 8, generator=gg) / math.sqrt(8), torch.zeros(C)
        ssd["out_layer.weight"], ssd["out_layer.bias"] = torch.randn(8, C, generator=gg) / math.sqrt(C), torch.zeros(8)
        ss = SparseStructureFlowModel(resolution=16, in_channels=8, model_channels=C, cond_channels=CC, out_channels=8,
                                      num_blocks=a.blocks, num_heads=HEADS, patch_size=1, use_fp16=True, qk_rms_norm=True,
                                      device=dev).load_state_dict(ssd)
        ss.use_graphs = True
        chs = [512, 128, 32]
        dd = {}

        def dconv(name, o, i, k=3):
            dd[name + ".weight"], dd[name + ".bias"] = torch.randn(o, i, k, k, k, generator=gg) / math.sqrt(i * k ** 3), torch.zeros(o)

        def dres(p_, c_):
            for nm in ("norm1", "norm2"):
                dd[p_ + nm + ".weight"], dd[p_ + nm + ".bias"] = torch.ones(c_), torch.zeros(c_)
            dconv(p_ + "conv1", c_, c_)
            dconv(p_ + "conv2", c_, c_)
        dconv("input_layer", chs[0], 8)
        for i_ in range(2):
            dres(f"middle_block.{i_}.", chs[0])
        bi_ = 0
        for lv, c_ in enumerate(chs):
            for _ in range(2):
                dres(f"blocks.{bi_}.", c_)
                bi_ += 1
            if lv < len(chs) - 1:
                dconv(f"blocks.{bi_}.conv", chs[lv + 1] * 8, c_)
                bi_ += 1
        dd["out_layer.0.weight"], dd["out_layer.0.bias"] = torch.ones(chs[-1]), torch.zeros(chs[-1])
        dconv("out_layer.2", 1, chs[-1])
        ssdec = SparseStructureDecoder(out_channels=1, latent_channels=8, num_res_blocks=2, channels=chs, device=dev).load_state_dict(dd)
        pipe = TrellisImageTo3DPipeline.from_args(args, {"slat_flow_model": model, "slat_decoder_gs": dec,
                                                         "sparse_structure_flow_model": ss, "sparse_structure_decoder": ssdec}, device=dev)
        pipe.sparse_structure_sampler = pipe.slat_sampler
        pipe.sparse_structure_sampler_params = {"steps": 25, "cfg_strength": 7.5, "cfg_interval": [0.5, 1.0], "rescale_t": 3.0}
        cd = {"cond": cond, "neg_cond": torch.zeros_like(cond)}
        calls = [0]
        fwd = model.forward_graphed
        def counted(*aa):
            calls[0] += 1
            return fwd(*aa)
        model.forward_graphed = counted
        pipe.sample_slat(cd, coords, noise=x)
        ncalls, calls[0] = calls[0], 0
        ms_sample = timed(lambda: pipe.sample_slat(cd, coords, noise=x), 3, warmup=1)
        slat = pipe.sample_slat(cd, coords, noise=x)
        ms_dec = timed(lambda: pipe.decode_slat(slat), 5, warmup=2)
        cd_ss = {"cond": cond, "neg_cond": torch.zeros_like(cond)}
        zn = torch.randn(1, 8, 16, 16, 16, generator=g).to(dev)
        pipe.sample_sparse_structure(cd_ss, noise=zn)
        ms_ss = timed(lambda: pipe.sample_sparse_structure(cd_ss, noise=zn), 3, warmup=1)
        z_lat = torch.randn(1, 8, 16, 16, 16, generator=g).to(dev)
        ms_ssdec = timed(lambda: ssdec(z_lat), 5, warmup=2)
        ms_ssflow = timed(lambda: ss(zn, t, cond), 10, warmup=3)
        print(json.dumps({"metric": "TRELLIS sampling stages at the shipped sizes (random weights): sample_sparse_structure + sample_slat + decode_slat",
                          "value": ms_ss + ms_sample + ms_dec, "unit": "ms", "higher_is_better": False,
                          "sample_sparse_structure_ms": ms_ss, "sparse_structure_flow_call_ms": ms_ssflow,
                          "sparse_structure_decoder_ms": ms_ssdec, "sample_slat_ms": ms_sample,
                          "model_calls": ncalls, "decode_slat_ms": ms_dec, "voxels": n, "gaussians": n * NG,
                          "coarse_tokens": nc, "cond_tokens": L}))
        return
    if a.conv_ab:
        from gvfdiffusion_b200.sparse.conv import SparseConv3d
        res = {}
        lv = {"fine": st, "coarse": SparseTensor(torch.zeros(nc, 8, device=dev), plan["coords"])}
        for name, level, cin, cout in (("fine 128->128", "fine", 128, 128), ("fine 256->128", "fine", 256, 128),
                                       ("coarse 128->1024", "coarse", 128, 1024), ("coarse 1024->1024", "coarse", 1024, 1024),
                                       ("fine 2048->128 (upsampled input, materialised)", "fine", 2048, 128)):
            s_ = lv[level]
            rows = s_.coords.shape[0]
            conv = SparseConv3d(cin, cout, 3, device=dev)
            conv.weight = (torch.randn(cout, 27 * cin, generator=g) / math.sqrt(27 * cin)).half().to(dev)
            conv.bias = torch.zeros(cout, device=dev)
            xa = torch.randn(rows, cin, generator=g).half().to(dev)
            nbr = conv.neighbor_map(s_)
            fused = timed(lambda: ops.sparse_conv_gemm(xa, nbr, conv.weight, conv.bias), 5)
            entry = {"rows": rows, "gflop": 2 * rows * 27 * cin * cout / 1e9, "fused_ms": fused}
            if rows * 27 * cin * 2 < (3 << 30):
                entry["im2col_gemm_ms"] = timed(lambda: ops.gemm(ops.sparse_im2col(xa, nbr), conv.weight, conv.bias, ops.EPI_F16), 5)
            res[name] = entry
        # the upsampled convolution as coarse per-tap products + gather-sum
        tap_w = (torch.randn(27 * 128, 2048, generator=g) / math.sqrt(27 * 2048)).half().to(dev)
        ac = torch.randn(nc, 2048, generator=g).half().to(dev)
        nbr = SparseConv3d(2048, 128, 3, device=dev).neighbor_map(st)
        res["fine 2048->128 (upsampled input) as tap products + gather-sum"] = {
            "gemm_ms": timed(lambda: ops.gemm(ac, tap_w, None, ops.EPI_F32), 5),
            "total_ms": timed(lambda: ops.sparse_tap_gather_sum(ops.gemm(ac, tap_w, None, ops.EPI_F32), nbr, plan["idx"], None), 5)}
        print(json.dumps({"conv_ab": res, "voxels": n, "coarse_tokens": nc}))
        return
    if a.profile_one:
        model.forward(st, t, cond)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        model.forward(st, t, cond)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    ms_eager = timed(lambda: model.forward(st, t, cond), a.steps)
    model.use_graphs = True
    assert torch.equal(model(st, t, cond).feats, out)
    ms = timed(lambda: model(st, t, cond), a.steps)
    # algorithmic flops of one call: transformer GEMMs + attention cores + the four ResBlocks' convolutions / skips
    blk = 2 * nc * C * C * (3 + 1 + 1 + 1 + 8) + 4 * nc * nc * C + 4 * nc * L * C
    conv = 2 * 27 * (n * IO * IO * 2 + nc * (IO * C + C * C) + n * (2 * C * IO + IO * IO) + n * (2 * IO * IO + IO * IO))
    flop = a.blocks * blk + conv + 2 * nc * IO * C + 2 * nc * 2 * C * IO + 2 * n * 2 * IO * IO
    line = {"metric": "structured-latent flow model call (TRELLIS stage, shipped widths)", "value": ms, "unit": "ms",
            "higher_is_better": False, "ms_eager": ms_eager, "execution": "CUDA-graph replay", "voxels": n, "coarse_tokens": nc, "cond_tokens": L, "blocks": a.blocks,
            "algorithmic_tflop": flop / 1e12, "tflops": flop / ms / 1e9, "dtype": "f16 (fp32 accumulate / residual stream)",
            "note": "K / V of the image tokens cached across calls (computed once per conditioning tensor)"}
    if not a.no_standin:
        ref = StandIn(sd, a.blocks, dev)
        pos = ops.ape(plan["coords"][:, 1:].float().contiguous(), C)
        r = ref(x, coords, plan, t, cond, pos)
        line["standin_rel_l2"] = float((out - r).norm() / r.norm())
        ms_ref = timed(lambda: ref(x, coords, plan, t, cond, pos), a.steps)
        line["standin"] = {"value": ms_ref, "unit": "ms", "kind": "stand-in",
                           "what": "reference forward in plain torch fp16: cuBLAS F.linear, flash_attn 2.8.3 varlen self / cross "
                                   "attention (K / V recomputed per call), torch LayerNorm / SiLU / scatter_reduce / gathers; "
                                   "submanifold conv as index-gather im2col + cuBLAS (spconv is absent from this image)"}
        line["speedup_vs_standin"] = ms_ref / ms
    print(json.dumps(line))


if __name__ == "__main__":
    main()
