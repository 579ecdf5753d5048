"""GPU parity of the DiT / DPM-Solver / VAE-decode engines against the CPU oracle
(oracle pinned to the reference by tests/test_oracle_golden.py).

Tolerance (north_star: "within 1e-3 relative fp16 ... on latent eps-hat"): relative L2 error
<= 1e-3 against the oracle emulating the reference's fp16 autocast, and <= 3e-3 against the
fp32 oracle (the reference's own fp16 path sits ~5e-4 from its fp32 path, see the goldens)."""
import os

import pytest
import torch

from oracle import dit as ODIT
from oracle import dpm as ODPM
from oracle import vae as OVAE

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEV = "cuda"


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / b.norm())


def _dit_from_golden():
    from gvfdiffusion_b200.model.dit import DiT
    g = torch.load(os.path.join(G, "dit_tiny.pt"), weights_only=False)
    m = DiT(**g["cfg"])
    m.load_state_dict(g["state_dict"])        # reference state-dict names load as-is
    return g, m.to(DEV).eval()


def test_dit_forward_golden_fixture():
    g, m = _dit_from_golden()
    cond = {k: g[k].to(DEV) for k in ("cond_images", "static_latent", "deformation_position_xyz")}
    y = m(g["x"].to(DEV), g["t"].to(DEV), **cond)
    assert rel(y, g["y_autocast_fp16"]) < 1.5e-3       # the reference itself under (CPU) fp16 autocast
    assert rel(y, g["y_fp32"]) < 3e-3                  # the reference in fp32
    y16 = ODIT.dit_forward(g["state_dict"], g["x"], g["t"], g["cond_images"], g["static_latent"],
                           g["deformation_position_xyz"], g["cfg"]["num_heads"], "fp16")
    assert rel(y, y16) < 1e-3


@pytest.mark.parametrize("gs", [(1.0, 1.0), (2.0, 1.5)])
def test_dpm_multistep_golden_fixture(gs):
    from gvfdiffusion_b200.model import dpmsolver as D
    g, m = _dit_from_golden()
    ns = D.NoiseScheduleVP("discrete", betas=torch.from_numpy(ODPM.reference_betas(1000)))
    assert ns.total_N == 996
    cond = {k: g[k][:1].to(DEV) for k in ("cond_images", "static_latent", "deformation_position_xyz")}
    unc = dict(cond)
    unc["cond_images"] = torch.zeros_like(cond["cond_images"])
    fn = D.model_wrapper(m, ns, model_type="v", guidance_type="classifier-free", condition=cond,
                         unconditional_condition=unc, guidance_scale=gs[0], guidance_scale2=gs[1])
    name = "g1" if gs == (1.0, 1.0) else "cfg"
    for steps in (6, 12):
        s = D.DPM_Solver(fn, ns, algorithm_type="dpmsolver++")
        x = s.sample(g["noise"].to(DEV), steps=steps, t_start=1.0, t_end=1 / 1000, order=2,
                     skip_type="time_uniform", method="multistep")
        assert s.nfe == steps
        assert rel(x, g[f"sample_{name}_{steps}"]) < 3e-3, (steps, rel(x, g[f"sample_{name}_{steps}"]))
    eps = fn(g["noise"].to(DEV), torch.tensor([0.37]))
    if name == "g1":
        assert rel(eps, g["eps_g1_t0.37"]) < 2e-3


def test_dpm_adaptive_golden_fixture():
    from gvfdiffusion_b200.model import dpmsolver as D
    g, m = _dit_from_golden()
    ns = D.NoiseScheduleVP("discrete", betas=torch.from_numpy(ODPM.reference_betas(1000)))
    cond = {k: g[k][:1].to(DEV) for k in ("cond_images", "static_latent", "deformation_position_xyz")}
    fn = D.model_wrapper(m, ns, model_type="v", guidance_type="classifier-free", condition=cond,
                         unconditional_condition=None)
    s = D.DPM_Solver(fn, ns)
    x = s.sample(g["noise"].to(DEV), t_start=1.0, t_end=1 / 1000, order=2, method="adaptive")
    assert s.adaptive_nfe == 28                       # the reference's data-dependent step count
    assert rel(x, g["sample_adaptive"]) < 5e-3


def test_schedule_scalars_match_oracle():
    from gvfdiffusion_b200.model import dpmsolver as D
    betas = ODPM.reference_betas(1000)
    ns, ons = D.NoiseScheduleVP("discrete", betas=torch.from_numpy(betas)), ODPM.NoiseScheduleVP(betas)
    for t in (1.0, 0.96875, 0.5, 0.0321, 0.001):
        tt = torch.tensor([t])
        assert abs(float(ns.marginal_lambda(t)) - float(ons.marginal_lambda(tt))) < 2e-5
        assert abs(float(ns.marginal_std(t)) - float(ons.marginal_std(tt))) < 1e-6
        assert abs(float(ns.marginal_alpha(t)) - float(ons.marginal_alpha(tt))) < 1e-6
    for lam in (-5.0, 0.3, 4.0):
        assert abs(float(ns.inverse_lambda(lam)) - float(ons.inverse_lambda(torch.tensor([lam])))) < 1e-6


def test_dit_full_width_one_block_vs_oracle():
    """Benchmark widths (C=512, 16 heads of 32, 1370 image / 4096 static tokens), 2 blocks, T=3."""
    from gvfdiffusion_b200.model.dit import DiT
    torch.manual_seed(0)
    cfg = dict(resolution=512, in_channels=16, model_channels=512, static_cond_channels=14, image_cond_channels=1024,
               out_channels=16, num_blocks=2, num_heads=16, mlp_ratio=4, pe_mode="ape", qk_rms_norm=True,
               use_fp16=True, no_temporal_attn=False)
    m = DiT(**cfg)
    gen = torch.Generator().manual_seed(3)
    for p in m.parameters():
        if p.abs().sum() == 0:
            p.data = torch.randn(p.shape, generator=gen) * 0.02
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    B, T, N = 1, 3, 512
    x = torch.randn(B, T, N, 16, generator=gen)
    t = torch.tensor([431.7])
    ci = torch.randn(B, T, 1370, 1024, generator=gen)
    sl = torch.randn(B, 4096, 14, generator=gen)
    xyz = torch.rand(B, N, 3, generator=gen) - 0.5
    y = m.to(DEV)(x.to(DEV), t.to(DEV), ci.to(DEV), sl.to(DEV), xyz.to(DEV)).clone()
    y16 = ODIT.dit_forward(sd, x, t, ci, sl, xyz, 16, "fp16")
    y32 = ODIT.dit_forward(sd, x, t, ci, sl, xyz, 16, "fp32")
    assert rel(y, y16) < 1e-3, rel(y, y16)
    assert rel(y, y32) < 3e-3, rel(y, y32)
    # the optional fused residual-Linear + LayerNorm kernels take the same rounding points
    m.engine().fuse_resid_ln = True
    try:
        yf = m(x.to(DEV), t.to(DEV), ci.to(DEV), sl.to(DEV), xyz.to(DEV)).clone()
    finally:
        m.engine().fuse_resid_ln = False
    assert rel(yf, y16) < 1e-3, rel(yf, y16)
    assert rel(yf, y) < 5e-4, rel(yf, y)


def test_vae_decode_golden_fixture_and_full_width():
    from gvfdiffusion_b200.model.autoencoder import GSKLTemporalVariationalAutoEncoder as VAE
    g = torch.load(os.path.join(G, "vae_tiny.pt"), weights_only=False)
    v = VAE(**g["cfg"])
    v.load_state_dict(g["state_dict"])
    with torch.no_grad():                              # the inference engine (tests/test_vae_train_gpu.py: training path)
        d = v.to(DEV).decode(g["z"].to(DEV), g["queries"].to(DEV))
    assert d.shape == g["delta_fp32"].shape
    assert rel(d, g["delta_autocast_fp16"]) < 2e-3
    assert rel(d, g["delta_fp32"]) < 4e-3
    # shipped width: dim 768, 12 heads of 64, 2 layers, chunked queries
    torch.manual_seed(1)
    cfg = dict(depth=2, dim=768, queries_dim=768, output_dim=14, num_inputs=8192, num_latents=512, latent_dim=16,
               heads=12, dim_head=-1, num_timesteps=2, chunk_size=192)
    v = VAE(**cfg)
    gen = torch.Generator().manual_seed(2)
    for p in v.parameters():
        if p.abs().sum() == 0:
            p.data = torch.randn(p.shape, generator=gen) * 0.05
    sd = {k: t.clone() for k, t in v.state_dict().items()}
    z = torch.randn(2 * 2, 512, 16, generator=gen)
    q = torch.randn(2, 300, 14, generator=gen) * 0.3
    with torch.no_grad():
        d = v.to(DEV).decode(z.to(DEV), q.to(DEV))
    d16 = OVAE.vae_decode(sd, z, q, 12, 2, "fp16")
    assert rel(d, d16) < 1.5e-3, rel(d, d16)


def test_precomputed_modulation_table_is_bit_identical():
    """DiTEngine.precompute_modulation (the model times of a fixed-step DPM-Solver run are known in advance): NFEs replayed
    with the table row copied into the workspace give the same bits as NFEs that compute the timestep MLP + adaLN GEMV
    inside the graph; the solver announces its times, so a sampled latent is identical with the table on and off."""
    from gvfdiffusion_b200.model import dpmsolver as DPM
    g, m = _dit_from_golden()
    cond = {k: g[k].to(DEV) for k in ("cond_images", "static_latent", "deformation_position_xyz")}
    eng = m.engine()
    x = g["x"].to(DEV)

    def nfe(t):
        return m.forward_branches(x, t, [cond]).clone()

    times = [873.25, 500.0, 12.5, 999.0, 640.0, 333.0, 250.75, 100.0, 77.0, 1.0]       # > 8: two table launches
    eng.use_premod = False
    want = [nfe(t) for t in times]
    eng.use_premod = True
    m.precompute_modulation(times)
    assert len(eng._modtab) == len(times)
    for t, w in zip(times, want):
        assert torch.equal(nfe(t), w), t
    assert torch.equal(nfe(55.5), m.forward_branches(x, 55.5, [cond]))                    # a time outside the table still works
    # whole sampler runs, table on / off
    ns = DPM.NoiseScheduleVP("discrete", betas=torch.from_numpy(ODPM.reference_betas(1000)))
    outs = []
    for on in (False, True):
        eng.use_premod = on
        eng._modtab.clear()
        fn = DPM.model_wrapper(m, ns, model_type="v", guidance_type="classifier-free", condition=cond,
                               unconditional_condition=None, guidance_scale=1.0, guidance_scale2=1.0)
        outs.append(DPM.DPM_Solver(fn, ns).sample(x, steps=6, order=2, method="multistep").clone())
        assert (len(eng._modtab) == 6) == on
    assert torch.equal(outs[0], outs[1])
