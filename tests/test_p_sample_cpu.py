"""Product-side ancestral sampler (BASELINE configs[0]: gaussian_diffusion.p_sample on a 4x8^3 dense
latent, CPU fp32) against fixtures recorded from the reference's own `create_gaussian_diffusion(...).p_sample`
(tests/golden/make_golden.py: gen_p_sample, gen_respace)."""
import os

import torch

from gvfdiffusion_b200.model import gaussian_diffusion as GD
from gvfdiffusion_b200.model.respace import SpacedDiffusion, create_gaussian_diffusion, space_timesteps

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DIFF_CFG = dict(steps=1000, learn_sigma=False, sigma_small=False, use_kl=False, noise_schedule="cosine",
                predict_type="v", predict_xstart=False, rescale_timesteps=True, rescale_learned_sigmas=True)


def _model(W):
    return lambda x, ts: torch.einsum("oc,bcdhw->bodhw", W, x) * torch.cos(ts / 1000.0).view(-1, 1, 1, 1, 1)


def test_p_sample_config0_matches_reference():
    g = torch.load(os.path.join(GOLD, "p_sample.pt"), weights_only=False)
    d = create_gaussian_diffusion(**DIFF_CFG)
    assert isinstance(d, SpacedDiffusion) and d.num_timesteps == 1000 and d.timestep_map == list(range(1000))
    sched = torch.load(os.path.join(GOLD, "schedule.pt"), weights_only=False)
    assert torch.equal(torch.from_numpy(d.betas), sched["betas"])            # float64 tables, bit for bit
    for tt, o in g["outs"].items():
        torch.manual_seed(100 + tt)                                          # the fixture drew its noise under this seed
        r = d.p_sample(_model(g["W"]), g["x"], torch.tensor([tt]))
        assert torch.equal(r["pred_xstart"], o["pred_xstart"]), tt
        assert torch.equal(r["sample"], o["sample"]), tt


def test_space_timesteps_and_respaced_p_sample_match_reference():
    g = torch.load(os.path.join(GOLD, "respace.pt"), weights_only=False)
    for (n, sec), want in g["space"].items():
        arg = eval(sec) if sec.startswith("[") else sec
        assert sorted(space_timesteps(n, arg)) == want, (n, sec)
    for case in g["cases"]:
        d = create_gaussian_diffusion(**case["cfg"])
        assert d.timestep_map == case["timestep_map"]
        assert torch.equal(torch.from_numpy(d.betas), case["betas"])
        for (tt, clip), o in case["outs"].items():
            t = torch.tensor([tt, max(tt - 1, 0)])
            torch.manual_seed(7 + tt)
            r = d.p_sample(_model(g["W"]), g["x"], t, clip_denoised=clip)
            assert torch.equal(r["pred_xstart"], o["pred_xstart"]), (case["cfg"]["predict_type"], tt, clip)
            assert torch.equal(r["sample"], o["sample"]), (case["cfg"]["predict_type"], tt, clip)


def test_p_sample_loop_runs_and_is_deterministic_given_noise():
    d = create_gaussian_diffusion(**dict(DIFF_CFG, timestep_respacing="8"))
    W = torch.eye(4) * 0.3

    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.p = torch.nn.Parameter(torch.zeros(1))

        def forward(self, x, ts):
            return _model(W)(x, ts)

    torch.manual_seed(0)
    a = d.p_sample_loop(M(), (1, 4, 8, 8, 8), noise=torch.ones(1, 4, 8, 8, 8))
    torch.manual_seed(0)
    b = d.p_sample_loop(M(), (1, 4, 8, 8, 8), noise=torch.ones(1, 4, 8, 8, 8))
    assert a.shape == (1, 4, 8, 8, 8) and torch.equal(a, b) and torch.isfinite(a).all()


def test_learned_sigma_raises():
    d = GD.GaussianDiffusion(betas=GD.get_named_beta_schedule("linear", 1000), model_mean_type=GD.ModelMeanType.EPSILON,
                             model_var_type=GD.ModelVarType.LEARNED_RANGE)
    try:
        d.p_mean_variance(lambda x, t: x, torch.zeros(1, 2), torch.tensor([3]))
    except NotImplementedError:
        return
    raise AssertionError("learned variances must raise")
