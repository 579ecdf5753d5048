"""TRELLIS flow-matching Euler samplers (reference trellis/pipelines/samplers/flow_euler.py:11-199) on the
fused step kernel against samples recorded from the reference's own classes (tests/golden/make_golden.py::gen_flow_euler,
CPU fp32).  The toy model runs in torch on the GPU (tanh / matmul differ from the CPU in the last bits), so the comparison
is 1e-5 relative; the step itself is bit-exact against the torch expressions evaluated on the same device."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
G = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "flow_euler.pt"), weights_only=False)


def _model(W):
    return lambda x, t, c: torch.tanh(x @ W) * torch.cos(t / 1000.0).view(-1, 1, 1) + 0.1 * c


def test_samplers_match_reference_fixture():
    from gvfdiffusion_b200.trellis.pipelines.samplers import FlowEulerCfgSampler, FlowEulerGuidanceIntervalSampler, FlowEulerSampler
    W, noise, cond, neg = (G[k].to(DEV) for k in ("W", "noise", "cond", "neg_cond"))
    model = _model(W)
    runs = {"plain": lambda a: FlowEulerSampler(1e-5).sample(model, noise, cond, verbose=False, **a),
            "rescaled": lambda a: FlowEulerSampler(1e-5).sample(model, noise, cond, verbose=False, **a),
            "cfg": lambda a: FlowEulerCfgSampler(1e-5).sample(model, noise, cond, neg, verbose=False, **a),
            "interval": lambda a: FlowEulerGuidanceIntervalSampler(1e-5).sample(model, noise, cond, neg, verbose=False, **a)}
    for name, case in G["cases"].items():
        r = runs[name](case["args"])
        assert len(r.pred_x_t) == case["args"]["steps"]
        for got, want in ((r.samples, case["samples"]), (r.pred_x_0[-1], case["pred_x_0"])):
            err = float((got.cpu() - want).norm() / want.norm())
            assert err < 1e-5, (name, err)


def test_step_is_bit_exact_against_torch_expressions():
    from gvfdiffusion_b200.trellis.pipelines.samplers import FlowEulerCfgSampler
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, 7, 5, generator=g).to(DEV)
    vc, vn = torch.randn(3, 7, 5, generator=g).to(DEV), torch.randn(3, 7, 5, generator=g).to(DEV)
    calls = iter([vc, vn])
    model = lambda x_t, t, c: next(calls)
    s = FlowEulerCfgSampler(0.123)
    t, t_prev, cfg = 0.7312345, 0.6012345, 2.5
    out = s.sample_once(model, x, t, t_prev, "c", neg_cond="n", cfg_strength=cfg)
    v = (1 + cfg) * vc - cfg * vn
    assert torch.equal(out.pred_x_prev, x - (t - t_prev) * v)
    assert torch.equal(out.pred_x_0, (1 - 0.123) * x - (0.123 + (1 - 0.123) * t) * v)
