"""The PRODUCT's host-side DPM-Solver scalars (gvfdiffusion_b200/model/dpmsolver.py: numpy float32 in the
reference's operation order, no device syncs) against the fixture produced by the reference's own
NoiseScheduleVP (tests/golden/schedule.pt) -- checked without a GPU: only the elementwise state updates run on
the device, everything that decides the coefficients is here."""
import os

import numpy as np
import torch

from gvfdiffusion_b200.model.dpmsolver import NoiseScheduleVP

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_noise_schedule_host_mirror_matches_reference_fixture():
    g = torch.load(os.path.join(G, "schedule.pt"), weights_only=False)
    ns = NoiseScheduleVP("discrete", betas=g["betas"])
    assert ns.total_N == g["total_N"] == 996
    assert np.array_equal(ns.log_alpha_array, g["log_alpha_array"].reshape(-1).numpy())
    assert np.array_equal(ns.t_array, g["t_array"].reshape(-1).numpy())
    for i, t in enumerate(g["ts"].tolist()):
        la, lam, std = float(g["log_alpha"][i]), float(g["lambda"][i]), float(g["std"][i])
        assert float(ns.marginal_log_mean_coeff(t)) == la                      # interpolation: bit-exact
        # std = sqrt(1 - exp(2 log_alpha)) cancels as t -> 0 (exp(..) ~ 0.9999): one ulp of numpy's vs torch's
        # float32 exp moves std by ~3e-6 and lambda by ~3e-4 there; elsewhere both are bit-exact
        assert abs(float(ns.marginal_std(t)) - std) <= 5e-6
        assert abs(float(ns.marginal_lambda(t)) - lam) <= (1e-3 if t < 0.005 else 2e-6)
    for lam, t_ref in zip(g["inv_lambda_in"].tolist(), g["inv_lambda"].reshape(-1).tolist()):
        assert abs(float(ns.inverse_lambda(lam)) - t_ref) <= 2e-6


def test_time_steps_and_coefficients_are_finite_and_ordered():
    g = torch.load(os.path.join(G, "schedule.pt"), weights_only=False)
    ns = NoiseScheduleVP("discrete", betas=g["betas"])
    ts = torch.linspace(1.0, 1.0 / ns.total_N, 33).numpy().astype(np.float32)      # the 32-step grid of the benchmark
    lam = np.array([ns.marginal_lambda(t) for t in ts])
    assert np.all(np.isfinite(lam)) and np.all(np.diff(lam) > 0)                      # lambda increases as t decreases
    std = np.array([ns.marginal_std(t) for t in ts])
    assert np.all(np.diff(std) < 0) and std[0] < 1.0 and std[-1] > 0.0
