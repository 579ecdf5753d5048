"""CPU oracle: noise schedule, model wrapper, DPM-Solver++ and IDDPM p_sample
(TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates, op for op in torch-CPU fp32 (tables in float64 numpy exactly where the
reference uses them):
  * cosine betas                     model/gaussian_diffusion.py:52-56,73-89
  * NoiseScheduleVP('discrete')      model/dpmsolver.py:7-168  (log-alpha table, clip at
                                     lambda=-5.1, piece-wise linear interpolation :1270-1309)
  * model_wrapper (v-pred, 3-way CFG) model/dpmsolver.py:171-351
  * DPM_Solver.sample multistep order 2 / adaptive order 2   :564-609,813-869,973-1027,1064-1262
  * GaussianDiffusion.p_sample (V, FIXED_LARGE, dynamic thresholding)
                                     model/gaussian_diffusion.py:198-215,279-460, model/respace.py:112-170

Pinned against the reference classes by tests/golden/make_golden.py.
"""
import math

import numpy as np
import torch


# ---------------------------------------------------------------- betas
def cosine_betas(n=1000, max_beta=0.999):
    f = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
    return np.array([min(1 - f((i + 1) / n) / f(i / n), max_beta) for i in range(n)],
                    dtype=np.float64)


def spaced_betas(betas):
    """SpacedDiffusion with every timestep kept still re-derives betas from the cumulative
    product (model/respace.py:123-131); `diffusion.betas` handed to NoiseScheduleVP at
    inference_dpm_latent.py:156 are these (last-bit different from cosine_betas)."""
    ac = np.cumprod(1.0 - np.asarray(betas, dtype=np.float64), axis=0)
    out, last = [], 1.0
    for a in ac:
        out.append(1 - a / last)
        last = a
    return np.array(out, dtype=np.float64)


def reference_betas(n=1000):
    return spaced_betas(cosine_betas(n))


# ---------------------------------------------------------------- noise schedule
def interpolate_fn(x, xp, yp):
    """model/dpmsolver.py:1270-1309, x [N,1], xp/yp [1,K]."""
    N, K = x.shape[0], xp.shape[1]
    all_x = torch.cat([x.unsqueeze(2), xp.unsqueeze(0).repeat((N, 1, 1))], dim=2)
    sorted_all_x, x_indices = torch.sort(all_x, dim=2)
    x_idx = torch.argmin(x_indices, dim=2)
    cand_start_idx = x_idx - 1
    start_idx = torch.where(x_idx == 0, torch.tensor(1),
                            torch.where(x_idx == K, torch.tensor(K - 2), cand_start_idx))
    end_idx = torch.where(start_idx == cand_start_idx, start_idx + 2, start_idx + 1)
    start_x = torch.gather(sorted_all_x, 2, start_idx.unsqueeze(2)).squeeze(2)
    end_x = torch.gather(sorted_all_x, 2, end_idx.unsqueeze(2)).squeeze(2)
    start_idx2 = torch.where(x_idx == 0, torch.tensor(0),
                             torch.where(x_idx == K, torch.tensor(K - 2), cand_start_idx))
    ype = yp.unsqueeze(0).expand(N, -1, -1)
    start_y = torch.gather(ype, 2, start_idx2.unsqueeze(2)).squeeze(2)
    end_y = torch.gather(ype, 2, (start_idx2 + 1).unsqueeze(2)).squeeze(2)
    return start_y + (x - start_x) * (end_y - start_y) / (end_x - start_x)


class NoiseScheduleVP:
    def __init__(self, betas):
        betas = torch.as_tensor(betas)
        log_alphas = 0.5 * torch.log(1 - betas).cumsum(dim=0)
        # numerical_clip_alpha, clipped_lambda = -5.1
        log_sigmas = 0.5 * torch.log(1.0 - torch.exp(2.0 * log_alphas))
        lambs = log_alphas - log_sigmas
        idx = torch.searchsorted(torch.flip(lambs, [0]), -5.1)
        if idx > 0:
            log_alphas = log_alphas[:-idx]
        self.T = 1.0
        self.log_alpha_array = log_alphas.reshape(1, -1).to(torch.float32)
        self.total_N = self.log_alpha_array.shape[1]
        self.t_array = torch.linspace(0.0, 1.0, self.total_N + 1)[1:].reshape(1, -1).to(torch.float32)

    def marginal_log_mean_coeff(self, t):
        return interpolate_fn(t.reshape(-1, 1), self.t_array, self.log_alpha_array).reshape(-1)

    def marginal_alpha(self, t):
        return torch.exp(self.marginal_log_mean_coeff(t))

    def marginal_std(self, t):
        return torch.sqrt(1.0 - torch.exp(2.0 * self.marginal_log_mean_coeff(t)))

    def marginal_lambda(self, t):
        lm = self.marginal_log_mean_coeff(t)
        return lm - 0.5 * torch.log(1.0 - torch.exp(2.0 * lm))

    def inverse_lambda(self, lamb):
        log_alpha = -0.5 * torch.logaddexp(torch.zeros((1,)), -2.0 * lamb)
        t = interpolate_fn(log_alpha.reshape(-1, 1), torch.flip(self.log_alpha_array, [1]),
                           torch.flip(self.t_array, [1]))
        return t.reshape(-1)


def _ex(v, dims):
    return v[(...,) + (None,) * (dims - 1)]


# ---------------------------------------------------------------- model wrapper
def make_model_fn(model, ns, condition, unconditional_condition=None,
                  guidance_scale=1.0, guidance_scale2=1.0):
    """model_wrapper(..., model_type='v', guidance_type='classifier-free') :273-347.
    model(x, t_input, **cond) -> v."""

    def noise_pred_fn(x, t_continuous, cond):
        t_input = (t_continuous - 1.0 / ns.total_N) * 1000.0
        out = model(x, t_input, **cond)
        alpha_t, sigma_t = ns.marginal_alpha(t_continuous), ns.marginal_std(t_continuous)
        return _ex(alpha_t, x.dim()) * out + _ex(sigma_t, x.dim()) * x

    def model_fn(x, t_continuous):
        if (guidance_scale == 1.0 and guidance_scale2 == 1.0) or unconditional_condition is None:
            return noise_pred_fn(x, t_continuous, condition)
        x_in = torch.cat([x] * 3)
        t_in = torch.cat([t_continuous] * 3)
        full_uncond = dict(unconditional_condition)
        full_uncond["static_latent"] = torch.zeros_like(full_uncond["static_latent"])
        c_in = {k: torch.cat([full_uncond[k], unconditional_condition[k], condition[k]])
                for k in condition}
        e_fu, e_u, e_c = noise_pred_fn(x_in, t_in, c_in).chunk(3)
        return e_fu + guidance_scale * (e_u - e_fu) + guidance_scale2 * (e_c - e_u)

    return model_fn


# ---------------------------------------------------------------- DPM-Solver++
class DPMSolverPP:
    """algorithm_type='dpmsolver++', no x0 / xt correctors (inference_dpm_latent.py:236)."""

    def __init__(self, model_fn, ns):
        self.model = lambda x, t: model_fn(x, t.expand(x.shape[0]))
        self.ns = ns
        self.nfe = 0

    def data_prediction_fn(self, x, t):
        self.nfe += 1
        noise = self.model(x, t)
        alpha_t, sigma_t = self.ns.marginal_alpha(t), self.ns.marginal_std(t)
        return (x - sigma_t * noise) / alpha_t

    def first_update(self, x, s, t, model_s=None):
        ns = self.ns
        h = ns.marginal_lambda(t) - ns.marginal_lambda(s)
        sigma_s, sigma_t = ns.marginal_std(s), ns.marginal_std(t)
        alpha_t = torch.exp(ns.marginal_log_mean_coeff(t))
        phi_1 = torch.expm1(-h)
        if model_s is None:
            model_s = self.data_prediction_fn(x, s)
        return sigma_t / sigma_s * x - alpha_t * phi_1 * model_s, model_s

    def singlestep_second_update(self, x, s, t, r1=0.5, model_s=None):
        # model/dpmsolver.py:611-691, dpmsolver++ / solver_type 'dpmsolver'
        ns = self.ns
        lambda_s, lambda_t = ns.marginal_lambda(s), ns.marginal_lambda(t)
        h = lambda_t - lambda_s
        s1 = ns.inverse_lambda(lambda_s + r1 * h)
        sigma_s, sigma_s1, sigma_t = ns.marginal_std(s), ns.marginal_std(s1), ns.marginal_std(t)
        alpha_s1 = torch.exp(ns.marginal_log_mean_coeff(s1))
        alpha_t = torch.exp(ns.marginal_log_mean_coeff(t))
        phi_11 = torch.expm1(-r1 * h)
        phi_1 = torch.expm1(-h)
        if model_s is None:
            model_s = self.data_prediction_fn(x, s)
        x_s1 = (sigma_s1 / sigma_s) * x - (alpha_s1 * phi_11) * model_s
        model_s1 = self.data_prediction_fn(x_s1, s1)
        return ((sigma_t / sigma_s) * x - (alpha_t * phi_1) * model_s
                - (0.5 / r1) * (alpha_t * phi_1) * (model_s1 - model_s))

    def multistep_second_update(self, x, model_prev_list, t_prev_list, t):
        ns = self.ns
        m1, m0 = model_prev_list[-2], model_prev_list[-1]
        t1, t0 = t_prev_list[-2], t_prev_list[-1]
        l1, l0, lt = ns.marginal_lambda(t1), ns.marginal_lambda(t0), ns.marginal_lambda(t)
        sigma0, sigma_t = ns.marginal_std(t0), ns.marginal_std(t)
        alpha_t = torch.exp(ns.marginal_log_mean_coeff(t))
        h_0 = l0 - l1
        h = lt - l0
        r0 = h_0 / h
        D1_0 = (1.0 / r0) * (m0 - m1)
        phi_1 = torch.expm1(-h)
        return (sigma_t / sigma0) * x - (alpha_t * phi_1) * m0 - 0.5 * (alpha_t * phi_1) * D1_0

    def sample(self, x, steps=20, t_start=1.0, t_end=1e-3, order=2, method="multistep",
               lower_order_final=True, atol=0.0078, rtol=0.05):
        assert order == 2
        if method == "adaptive":
            return self._adaptive(x, t_start, t_end, atol=atol, rtol=rtol)
        assert method == "multistep" and steps >= order
        timesteps = torch.linspace(t_start, t_end, steps + 1)
        t = timesteps[0]
        t_prev = [t]
        m_prev = [self.data_prediction_fn(x, t)]
        t = timesteps[1]
        x, _ = self.first_update(x, t_prev[-1], t, model_s=m_prev[-1])
        t_prev.append(t)
        m_prev.append(self.data_prediction_fn(x, t))
        for step in range(2, steps + 1):
            t = timesteps[step]
            step_order = min(order, steps + 1 - step) if (lower_order_final and steps < 10) else order
            if step_order == 1:
                x, _ = self.first_update(x, t_prev[-1], t, model_s=m_prev[-1])
            else:
                x = self.multistep_second_update(x, m_prev, t_prev, t)
            t_prev[0], m_prev[0] = t_prev[1], m_prev[1]
            t_prev[1] = t
            if step < steps:
                m_prev[1] = self.data_prediction_fn(x, t)
        return x

    def _adaptive(self, x, t_T, t_0, h_init=0.05, atol=0.0078, rtol=0.05, theta=0.9, t_err=1e-5):
        ns = self.ns
        order = 2
        s = t_T * torch.ones((1,))
        lambda_s = ns.marginal_lambda(s)
        lambda_0 = ns.marginal_lambda(t_0 * torch.ones_like(s))
        h = h_init * torch.ones_like(s)
        x_prev = x
        while torch.abs(s - t_0).mean() > t_err:
            t = ns.inverse_lambda(lambda_s + h)
            x_lower, model_s = self.first_update(x, s, t)
            x_higher = self.singlestep_second_update(x, s, t, r1=0.5, model_s=model_s)
            delta = torch.max(torch.ones_like(x) * atol,
                              rtol * torch.max(torch.abs(x_lower), torch.abs(x_prev)))
            norm_fn = lambda v: torch.sqrt(torch.square(v.reshape(v.shape[0], -1)).mean(dim=-1, keepdim=True))
            E = norm_fn((x_higher - x_lower) / delta).max()
            if torch.all(E <= 1.0):
                x = x_higher
                s = t
                x_prev = x_lower
                lambda_s = ns.marginal_lambda(s)
            h = torch.min(theta * h * torch.float_power(E, -1.0 / order).float(), lambda_0 - lambda_s)
        return x


# ---------------------------------------------------------------- IDDPM p_sample (cfg 1)
class GaussianDiffusionV:
    """create_gaussian_diffusion(steps, noise_schedule='cosine', predict_type='v',
    rescale_timesteps=True) with no respacing: SpacedDiffusion == base diffusion."""

    def __init__(self, steps=1000):
        betas = reference_betas(steps)
        self.num_timesteps = steps
        self.betas = betas
        ac = np.cumprod(1.0 - betas)
        ac_prev = np.append(1.0, ac[:-1])
        self.sqrt_ac = np.sqrt(ac)
        self.sqrt_1mac = np.sqrt(1.0 - ac)
        self.post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
        self.coef1 = betas * np.sqrt(ac_prev) / (1.0 - ac)
        self.coef2 = (1.0 - ac_prev) * np.sqrt(1.0 - betas) / (1.0 - ac)
        self.logvar_large = np.log(np.append(self.post_var[1], betas[1:]))

    @staticmethod
    def _x(arr, t, shape):
        res = torch.from_numpy(arr)[t].float()
        while res.dim() < len(shape):
            res = res[..., None]
        return res.expand(shape)

    def p_sample(self, model, x, t, noise, clip_denoised=True, p=0.99):
        ts = t.float() * (1000.0 / self.num_timesteps)          # _WrappedModel, respace.py:165-170
        v = model(x, ts)
        x0 = self._x(self.sqrt_ac, t, x.shape) * x - self._x(self.sqrt_1mac, t, x.shape) * v
        if clip_denoised:                                        # dynamic_thresholding :198-215
            s = torch.quantile(x0.abs().reshape(x0.shape[0], -1), p, dim=-1)
            x0 = torch.clip(x0.reshape(x0.shape[0], -1).T, -s, s).T.reshape(x0.shape)
        mean = self._x(self.coef1, t, x.shape) * x0 + self._x(self.coef2, t, x.shape) * x
        logvar = self._x(self.logvar_large, t, x.shape)
        nz = (t != 0).float().view(-1, *([1] * (x.dim() - 1)))
        return {"sample": mean + nz * torch.exp(0.5 * logvar) * noise, "pred_xstart": x0}
