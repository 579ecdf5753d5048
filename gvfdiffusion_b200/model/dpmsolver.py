"""Host side of the DPM-Solver++ sampler: same names and call surface as the reference's
`model/dpmsolver.py` (NoiseScheduleVP :7, model_wrapper :171, DPM_Solver :354 with
`.sample(x, steps, t_start, t_end, order, skip_type, method)` :1064) for the configurations the
inference path uses (discrete VP schedule, v-prediction, classifier-free 3-way guidance,
dpmsolver++ multistep order <= 2 and adaptive order 2).

The reference evaluates every schedule quantity with ~20 tiny CUDA kernels per call
(interpolate_fn: cat + sort + gather, :1270-1309) roughly ten times per step.  Here the
scalars are computed on the host in float32 following the same operation order, and the
state update runs as one fused kernel per step (gvf_dpm_x0 / gvf_dpm_update).
"""
import math

import numpy as np
import torch

from .. import ops

f32 = np.float32


class NoiseScheduleVP:
    """reference model/dpmsolver.py:7-168, schedule='discrete' only."""

    def __init__(self, schedule="discrete", betas=None, alphas_cumprod=None, dtype=torch.float32):
        if schedule != "discrete":
            raise ValueError("only the 'discrete' schedule is on the GVFDiffusion inference path")
        self.schedule = schedule
        if betas is not None:
            betas = torch.as_tensor(betas).detach().cpu().double()
            log_alphas = 0.5 * torch.log(1 - betas).cumsum(dim=0)
        else:
            log_alphas = 0.5 * torch.log(torch.as_tensor(alphas_cumprod).detach().cpu().double())
        log_sigmas = 0.5 * torch.log(1.0 - torch.exp(2.0 * log_alphas))
        lambs = log_alphas - log_sigmas
        idx = int(torch.searchsorted(torch.flip(lambs, [0]), -5.1))        # numerical_clip_alpha :115-126
        if idx > 0:
            log_alphas = log_alphas[:-idx]
        self.T = 1.0
        self.log_alpha_array = log_alphas.to(torch.float32).numpy()
        self.total_N = int(self.log_alpha_array.shape[0])
        self.t_array = torch.linspace(0.0, 1.0, self.total_N + 1)[1:].to(torch.float32).numpy()

    @staticmethod
    def _interp(x, xp, yp):
        """interpolate_fn (:1270-1309) for a scalar x: piece-wise linear with linear extrapolation,
        float32 arithmetic in the reference's order."""
        x = f32(x)
        K = xp.shape[0]
        # position of x among the sorted keypoints (ties: x sorts first, as torch.sort is stable on cat([x, xp]))
        x_idx = int(np.searchsorted(xp, x, side="left"))
        if x_idx == 0:
            i0 = 0
        elif x_idx == K:
            i0 = K - 2
        else:
            i0 = x_idx - 1
        sx, ex, sy, ey = xp[i0], xp[i0 + 1], yp[i0], yp[i0 + 1]
        return f32(sy + f32(f32(x - sx) * f32(ey - sy)) / f32(ex - sx))

    def marginal_log_mean_coeff(self, t):
        return self._interp(t, self.t_array, self.log_alpha_array)

    def marginal_alpha(self, t):
        return f32(np.exp(self.marginal_log_mean_coeff(t)))

    def marginal_std(self, t):
        return f32(np.sqrt(f32(1.0) - np.exp(f32(2.0) * self.marginal_log_mean_coeff(t))))

    def marginal_lambda(self, t):
        lm = self.marginal_log_mean_coeff(t)
        return f32(lm - f32(0.5) * np.log(f32(1.0) - np.exp(f32(2.0) * lm)))

    def inverse_lambda(self, lamb):
        lamb = f32(lamb)
        log_alpha = f32(-0.5) * f32(np.logaddexp(f32(0.0), f32(-2.0) * lamb))
        return self._interp(log_alpha, self.log_alpha_array[::-1], self.t_array[::-1])


class _WrappedModel:
    """What `model_wrapper` returns: callable like the reference's model_fn (x, t_continuous) ->
    noise, plus the pieces DPM_Solver's fused path needs."""

    def __init__(self, model, ns, model_type, condition, unconditional_condition, s1, s2):
        self.model, self.ns, self.model_type = model, ns, model_type
        self.condition, self.uncond, self.s1, self.s2 = condition, unconditional_condition, float(s1), float(s2)
        self.use_cfg = not ((self.s1 == 1.0 and self.s2 == 1.0) or unconditional_condition is None)
        self._prepared = None

    def t_input(self, t):
        return (f32(t) - f32(1.0) / f32(self.ns.total_N)) * f32(1000.0)          # :273-281

    def branches(self):
        """Conditioning of each guidance branch in the reference's order (:328-345):
        full-uncond (image zeros, static zeros), image-uncond, cond."""
        if self._prepared is None:
            if not self.use_cfg:
                self._prepared = [self.condition]
            else:
                full = dict(self.uncond)
                full["static_latent"] = torch.zeros_like(full["static_latent"])
                self._prepared = [full, self.uncond, self.condition]
        return self._prepared

    def raw_outputs(self, x, t):
        """Model outputs of all branches stacked on dim 0: [branches*B, ...]."""
        conds = self.branches()
        tin = float(self.t_input(t))
        if hasattr(self.model, "forward_branches"):
            return self.model.forward_branches(x, tin, conds)
        tt = torch.full((x.shape[0],), tin, device=x.device, dtype=torch.float32)
        return torch.cat([self.model(x, tt, **c) for c in conds], 0)

    def __call__(self, x, t_continuous):
        t = float(t_continuous.reshape(-1)[0]) if torch.is_tensor(t_continuous) else float(t_continuous)
        v = self.raw_outputs(x, t).contiguous()
        a, s = self.ns.marginal_alpha(t), self.ns.marginal_std(t)
        # eps through the fused kernel: x0 = (x - s eps)/a  =>  eps = (x - a x0)/s ; do it directly instead
        x0 = torch.empty_like(x)
        ops.dpm_x0(x.contiguous(), v, 3 if self.use_cfg else 1, float(a), float(s), self.s1, self.s2, x0,
                   model_type=1 if self.model_type == "v" else 0)
        eps = torch.empty_like(x)
        # eps = (x - a*x0)/s  ==  (1/s) x - (a/s) x0
        ops.dpm_update(x.contiguous(), x0, None, float(1.0 / s), float(a / s), 0.0, 1, eps)
        return eps


def model_wrapper(model, noise_schedule, model_type="noise", model_kwargs={}, guidance_type="uncond",
                  condition=None, unconditional_condition=None, guidance_scale=1.0, guidance_scale2=1.0,
                  classifier_fn=None, classifier_kwargs={}):
    """reference model/dpmsolver.py:171-351 for model_type in {'noise','v'} and guidance_type in
    {'uncond', 'classifier-free'} (the 3-way CAT4D-style guidance of :328-347)."""
    if model_type not in ("noise", "v"):
        raise NotImplementedError(f"model_type {model_type!r} is not on the inference path")
    if guidance_type == "classifier":
        raise NotImplementedError("classifier guidance is not on the inference path")
    if model_kwargs:
        condition = {**(condition or {}), **model_kwargs}
    if guidance_type == "uncond":
        unconditional_condition = None
        condition = condition or {}
    return _WrappedModel(model, noise_schedule, model_type, condition, unconditional_condition,
                         guidance_scale, guidance_scale2)


class DPM_Solver:
    """reference model/dpmsolver.py:354-1262, algorithm_type='dpmsolver++', methods 'multistep'
    (order 1/2) and 'adaptive' (order 2), skip_type 'time_uniform'."""

    def __init__(self, model_fn, noise_schedule, algorithm_type="dpmsolver++", correcting_x0_fn=None,
                 correcting_xt_fn=None, thresholding_max_val=1.0, dynamic_thresholding_ratio=0.995):
        if algorithm_type != "dpmsolver++":
            raise NotImplementedError("only dpmsolver++ is on the inference path")
        if correcting_x0_fn is not None or correcting_xt_fn is not None:
            raise NotImplementedError("x0 / xt correctors are not on the inference path")
        if not isinstance(model_fn, _WrappedModel):
            raise TypeError("model_fn must come from gvfdiffusion_b200.model.dpmsolver.model_wrapper")
        self.fn, self.ns = model_fn, noise_schedule
        self.nfe = 0

    # x0 prediction (data_prediction_fn :450-459 with noise_pred_fn :284-300 and CFG :328-347 folded in)
    def _x0(self, x, t, out):
        self.nfe += 1
        v = self.fn.raw_outputs(x, t)
        a, s = self.ns.marginal_alpha(t), self.ns.marginal_std(t)
        ops.dpm_x0(x, v, 3 if self.fn.use_cfg else 1, float(a), float(s), self.fn.s1, self.fn.s2, out,
                   model_type=1 if self.fn.model_type == "v" else 0)
        return out

    def _first(self, x, s, t, m_s, out):                              # :564-597
        ns = self.ns
        h = f32(ns.marginal_lambda(t) - ns.marginal_lambda(s))
        cx = f32(ns.marginal_std(t) / ns.marginal_std(s))
        cm = f32(f32(np.exp(ns.marginal_log_mean_coeff(t))) * f32(np.expm1(-h)))
        return ops.dpm_update(x, m_s, None, float(cx), float(cm), 0.0, 1, out)

    def _second_multistep(self, x, m1, m0, t1, t0, t, out):           # :813-848
        ns = self.ns
        l1, l0, lt = ns.marginal_lambda(t1), ns.marginal_lambda(t0), ns.marginal_lambda(t)
        h_0, h = f32(l0 - l1), f32(lt - l0)
        r0 = f32(h_0 / h)
        cx = f32(ns.marginal_std(t) / ns.marginal_std(t0))
        cm = f32(f32(np.exp(ns.marginal_log_mean_coeff(t))) * f32(np.expm1(-h)))
        return ops.dpm_update(x, m0, m1, float(cx), float(cm), float(f32(1.0) / r0), 2, out)

    def _announce_times(self, ts):
        """A fixed-step run knows every model time in advance: let the model precompute what depends on time only."""
        m = self.fn
        inner = getattr(m, "model", None)
        if isinstance(m, _WrappedModel) and hasattr(inner, "precompute_modulation"):
            inner.precompute_modulation([float(m.t_input(t)) for t in ts])

    def sample(self, x, steps=20, t_start=None, t_end=None, order=2, skip_type="time_uniform",
               method="multistep", lower_order_final=True, denoise_to_zero=False, solver_type="dpmsolver",
               atol=0.0078, rtol=0.05, return_intermediate=False):
        if skip_type != "time_uniform" or solver_type != "dpmsolver" or denoise_to_zero or return_intermediate:
            raise NotImplementedError("only time_uniform / dpmsolver / no denoise_to_zero is on the inference path")
        if order not in (1, 2):
            raise NotImplementedError("order must be 1 or 2")
        t_0 = 1.0 / self.ns.total_N if t_end is None else t_end
        t_T = self.ns.T if t_start is None else t_start
        assert t_0 > 0 and t_T > 0
        x = x.to(torch.float32).contiguous().clone()
        with torch.no_grad():
            if method == "adaptive":
                return self._adaptive(x, t_T, t_0, atol, rtol)
            if method != "multistep":
                raise NotImplementedError("method must be 'multistep' or 'adaptive'")
            assert steps >= order
            ts = torch.linspace(t_T, t_0, steps + 1).numpy().astype(f32)          # get_time_steps :491
            self._announce_times(ts[:-1])                                        # the model is evaluated at ts[0 .. steps - 1]
            bufs = [torch.empty_like(x) for _ in range(2)]
            xn = torch.empty_like(x)
            m_prev = [self._x0(x, ts[0], bufs[0])]
            t_prev = [ts[0]]
            for step in range(1, order):                                         # init by lower order :1203-1211
                self._first(x, t_prev[-1], ts[step], m_prev[-1], xn)
                x, xn = xn, x
                t_prev.append(ts[step])
                m_prev.append(self._x0(x, ts[step], bufs[1]))
            for step in range(order, steps + 1):
                t = ts[step]
                so = min(order, steps + 1 - step) if (lower_order_final and steps < 10) else order
                if so == 1:
                    self._first(x, t_prev[-1], t, m_prev[-1], xn)
                else:
                    self._second_multistep(x, m_prev[-2], m_prev[-1], t_prev[-2], t_prev[-1], t, xn)
                x, xn = xn, x
                if order == 2:
                    t_prev[0], m_prev[0], m_prev[1] = t_prev[1], m_prev[1], m_prev[0]
                    t_prev[1] = t
                else:
                    t_prev[0] = t
                if step < steps:
                    self._x0(x, t, m_prev[-1])
            return x

    def _adaptive(self, x, t_T, t_0, atol, rtol, h_init=0.05, theta=0.9, t_err=1e-5):   # :973-1027, order 2
        ns = self.ns
        s = f32(t_T)
        lambda_s, lambda_0 = ns.marginal_lambda(s), ns.marginal_lambda(f32(t_0))
        h = f32(h_init)
        r1 = f32(0.5)
        B = x.shape[0]
        m_s, m_s1 = torch.empty_like(x), torch.empty_like(x)
        x_lower, x_higher, x_s1 = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
        x_prev = x.clone()
        E2 = torch.zeros(B, dtype=torch.float32, device=x.device)
        nfe = 0
        while abs(float(s) - t_0) > t_err:
            t = ns.inverse_lambda(f32(lambda_s + h))
            self._x0(x, s, m_s)
            self._first(x, s, t, m_s, x_lower)
            # singlestep second-order update (:611-660) reusing m_s
            lam_t = ns.marginal_lambda(t)
            hh = f32(lam_t - lambda_s)
            s1 = ns.inverse_lambda(f32(lambda_s + r1 * hh))
            sig_s, sig_s1, sig_t = ns.marginal_std(s), ns.marginal_std(s1), ns.marginal_std(t)
            a_s1, a_t = f32(np.exp(ns.marginal_log_mean_coeff(s1))), f32(np.exp(ns.marginal_log_mean_coeff(t)))
            phi_11, phi_1 = f32(np.expm1(-r1 * hh)), f32(np.expm1(-hh))
            ops.dpm_update(x, m_s, None, float(sig_s1 / sig_s), float(a_s1 * phi_11), 0.0, 1, x_s1)
            self._x0(x_s1, s1, m_s1)
            # x_t = (sig_t/sig_s) x - (a_t phi_1) m_s - (0.5/r1)(a_t phi_1)(m_s1 - m_s)
            ops.dpm_update(x, m_s, m_s1, float(sig_t / sig_s), float(a_t * phi_1), float(-1.0 / r1), 2, x_higher)
            E2.zero_()
            ops.dpm_error_sq(x_higher, x_lower, x_prev, atol, rtol, E2)
            E = float(torch.sqrt(E2 / (x.numel() // B)).max())                   # host sync, as in the reference
            if E <= 1.0:
                x, x_higher = x_higher, x
                s = t
                x_prev, x_lower = x_lower, x_prev
                lambda_s = ns.marginal_lambda(s)
            if not math.isfinite(E):
                raise FloatingPointError(f"adaptive DPM-Solver: error estimate is {E} (non-finite model output)")
            # torch.float_power(E, -1/2) semantics (:1022): E == 0 gives inf, and min() then takes the remaining interval
            with np.errstate(divide="ignore"):
                grow = np.float_power(np.float64(E), -0.5)
            h = f32(min(f32(theta * h * f32(grow)), f32(lambda_0 - lambda_s)))
            nfe += 2
        self.adaptive_nfe = nfe
        return x
