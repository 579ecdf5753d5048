"""Ancestral sampler of the reference (`model/gaussian_diffusion.py:279-560`, BASELINE.json configs[0]:
`gaussian_diffusion.p_sample`, one step on a 4x8^3 dense latent, CPU fp32 -- plumbing, no extension involved on
either side).  Same class / method names and keyword sets as the reference so that `create_gaussian_diffusion`
callers (`utils/script_util.py:7-61`) switch over unchanged.

Host design: every schedule table is a float64 numpy array built once; a step gathers its per-batch scalars on
the host, casts them to float32 exactly where the reference's `_extract_into_tensor` does (`arr[t].float()`), and
applies them with device-agnostic torch expressions (the tensor stays where the caller put it; the DiT behind
`model` is the sm_100a engine when it is a `gvfdiffusion_b200.model.dit.DiT`).
"""
import enum
import math

import numpy as np
import torch


class ModelMeanType(enum.Enum):
    PREVIOUS_X = enum.auto()
    START_X = enum.auto()
    EPSILON = enum.auto()
    V = enum.auto()


class ModelVarType(enum.Enum):
    LEARNED = enum.auto()
    FIXED_SMALL = enum.auto()
    FIXED_LARGE = enum.auto()
    LEARNED_RANGE = enum.auto()


class LossType(enum.Enum):
    MSE = enum.auto()
    RESCALED_MSE = enum.auto()
    KL = enum.auto()
    RESCALED_KL = enum.auto()


def betas_for_alpha_bar(n, alpha_bar, max_beta=0.999):
    """beta_i = min(1 - abar((i+1)/n) / abar(i/n), max_beta) (reference :75-89), float64."""
    return np.array([min(1.0 - alpha_bar((i + 1) / n) / alpha_bar(i / n), max_beta) for i in range(n)])


def get_named_beta_schedule(schedule_name, num_diffusion_timesteps, beta_start=0.0001, beta_end=0.02):
    n = num_diffusion_timesteps
    if schedule_name == "linear":
        k = 1000 / n
        return np.linspace(k * beta_start, k * beta_end, n, dtype=np.float64)
    if schedule_name == "cosine":
        return betas_for_alpha_bar(n, lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2)
    raise NotImplementedError(f"unknown beta schedule: {schedule_name}")


def _extract_into_tensor(arr, timesteps, broadcast_shape):
    """float64 table -> float32 per-batch column broadcast to `broadcast_shape` (reference :787-800)."""
    res = torch.from_numpy(np.asarray(arr))[timesteps.cpu()].float().to(timesteps.device)
    return res.reshape(-1, *([1] * (len(broadcast_shape) - 1))).expand(broadcast_shape)


class GaussianDiffusion:
    def __init__(self, *, betas, model_mean_type, model_var_type, loss_type=LossType.MSE, rescale_timesteps=False,
                 min_snr=False):
        self.model_mean_type, self.model_var_type, self.loss_type = model_mean_type, model_var_type, loss_type
        self.rescale_timesteps, self.min_snr = rescale_timesteps, min_snr
        b = np.array(betas, dtype=np.float64)
        if b.ndim != 1 or not ((b > 0).all() and (b <= 1).all()):
            raise ValueError("betas must be a 1-D array in (0, 1]")
        self.betas, self.num_timesteps = b, int(b.shape[0])
        a = 1.0 - b
        ac = np.cumprod(a)
        self.alphas_cumprod = ac
        self.alphas_cumprod_prev = np.append(1.0, ac[:-1])
        self.alphas_cumprod_next = np.append(ac[1:], 0.0)
        self.sqrt_alphas_cumprod = np.sqrt(ac)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - ac)
        self.log_one_minus_alphas_cumprod = np.log(1.0 - ac)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / ac)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / ac - 1)
        pv = b * (1.0 - self.alphas_cumprod_prev) / (1.0 - ac)
        self.posterior_variance = pv
        self.posterior_log_variance_clipped = np.log(np.append(pv[1], pv[1:]))
        self.posterior_mean_coef1 = b * np.sqrt(self.alphas_cumprod_prev) / (1.0 - ac)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(a) / (1.0 - ac)

    # ---- forward process -------------------------------------------------------------------------------------
    def q_sample(self, x_start, t, noise=None):
        noise = torch.randn_like(x_start) if noise is None else noise
        return (_extract_into_tensor(self.sqrt_alphas_cumprod, t, x_start.shape) * x_start
                + _extract_into_tensor(self.sqrt_one_minus_alphas_cumprod, t, x_start.shape) * noise)

    def get_v(self, x, noise, t):
        return (_extract_into_tensor(self.sqrt_alphas_cumprod, t, x.shape) * noise
                - _extract_into_tensor(self.sqrt_one_minus_alphas_cumprod, t, x.shape) * x)

    def q_posterior_mean_variance(self, x_start, x_t, t):
        mean = (_extract_into_tensor(self.posterior_mean_coef1, t, x_t.shape) * x_start
                + _extract_into_tensor(self.posterior_mean_coef2, t, x_t.shape) * x_t)
        return (mean, _extract_into_tensor(self.posterior_variance, t, x_t.shape),
                _extract_into_tensor(self.posterior_log_variance_clipped, t, x_t.shape))

    # ---- reverse process -------------------------------------------------------------------------------------
    @staticmethod
    def dynamic_thresholding(x, p=0.995, c=1.7):
        """Per-sample clip to the p-quantile of |x| (reference :197-215; `c` is unused there too)."""
        flat = x.reshape(x.shape[0], -1)
        s = torch.quantile(flat.abs(), p, dim=-1)
        return torch.clip(flat.T, -s, s).T.reshape(x.shape)

    def _scale_timesteps(self, t):
        return t.float() * (1000.0 / self.num_timesteps) if self.rescale_timesteps else t

    def _predict_xstart_from_eps(self, x_t, t, eps):
        return (_extract_into_tensor(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t
                - _extract_into_tensor(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape) * eps)

    def _predict_start_from_z_and_v(self, x_t, t, v):
        return (_extract_into_tensor(self.sqrt_alphas_cumprod, t, x_t.shape) * x_t
                - _extract_into_tensor(self.sqrt_one_minus_alphas_cumprod, t, x_t.shape) * v)

    def _predict_eps_from_z_and_v(self, x_t, t, v):
        return (_extract_into_tensor(self.sqrt_alphas_cumprod, t, x_t.shape) * v
                + _extract_into_tensor(self.sqrt_one_minus_alphas_cumprod, t, x_t.shape) * x_t)

    def p_mean_variance(self, model, x, t, clip_denoised=True, dynamic_thresholding_p=0.99, dynamic_thresholding_c=1.7,
                        denoised_fn=None, model_kwargs=None):
        B = x.shape[0]
        if tuple(t.shape) != (B,):
            raise ValueError("t must have shape (batch,)")
        out = model(x, self._scale_timesteps(t), **(model_kwargs or {}))
        if self.model_var_type in (ModelVarType.LEARNED, ModelVarType.LEARNED_RANGE):
            raise NotImplementedError("learned variances are not used by the shipped configs (learn_sigma: false)")
        if self.model_var_type == ModelVarType.FIXED_LARGE:
            var = np.append(self.posterior_variance[1], self.betas[1:])
            logvar = np.log(var)
        else:
            var, logvar = self.posterior_variance, self.posterior_log_variance_clipped
        variance = _extract_into_tensor(var, t, x.shape)
        log_variance = _extract_into_tensor(logvar, t, x.shape)

        def process(x0):
            if denoised_fn is not None:
                x0 = denoised_fn(x0)
            if clip_denoised:
                x0 = self.dynamic_thresholding(x0, p=dynamic_thresholding_p, c=dynamic_thresholding_c)
            return x0

        if self.model_mean_type == ModelMeanType.V:
            pred_xstart = process(self._predict_start_from_z_and_v(x, t, out))
        elif self.model_mean_type == ModelMeanType.EPSILON:
            pred_xstart = process(self._predict_xstart_from_eps(x, t, out))
        elif self.model_mean_type == ModelMeanType.START_X:
            pred_xstart = process(out)
        else:
            raise NotImplementedError(self.model_mean_type)
        mean, _, _ = self.q_posterior_mean_variance(pred_xstart, x, t)
        return {"mean": mean, "variance": variance, "log_variance": log_variance, "pred_xstart": pred_xstart}

    def p_sample(self, model, x, t, clip_denoised=True, dynamic_thresholding_p=0.99, dynamic_thresholding_c=1.7,
                 denoised_fn=None, model_kwargs=None, inpainting_mask=None):
        out = self.p_mean_variance(model, x, t, clip_denoised=clip_denoised, dynamic_thresholding_p=dynamic_thresholding_p,
                                   dynamic_thresholding_c=dynamic_thresholding_c, denoised_fn=denoised_fn,
                                   model_kwargs=model_kwargs)
        noise = torch.randn_like(x)
        nonzero = (t != 0).float().view(-1, *([1] * (x.dim() - 1)))
        sample = out["mean"] + nonzero * torch.exp(0.5 * out["log_variance"]) * noise
        if inpainting_mask is not None:
            sample = (1 - inpainting_mask) * x + inpainting_mask * sample
        return {"sample": sample, "pred_xstart": out["pred_xstart"]}

    def p_sample_loop_progressive(self, model, shape, inpainting_mask=None, noise=None, clip_denoised=True,
                                  dynamic_thresholding_p=0.99, dynamic_thresholding_c=1.7, denoised_fn=None,
                                  model_kwargs=None, device=None, progress=False):
        if device is None:
            device = next(model.parameters()).device
        img = noise if noise is not None else torch.randn(*shape, device=device)
        for i in reversed(range(self.num_timesteps)):
            t = torch.tensor([i] * shape[0], device=device)
            with torch.no_grad():
                out = self.p_sample(model, img, t, clip_denoised=clip_denoised,
                                    dynamic_thresholding_p=dynamic_thresholding_p,
                                    dynamic_thresholding_c=dynamic_thresholding_c, denoised_fn=denoised_fn,
                                    inpainting_mask=inpainting_mask, model_kwargs=model_kwargs)
            yield out
            img = out["sample"]

    def p_sample_loop(self, model, shape, noise=None, clip_denoised=True, dynamic_thresholding_p=0.99,
                      dynamic_thresholding_c=1.7, inpainting_mask=None, denoised_fn=None, model_kwargs=None,
                      device=None, progress=False, sample_fn=None):
        final = None
        for k, sample in enumerate(self.p_sample_loop_progressive(
                model, shape, noise=noise, clip_denoised=clip_denoised, dynamic_thresholding_p=dynamic_thresholding_p,
                dynamic_thresholding_c=dynamic_thresholding_c, denoised_fn=denoised_fn,
                inpainting_mask=inpainting_mask, model_kwargs=model_kwargs, device=device, progress=progress)):
            final = sample_fn(k, sample) if sample_fn is not None else sample
        return final["sample"]
