"""`SpacedDiffusion` / `space_timesteps` / `create_gaussian_diffusion` of the reference (`model/respace.py:8-170`,
`utils/script_util.py:7-61`): a diffusion process restricted to a subset of the base timesteps; the wrapped model
sees the ORIGINAL timestep numbers (optionally rescaled to [0, 1000))."""
import numpy as np
import torch

from . import gaussian_diffusion as gd


def space_timesteps(num_timesteps, section_counts):
    """Timesteps kept when each equal section of the base process is strided down to `section_counts[i]` steps
    ("ddimN": the single integer stride giving exactly N steps; "fast27": the reference's hand-tuned split)."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            want = int(section_counts[4:])
            for stride in range(1, num_timesteps):
                if len(range(0, num_timesteps, stride)) == want:
                    return set(range(0, num_timesteps, stride))
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        if section_counts == "fast27":
            steps = space_timesteps(num_timesteps, "10,10,3,2,2")
            steps.remove(num_timesteps - 1)
            steps.add(num_timesteps - 3)
            return steps
        section_counts = [int(v) for v in section_counts.split(",")]
    base, extra = divmod(num_timesteps, len(section_counts))
    kept, start = [], 0
    for i, count in enumerate(section_counts):
        size = base + (1 if i < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        stride = 1 if count <= 1 else (size - 1) / (count - 1)
        pos = 0.0                                   # accumulated like the reference (:97-101): same roundings
        for _ in range(count):
            kept.append(start + round(pos))
            pos += stride
        start += size
    return set(kept)


class _WrappedModel:
    def __init__(self, model, timestep_map, rescale_timesteps, original_num_steps):
        self.model, self.timestep_map = model, timestep_map
        self.rescale_timesteps, self.original_num_steps = rescale_timesteps, original_num_steps

    def __call__(self, x, ts, **kwargs):
        new_ts = torch.tensor(self.timestep_map, device=ts.device, dtype=ts.dtype)[ts]
        if self.rescale_timesteps:
            new_ts = new_ts.float() * (1000.0 / self.original_num_steps)
        return self.model(x, new_ts, **kwargs)

    def parameters(self):
        return self.model.parameters()


class SpacedDiffusion(gd.GaussianDiffusion):
    def __init__(self, use_timesteps, **kwargs):
        self.use_timesteps = set(use_timesteps)
        self.original_num_steps = len(kwargs["betas"])
        base = gd.GaussianDiffusion(**kwargs)
        self.timestep_map, new_betas, last = [], [], 1.0
        for i, ac in enumerate(base.alphas_cumprod):
            if i in self.use_timesteps:
                new_betas.append(1 - ac / last)
                last = ac
                self.timestep_map.append(i)
        kwargs["betas"] = np.array(new_betas)
        super().__init__(**kwargs)

    def _wrap_model(self, model):
        if isinstance(model, _WrappedModel):
            return model
        return _WrappedModel(model, self.timestep_map, self.rescale_timesteps, self.original_num_steps)

    def p_mean_variance(self, model, *args, **kwargs):
        return super().p_mean_variance(self._wrap_model(model), *args, **kwargs)

    def _scale_timesteps(self, t):
        return t            # the wrapped model rescales


def create_gaussian_diffusion(*, steps=1000, learn_sigma=False, sigma_small=False, noise_schedule="linear", use_kl=False,
                              predict_type="eps", predict_xstart=False, rescale_timesteps=False,
                              rescale_learned_sigmas=False, timestep_respacing="", beta_start=0.0001, beta_end=0.02,
                              min_snr=False):
    betas = gd.get_named_beta_schedule(noise_schedule, steps, beta_start, beta_end)
    loss = gd.LossType.RESCALED_KL if use_kl else gd.LossType.RESCALED_MSE if rescale_learned_sigmas else gd.LossType.MSE
    mean = {"eps": gd.ModelMeanType.EPSILON, "xstart": gd.ModelMeanType.START_X, "v": gd.ModelMeanType.V}.get(predict_type)
    if mean is None:
        raise ValueError(f"Unknown predict_type for diffusion model: {predict_type}")
    var = gd.ModelVarType.LEARNED_RANGE if learn_sigma else (
        gd.ModelVarType.FIXED_SMALL if sigma_small else gd.ModelVarType.FIXED_LARGE)
    return SpacedDiffusion(use_timesteps=space_timesteps(steps, timestep_respacing or [steps]), betas=betas,
                           model_mean_type=mean, model_var_type=var, loss_type=loss,
                           rescale_timesteps=rescale_timesteps, min_snr=min_snr)
