"""`FlowEulerSampler` / `FlowEulerCfgSampler` / `FlowEulerGuidanceIntervalSampler` of the TRELLIS stage (reference
trellis/pipelines/samplers/flow_euler.py:11-199, classifier_free_guidance_mixin.py, guidance_interval_mixin.py): Euler
integration of a flow-matching velocity model from t = 1 to 0 over the rescaled time grid.

Same class names, `sample` / `sample_once` keyword sets and result dict.  Host scalars (the time grid, float64 like the
reference's numpy linspace) + ONE fused launch per step (gvf_flow_euler_step: guidance mix, Euler update and the x_0
prediction -- the reference spends ~10 elementwise torch launches on them); CUDA tensors only.
"""
import numpy as np
import torch

from .... import _lib
from ...._lib import check, current_stream, ptr


class edict(dict):
    __getattr__ = dict.get

    def __setattr__(self, k, v):
        self[k] = v


class FlowEulerSampler:
    def __init__(self, sigma_min: float):
        self.sigma_min = sigma_min

    # guidance hooks (overridden by the two subclasses)
    def _predictions(self, model, x_t, t, cond, **kwargs):
        """-> (v_cond, v_neg or None, cfg_strength)"""
        return self._inference_model(model, x_t, t, cond, **kwargs), None, 0.0

    def _inference_model(self, model, x_t, t, cond=None, **kwargs):
        tt = torch.tensor([1000 * t] * x_t.shape[0], device=x_t.device, dtype=torch.float32)      # SparseTensor: shape[0] = batch
        return model(x_t, tt, cond, **kwargs)

    @torch.no_grad()
    def sample_once(self, model, x_t, t: float, t_prev: float, cond=None, **kwargs):
        """x_t: a dense tensor, or a SparseTensor (the structured-latent stage, trellis_image_to_3d.py:238-248) whose
        feature rows are stepped and re-wrapped on the same coordinates."""
        sparse = hasattr(x_t, "feats") and hasattr(x_t, "replace")
        if not (x_t.feats if sparse else x_t).is_cuda:
            raise RuntimeError("FlowEulerSampler runs on CUDA tensors only (no CPU fallback)")
        v, vn, s = self._predictions(model, x_t, t, cond, **kwargs)
        feats = lambda a: a.feats if sparse else a
        x = feats(x_t).float().contiguous()
        v = feats(v).float().contiguous()
        vn = None if vn is None else feats(vn).float().contiguous()
        x_prev, x_0 = torch.empty_like(x), torch.empty_like(x)
        check(_lib.lib().gvf_flow_euler_step(ptr(x), ptr(v), ptr(vn), x.numel(), float(s), float(t), float(t_prev),
                                             float(self.sigma_min), ptr(x_prev), ptr(x_0), current_stream()),
              "gvf_flow_euler_step")
        if sparse:
            x_prev, x_0 = x_t.replace(x_prev), x_t.replace(x_0)
        return edict({"pred_x_prev": x_prev, "pred_x_0": x_0})

    @torch.no_grad()
    def sample(self, model, noise, cond=None, steps: int = 50, rescale_t: float = 1.0, verbose: bool = True, **kwargs):
        t_seq = np.linspace(1, 0, steps + 1)
        t_seq = rescale_t * t_seq / (1 + (rescale_t - 1) * t_seq)
        ret = edict({"samples": None, "pred_x_t": [], "pred_x_0": []})
        sample = noise
        for i in range(steps):
            out = self.sample_once(model, sample, t_seq[i], t_seq[i + 1], cond, **kwargs)
            sample = out.pred_x_prev
            ret.pred_x_t.append(out.pred_x_prev)
            ret.pred_x_0.append(out.pred_x_0)
        ret.samples = sample
        return ret


class FlowEulerCfgSampler(FlowEulerSampler):
    def _predictions(self, model, x_t, t, cond, neg_cond=None, cfg_strength=3.0, **kwargs):
        return (self._inference_model(model, x_t, t, cond, **kwargs), self._inference_model(model, x_t, t, neg_cond, **kwargs),
                cfg_strength)

    @torch.no_grad()
    def sample(self, model, noise, cond, neg_cond, steps: int = 50, rescale_t: float = 1.0, cfg_strength: float = 3.0,
               verbose: bool = True, **kwargs):
        return super().sample(model, noise, cond, steps, rescale_t, verbose, neg_cond=neg_cond, cfg_strength=cfg_strength,
                              **kwargs)


class FlowEulerGuidanceIntervalSampler(FlowEulerSampler):
    def _predictions(self, model, x_t, t, cond, neg_cond=None, cfg_strength=3.0, cfg_interval=(0.0, 1.0), **kwargs):
        v = self._inference_model(model, x_t, t, cond, **kwargs)
        if cfg_interval[0] <= t <= cfg_interval[1]:
            return v, self._inference_model(model, x_t, t, neg_cond, **kwargs), cfg_strength
        return v, None, 0.0

    @torch.no_grad()
    def sample(self, model, noise, cond, neg_cond, steps: int = 50, rescale_t: float = 1.0, cfg_strength: float = 3.0,
               cfg_interval=(0.0, 1.0), verbose: bool = True, **kwargs):
        return super().sample(model, noise, cond, steps, rescale_t, verbose, neg_cond=neg_cond, cfg_strength=cfg_strength,
                              cfg_interval=cfg_interval, **kwargs)
