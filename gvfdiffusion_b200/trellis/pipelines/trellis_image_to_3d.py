"""The structured-latent half of `TrellisImageTo3DPipeline` (reference trellis/pipelines/trellis_image_to_3d.py:197-256):
`sample_slat` -- the flow sampler over `SLatFlowModel` on the active voxels, then de-normalisation -- and `decode_slat` to
canonical Gaussians.  Same method names, arguments and `models` / `slat_sampler` / `slat_sampler_params` /
`slat_normalization` attributes as the reference class.  Out of scope here: image preprocessing (rembg), the DINOv2
conditioning encoder, the dense sparse-structure stage that produces `coords`, the mesh / radiance-field decoders."""
import torch

from ... import ops
from ...sparse.basic import SparseTensor
from . import samplers


class TrellisImageTo3DPipeline:
    def __init__(self, models=None, slat_sampler=None, slat_normalization=None, slat_sampler_params=None, device="cuda"):
        self.models = models or {}
        self.slat_sampler = slat_sampler
        self.slat_sampler_params = dict(slat_sampler_params or {})
        self.slat_normalization = slat_normalization
        self.device = torch.device(device)
        self._norm = None

    @staticmethod
    def from_args(args, models, device="cuda"):
        """`args` = the `args` block of the reference's pipeline.json (:55-67): slat_sampler {name, args, params},
        slat_normalization {mean, std}."""
        s = args["slat_sampler"]
        return TrellisImageTo3DPipeline(models, getattr(samplers, s["name"])(**s["args"]), args["slat_normalization"],
                                        s["params"], device)

    @torch.no_grad()
    def sample_slat(self, cond: dict, coords: torch.Tensor, sampler_params: dict = {}, noise=None) -> SparseTensor:
        """cond: {'cond': [B, L, C], 'neg_cond': ...}; coords int [N, 4] -> the structured latent (:223-256)."""
        flow_model = self.models["slat_flow_model"]
        if noise is None:
            noise = torch.randn(coords.shape[0], flow_model.in_channels).to(self.device)
        noise = SparseTensor(noise.to(self.device, torch.float32), coords.to(self.device))
        params = {**self.slat_sampler_params, **sampler_params}
        slat = self.slat_sampler.sample(flow_model, noise, **cond, **params, verbose=False).samples
        if self._norm is None:
            self._norm = (torch.tensor(self.slat_normalization["std"], dtype=torch.float32, device=self.device),
                          torch.tensor(self.slat_normalization["mean"], dtype=torch.float32, device=self.device))
        return slat.replace(ops.affine_lastdim(slat.feats.contiguous(), self._norm[0], self._norm[1]))      # slat * std + mean

    @torch.no_grad()
    def decode_slat(self, slat: SparseTensor, formats=("gaussian",)) -> dict:
        ret = {}
        for f in formats:
            if f != "gaussian":
                raise NotImplementedError(f"format {f!r}: only the Gaussian decoder is on the GVF path")
            ret["gaussian"] = self.models["slat_decoder_gs"](slat)
        return ret
