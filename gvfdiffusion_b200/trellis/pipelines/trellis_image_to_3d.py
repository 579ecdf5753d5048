"""The sampling / decoding half of `TrellisImageTo3DPipeline` (reference trellis/pipelines/trellis_image_to_3d.py:165-284):
`sample_sparse_structure` -- the flow sampler over `SparseStructureFlowModel`, then the occupancy decoder and
`argwhere(> 0)` -- `sample_slat` -- the flow sampler over `SLatFlowModel` on the active voxels, then de-normalisation --
`decode_slat` to canonical Gaussians, and `run_from_cond` = the body of `run()` after `get_cond`.  Same method names,
arguments and `models` / `*_sampler` / `*_sampler_params` / `slat_normalization` attributes as the reference class.
`models['sparse_structure_decoder']` is `trellis.models.SparseStructureDecoder` (or any callable z_s [B, C, 16, 16, 16] ->
occupancy logits [B, 1, 64, 64, 64]).  Out of scope here: image preprocessing (rembg), the DINOv2 conditioning encoder
(`get_cond`), the mesh / radiance-field decoders."""
import torch

from ... import ops
from ...sparse.basic import SparseTensor
from . import samplers


class TrellisImageTo3DPipeline:
    def __init__(self, models=None, slat_sampler=None, slat_normalization=None, slat_sampler_params=None, device="cuda",
                 sparse_structure_sampler=None, sparse_structure_sampler_params=None):
        self.models = models or {}
        self.sparse_structure_sampler = sparse_structure_sampler
        self.sparse_structure_sampler_params = dict(sparse_structure_sampler_params or {})
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
        ss = args.get("sparse_structure_sampler")
        return TrellisImageTo3DPipeline(models, getattr(samplers, s["name"])(**s["args"]), args["slat_normalization"],
                                        s["params"], device,
                                        getattr(samplers, ss["name"])(**ss["args"]) if ss else None, ss["params"] if ss else None)

    @torch.no_grad()
    def sample_sparse_structure(self, cond: dict, num_samples: int = 1, sampler_params: dict = {}, noise=None) -> torch.Tensor:
        """-> coords int32 [N, 4] = (batch, x, y, z) of the occupied voxels (:165-195)."""
        flow_model = self.models["sparse_structure_flow_model"]
        reso = flow_model.resolution
        if noise is None:
            noise = torch.randn(num_samples, flow_model.in_channels, reso, reso, reso).to(self.device)
        params = {**self.sparse_structure_sampler_params, **sampler_params}
        z_s = self.sparse_structure_sampler.sample(flow_model, noise.to(self.device, torch.float32), **cond, **params,
                                                   verbose=False).samples
        decoder = self.models["sparse_structure_decoder"]
        return torch.argwhere(decoder(z_s) > 0)[:, [0, 2, 3, 4]].int()

    @torch.no_grad()
    def run_from_cond(self, cond: dict, num_samples: int = 1, sparse_structure_sampler_params: dict = {},
                      slat_sampler_params: dict = {}, formats=("gaussian",)) -> dict:
        """`run()` after `get_cond` (:279-284)."""
        coords = self.sample_sparse_structure(cond, num_samples, sparse_structure_sampler_params)
        slat = self.sample_slat(cond, coords, slat_sampler_params)
        return self.decode_slat(slat, formats)

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
