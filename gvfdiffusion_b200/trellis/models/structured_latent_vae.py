"""`SLatGaussianDecoder`: structured latents -> canonical Gaussians, the last model of the TRELLIS stage in front of the
GVF path (reference trellis/models/structured_latent_vae/decoder_gs.py:10-122 over
SparseTransformerBase, base.py:36-117; called by trellis/pipelines/trellis_image_to_3d.py:199-221 `decode_slat`).

    input_layer + APE(coords) -> swin SparseTransformerBlocks -> layer_norm -> out_layer -> to_representation

It is the architecture of GVF's own static-VAE decoder (model/sparse_voxel_diffusion/sparse_transformer_vae.py:178-188 with
norm_output) under other state-dict names, so it runs on the same device engine -- `sparse.transformer.SparseTransformerVAE`
(packed window attention with the gather fused, tcgen05 GEMMs with fused epilogues, native launch driver) -- and the same
`gvf_to_representation` kernel; this class keeps the reference's constructor arguments, key names (`input_layer`,
`blocks.{i}.attn.to_qkv / to_out`, `blocks.{i}.mlp.mlp.{0,2}`, `out_layer`, buffer `offset_perturbation`) and return type
(one Gaussian model per batch entry).  Inference only, CUDA only."""
import torch

from ...model.sparse_voxel_diffusion.sparse_vae import SparseVAE
from ...sparse.basic import SparseTensor
from ...sparse.transformer import SparseTransformerVAE


class SLatGaussianDecoder:
    def __init__(self, resolution, model_channels, latent_channels, num_blocks, num_heads=None, num_head_channels=64,
                 mlp_ratio=4, attn_mode="swin", window_size=8, pe_mode="ape", use_fp16=False, use_checkpoint=False,
                 qk_rms_norm=False, representation_config=None, device="cuda"):
        if attn_mode != "swin" or pe_mode != "ape" or qk_rms_norm or mlp_ratio != 4:
            raise NotImplementedError("the shipped decoder (slat_dec_gs_swin8_*) is swin / ape / no q-k norm / mlp_ratio 4")
        self.resolution, self.model_channels, self.latent_channels = resolution, model_channels, latent_channels
        self.num_blocks, self.window_size, self.use_fp16 = num_blocks, window_size, use_fp16
        self.num_heads = num_heads or model_channels // num_head_channels
        self.rep_config = dict(representation_config)
        self.device = torch.device(device)
        # decoder_gs.py:101-110 is the static VAE's MipGS branch with the soft in-voxel offset (sparse_vae.py:165-180)
        self._rep = SparseVAE(resolution=resolution, representation_config={"MipGS": dict(self.rep_config, reg_mode="soft_invoxel")},
                              device=device)
        self.out_channels = self._rep.out_channels
        self.layout = {k: dict(v, shape=shape) for (k, v), shape in zip(
            self._rep.layouts["MipGS"].items(),
            [(self.rep_config["num_gaussians"], 3), (self.rep_config["num_gaussians"], 1, 3), (self.rep_config["num_gaussians"], 3),
             (self.rep_config["num_gaussians"], 4), (self.rep_config["num_gaussians"], 1)])}
        self.offset_perturbation = self._rep.perturbation.get("MipGS")
        self.engine = None

    def load_state_dict(self, sd, strict=True):
        ren = {}
        for k, v in sd.items():
            if k.startswith("blocks."):
                ren["decoder." + k[len("blocks."):]] = v
            elif k.startswith("input_layer."):
                ren["from_latent." + k[len("input_layer."):]] = v
            elif k.startswith("out_layer."):
                ren[k] = v
        if strict and tuple(sd["out_layer.weight"].shape) != (self.out_channels, self.model_channels):
            raise ValueError("out_layer does not match representation_config")
        self.engine = SparseTransformerVAE(ren, self.num_blocks, self.num_heads, self.window_size, use_fp16=self.use_fp16,
                                           norm_output=True, device=self.device)
        return self

    @torch.no_grad()
    def decode_rows(self, x: SparseTensor):
        """-> SparseTensor of the out_layer rows [N, 14 * num_gaussians] (decoder_gs.py:117-121)."""
        if self.engine is None:
            raise RuntimeError("load_state_dict first")
        if not x.feats.is_cuda:
            raise RuntimeError("SLatGaussianDecoder runs on CUDA tensors only (no CPU fallback)")
        return x.replace(self.engine.decode(x.feats.float(), x.coords))

    def to_representation(self, x: SparseTensor):
        """decoder_gs.py:81-115 -> [Gaussian model per batch entry]."""
        return self._rep.to_representation(x)["MipGS"]

    def forward(self, x: SparseTensor):
        return self.to_representation(self.decode_rows(x))

    __call__ = forward
