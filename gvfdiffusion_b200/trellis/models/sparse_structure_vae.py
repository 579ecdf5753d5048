"""`SparseStructureDecoder`: occupancy latent -> occupancy logits, the dense Conv3d ResNet between the two flow models of
the TRELLIS stage (reference trellis/models/sparse_structure_vae.py:209-306; `torch.argwhere(decoder(z_s) > 0)` gives the
active voxels, trellis_image_to_3d.py:192-193).  Same constructor arguments and state-dict keys as the reference class
(norm_type 'layer', upsampling by convolution + pixel shuffle: the shipped ss_dec_conv3d_16l8 configuration).

A dense 3 x 3 x 3 convolution with zero padding IS the submanifold convolution on a fully active grid, so the decoder runs on
the voxel-side operators of this library with rows = voxels in (batch, x, y, z) order and channels last: neighbour map of the
full grid (gvf_sparse_neighbor_map, once per grid size) -> fp16 im2col operand -> tcgen05 GEMM with the bias (and the
residual of a ResBlock's skip path) in its epilogue; ChannelLayerNorm32 + SiLU is one gvf_ln_mod_act_f16 pass that writes the
next convolution's operand; pixel_shuffle_3d is an index permutation of the rows (one torch copy).  Inference only, CUDA
only."""
import torch

from ... import ops
from ...sparse.basic import SparseTensor
from ...sparse.conv import SparseConv3d

F16, F32 = torch.float16, torch.float32


def _conv_weight(w):
    """nn.Conv3d weight [Cout, Cin, kx, ky, kz] -> [Cout, kx * ky * kz * Cin] fp16 rows (tap-major, channel-minor)."""
    return w.detach().permute(0, 2, 3, 4, 1).reshape(w.shape[0], -1)


class SparseStructureDecoder:
    def __init__(self, out_channels, latent_channels, num_res_blocks, channels, num_res_blocks_middle=2, norm_type="layer",
                 use_fp16=False, device="cuda"):
        if norm_type != "layer":
            raise NotImplementedError("the shipped decoder uses norm_type 'layer' (ChannelLayerNorm32)")
        self.out_channels, self.latent_channels, self.num_res_blocks = out_channels, latent_channels, num_res_blocks
        self.channels, self.num_res_blocks_middle = list(channels), num_res_blocks_middle
        self.device = torch.device(device)
        self._grids = {}                          # (B, R) -> SparseTensor of the full grid (neighbour-map cache owner)
        self._loaded = False

    def load_state_dict(self, sd, strict=True):
        dev = self.device
        h = lambda t: t.detach().to(dev, F16).contiguous()
        f = lambda t: t.detach().to(dev, F32).contiguous()

        def conv(name, pad_out=0):
            w, b = _conv_weight(sd[name + ".weight"]), sd[name + ".bias"].detach().float()
            if pad_out:
                w = torch.cat([w, torch.zeros(pad_out, w.shape[1], dtype=w.dtype)], 0)
                b = torch.cat([b, torch.zeros(pad_out)], 0)
            return h(w), f(b)

        def res(p):
            if p + "skip_connection.weight" in sd:
                raise NotImplementedError("ResBlock3d with a channel change is not used by the decoder")
            return dict(n1=(f(sd[p + "norm1.weight"]), f(sd[p + "norm1.bias"])), n2=(f(sd[p + "norm2.weight"]), f(sd[p + "norm2.bias"])),
                        c1=conv(p + "conv1"), c2=conv(p + "conv2"))
        if any(c % 8 for c in self.channels) or self.latent_channels % 8:
            raise ValueError("channel counts must be multiples of 8 (16 B rows)")
        self.input = conv("input_layer")
        self.middle = [res(f"middle_block.{i}.") for i in range(self.num_res_blocks_middle)]
        self.stages, bi = [], 0
        for lvl, ch in enumerate(self.channels):
            blocks = []
            for _ in range(self.num_res_blocks):
                blocks.append(res(f"blocks.{bi}."))
                bi += 1
            up = None
            if lvl < len(self.channels) - 1:
                up = conv(f"blocks.{bi}.conv")
                bi += 1
            self.stages.append((blocks, up))
        self.out_norm = (f(sd["out_layer.0.weight"]), f(sd["out_layer.0.bias"]))
        self.out_conv = conv("out_layer.2", (-self.out_channels) % 8)
        self._mapper = SparseConv3d(8, 8, 3, device=dev)           # only its neighbour-map builder is used
        self._loaded = True
        return self

    def _grid(self, B, R):
        g = self._grids.get((B, R))
        if g is None:
            ax = torch.arange(R, device=self.device, dtype=torch.int32)
            b, x, y, z = torch.meshgrid(torch.arange(B, device=self.device, dtype=torch.int32), ax, ax, ax, indexing="ij")
            coords = torch.stack([b, x, y, z], -1).reshape(-1, 4).contiguous()
            layout = [slice(i * R ** 3, (i + 1) * R ** 3) for i in range(B)]
            g = SparseTensor(torch.empty((coords.shape[0], 1), device=self.device), coords, torch.Size([B, 1]), layout)
            g.R = R
            if len(self._grids) >= 8:
                self._grids.clear()
            self._grids[(B, R)] = g
        return g

    def _conv(self, grid, a, wb, epilogue=ops.EPI_F16, out=None):
        """3 x 3 x 3 convolution with zero padding on the full grid: im2col + GEMM (rows = voxels, channels last)."""
        nbr = self._mapper.neighbor_map(grid, grid_size=grid.R)                                   # cached on the grid
        return ops.gemm(ops.sparse_im2col(a, nbr), wb[0], wb[1], epilogue, out=out)

    def _res(self, grid, blk, x):
        """ResBlock3d.forward (:37-45): x fp16 [rows, C]."""
        a = ops.ln_mod_act(x, eps=1e-5, w=blk["n1"][0], b=blk["n1"][1], act=1)
        h = self._conv(grid, a, blk["c1"])
        a = ops.ln_mod_act(h, eps=1e-5, w=blk["n2"][0], b=blk["n2"][1], act=1)
        return self._conv(grid, a, blk["c2"], ops.EPI_RESID_F16, out=x.clone())

    @torch.no_grad()
    def forward(self, z):
        if not self._loaded:
            raise RuntimeError("load_state_dict first")
        if not z.is_cuda:
            raise RuntimeError("SparseStructureDecoder runs on CUDA tensors only (no CPU fallback)")
        B, C, R = z.shape[0], z.shape[1], z.shape[2]
        assert C == self.latent_channels and tuple(z.shape[2:]) == (R, R, R)
        grid = self._grid(B, R)
        rows = z.to(F32).permute(0, 2, 3, 4, 1).reshape(B * R ** 3, C).contiguous()           # channels last
        h = self._conv(grid, rows, self.input)
        for blk in self.middle:
            h = self._res(grid, blk, h)
        for blocks, up in self.stages:
            for blk in blocks:
                h = self._res(grid, blk, h)
            if up is not None:                                                                 # UpsampleBlock3d (:97-101)
                u = self._conv(grid, h, up)                                                    # [rows, 8 c]: c-major, (sx, sy, sz)-minor
                c = u.shape[1] // 8
                u = u.view(B, R, R, R, c, 2, 2, 2).permute(0, 1, 5, 2, 6, 3, 7, 4)             # pixel_shuffle_3d on rows
                R *= 2
                h = u.reshape(B * R ** 3, c).contiguous()
                grid = self._grid(B, R)
        a = ops.ln_mod_act(h, eps=1e-5, w=self.out_norm[0], b=self.out_norm[1], act=1)
        y = torch.empty((B * R ** 3, self.out_conv[0].shape[0]), dtype=F32, device=self.device)
        self._conv(grid, a, self.out_conv, ops.EPI_F32, out=y)
        return y[:, :self.out_channels].reshape(B, R, R, R, self.out_channels).permute(0, 4, 1, 2, 3).contiguous()

    __call__ = forward
