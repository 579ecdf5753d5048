"""`SparseConv3d`: submanifold sparse 3-D convolution on the device, mirroring the reference's
sparse/conv/conv_spconv.py:6-46 (`spconv.SubMConv3d` when stride == 1 and padding is None -- the only form its
callers use: trellis/models/structured_latent_flow.py:34-35, structured_latent_vae/decoder_mesh.py:43-52).

    out[i] = bias + sum_k W[:, k, :] x[row of the voxel at coords[i] + dilation (k - ks // 2)]      (absent = 0)

State-dict keys and layout follow spconv 2.x: `conv.weight` [Cout, kx, ky, kz, Cin], `conv.bias` [Cout].
Execution: neighbour map (cached per coordinate set under `indice_key`, like spconv's indice_dict) -> ONE tcgen05 GEMM
whose TMA producer gathers the neighbour rows itself (`cp.async.bulk.tensor ... tile::gather4`, absent voxels zero-filled),
bias in the epilogue; fp32 features or Cin % 64 != 0 go through the im2col gather + plain GEMM instead.
Strided / padded SparseConv3d (spatial resampling) is not part of the path and raises."""
import torch

from .. import ops
from .basic import SparseTensor

F16, F32 = torch.float16, torch.float32


class SparseConv3d:
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, dilation=1, padding=None, bias=True,
                 indice_key=None, device="cuda"):
        if stride != 1 or padding is not None:
            raise NotImplementedError("only the submanifold form (stride 1, padding None) is on the path")
        if in_channels % 8 or out_channels % 8:
            raise ValueError("channel counts must be multiples of 8 (16 B rows)")
        self.in_channels, self.out_channels, self.kernel_size, self.dilation = in_channels, out_channels, kernel_size, dilation
        self.indice_key = indice_key
        self.fused_gather = True            # False: im2col + GEMM (the only path for fp32 features / Cin % 64 != 0)
        self.device = torch.device(device)
        self.weight = torch.zeros(out_channels, kernel_size ** 3 * in_channels, dtype=F16, device=self.device)
        self.bias = torch.zeros(out_channels, dtype=F32, device=self.device) if bias else None

    def load_state_dict(self, sd, prefix=""):
        w = sd[prefix + "conv.weight"]
        ks = self.kernel_size
        assert tuple(w.shape) == (self.out_channels, ks, ks, ks, self.in_channels), tuple(w.shape)
        self.weight = w.detach().to(self.device, F16).reshape(self.out_channels, -1).contiguous()
        if self.bias is not None:
            # spconv adds the bias to the fp16 result under autocast; the GEMM epilogue adds it in fp32 before
            # the single fp16 rounding (one rounding fewer, covered by the parity tolerance)
            self.bias = sd[prefix + "conv.bias"].detach().to(self.device, F32).contiguous()
        return self

    def neighbor_map(self, x: SparseTensor, grid_size=None):
        key = f"submconv_{self.kernel_size}_{self.dilation}_{self.indice_key}"
        nbr = x.get_spatial_cache(key)
        if nbr is None or nbr.shape[0] != x.coords.shape[0]:
            if grid_size is None:
                grid_size = int(x.coords[:, 1:].max()) + 1          # one host sync per coordinate set, then cached
            status = torch.zeros(1, dtype=torch.int32, device=x.coords.device)
            nbr = ops.sparse_neighbor_map(x.coords, x.shape[0], grid_size, self.kernel_size, self.dilation, status=status)
            st = int(status.item())
            if st & 1:
                raise ValueError("SparseConv3d: coordinates outside [0, grid_size) or batch index out of range")
            if st & 2:
                raise ValueError("SparseConv3d: duplicate voxel coordinates")
            x.register_spatial_cache(key, nbr)
        return nbr

    def forward(self, x: SparseTensor, grid_size=None, epilogue=ops.EPI_F16, out=None) -> SparseTensor:
        if not x.feats.is_cuda:
            raise RuntimeError("SparseConv3d runs on the device only (no CPU fallback)")
        nbr = self.neighbor_map(x, grid_size)
        if (self.fused_gather and self.in_channels % 64 == 0 and x.feats.dtype == F16 and x.feats.stride(1) == 1
                and epilogue in (ops.EPI_F16, ops.EPI_F32)):
            # one kernel: the GEMM's TMA producer gathers the neighbour rows (tile::gather4), no im2col operand
            y = ops.sparse_conv_gemm(x.feats, nbr, self.weight, self.bias, out_f32=epilogue == ops.EPI_F32, out=out)
            return x.replace(y)
        a = ops.sparse_im2col(x.feats if x.feats.dtype in (F16, F32) else x.feats.float(), nbr)
        y = ops.gemm(a, self.weight, self.bias, epilogue, out=out)
        return x.replace(y)

    __call__ = forward
