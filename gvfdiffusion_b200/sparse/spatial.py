"""`SparseDownsample` / `SparseUpsample` of the TRELLIS stage (reference trellis/modules/sparse/spatial.py:13-80).

Downsample(f): coordinates // f, cells in lexicographic (batch, x, y, z) order -- the order of the reference's
`code.unique()` -- and every cell's features = (sum of its fine rows) / (count + 1): the reference pools with
`torch.scatter_reduce(zeros, ..., reduce='mean')`, whose default `include_self=True` counts the zero it starts from
(pinned by tests/golden/slat_flow_tiny.pt).  The integer bookkeeping (cell code, unique, the fine rows grouped by cell) is
torch index work done ONCE per coordinate set and cached on the SparseTensor like the reference's `upsample_*` entries;
the feature pass is gvf_sparse_pool_mean_f16 (deterministic fp32 sums, one fp16 rounding; the reference accumulates fp16
atomics).  Upsample(f): `feats[idx]` through the cached cell index -- gvf_gather_concat_f16."""
import torch

from .. import ops
from .basic import SparseTensor


def _rescale(scale, f):
    """The level tag of the spatial cache.  The reference computes `s // f` on its (1, 1, 1) start value, which collapses
    every level to 0 (harmless there: its cache holds nothing level-specific).  Here the neighbour maps of the convolutions
    live in that cache, so levels stay distinct: 1 -> 0.5 -> 1."""
    out = []
    for s in scale:
        v = s * f
        out.append(int(v) if float(v).is_integer() else v)
    return tuple(out)


def downsample_plan(x: SparseTensor, factor: int):
    """-> dict(coords [cells, 4] int32, idx [N] int32 cell of every fine row, order [N] int32 fine rows grouped by cell,
    offsets [cells + 1] int32), cached per coordinate set."""
    key = f"downsample_{factor}_plan"
    plan = x.get_spatial_cache(key)
    if plan is None or plan["n"] != x.coords.shape[0]:
        c = x.coords.long()
        cc = c.clone()
        cc[:, 1:] //= factor
        M = int(cc[:, 1:].max()) + 1                       # one host sync per coordinate set
        code = ((cc[:, 0] * M + cc[:, 1]) * M + cc[:, 2]) * M + cc[:, 3]
        ucode, idx = code.unique(return_inverse=True)
        coords = torch.stack([ucode // M ** 3, (ucode // M ** 2) % M, (ucode // M) % M, ucode % M], -1).int()
        order = torch.sort(idx, stable=True).indices.int()
        offsets = torch.zeros(ucode.shape[0] + 1, dtype=torch.int32, device=idx.device)
        offsets[1:] = torch.cumsum(torch.bincount(idx, minlength=ucode.shape[0]), 0)
        counts = torch.bincount(coords[:, 0].long(), minlength=x.shape[0]).tolist()
        off, layout = 0, []
        for cnt in counts:                                  # the coarse level's batch layout, like SparseTensor would derive it
            layout.append(slice(off, off + cnt))
            off += cnt
        plan = dict(n=x.coords.shape[0], coords=coords.contiguous(), idx=idx.int().contiguous(), order=order.contiguous(),
                    offsets=offsets, layout=layout)
        x.register_spatial_cache(key, plan)
    return plan


class SparseDownsample:
    def __init__(self, factor):
        if not isinstance(factor, int):
            raise NotImplementedError("isotropic integer factors only (the flow model uses 2)")
        self.factor = factor

    def forward(self, input: SparseTensor) -> SparseTensor:
        if not (input.feats.is_cuda and input.feats.dtype == torch.float16):
            raise RuntimeError("SparseDownsample runs on CUDA fp16 features only (no CPU fallback)")
        f = self.factor
        plan = downsample_plan(input, f)
        feats = ops.sparse_pool_mean(input.feats, plan["order"], plan["offsets"])
        out = SparseTensor(feats, plan["coords"], torch.Size([input.shape[0], *feats.shape[1:]]), plan["layout"],
                           input._spatial_cache, _rescale(input._scale, 1.0 / f))
        # what SparseUpsample looks up (spatial.py:49-51)
        fac = (f,) * 3
        if out.get_spatial_cache(f"upsample_{fac}_idx") is None:
            out.register_spatial_cache(f"upsample_{fac}_coords", input.coords)
            out.register_spatial_cache(f"upsample_{fac}_layout", input.layout)
            out.register_spatial_cache(f"upsample_{fac}_idx", plan["idx"])
        return out

    __call__ = forward


class SparseUpsample:
    def __init__(self, factor):
        if not isinstance(factor, int):
            raise NotImplementedError("isotropic integer factors only (the flow model uses 2)")
        self.factor = factor

    def forward(self, input: SparseTensor) -> SparseTensor:
        fac = (self.factor,) * 3
        coords = input.get_spatial_cache(f"upsample_{fac}_coords")
        layout = input.get_spatial_cache(f"upsample_{fac}_layout")
        idx = input.get_spatial_cache(f"upsample_{fac}_idx")
        if coords is None or layout is None or idx is None:
            raise ValueError("Upsample cache not found. SparseUpsample must be paired with SparseDownsample.")
        feats = ops.gather_concat(a=input.feats, idx=idx)
        return SparseTensor(feats, coords, torch.Size([input.shape[0], *feats.shape[1:]]), layout, input._spatial_cache,
                            _rescale(input._scale, self.factor))

    __call__ = forward
