"""`SparseTensor`: the minimal container of the voxel-side operators, mirroring the attribute surface of the
reference's sparse/basic.py:18-175 that the path touches (`feats`, `coords`, `shape`, `layout`, `replace`,
`_spatial_cache`).  The reference wraps a spconv.SparseConvTensor; here it is two torch tensors:
feats [N, C] and coords int32 [N, 4] = (batch, x, y, z), rows grouped by batch entry."""
import torch


class SparseTensor:
    def __init__(self, feats, coords, shape=None, layout=None, spatial_cache=None, scale=(1, 1, 1)):
        assert feats.shape[0] == coords.shape[0] and coords.shape[1] == 4
        self.feats = feats
        self.coords = coords.to(torch.int32).contiguous()
        if shape is None:                                   # sparse/basic.py:118-122
            bs = int(self.coords[:, 0].max()) + 1 if coords.shape[0] else 0
            shape = torch.Size([bs, *feats.shape[1:]])
        self.shape = torch.Size(shape)
        if layout is None:                                  # sparse/basic.py:124-128
            counts = torch.bincount(self.coords[:, 0].long(), minlength=self.shape[0]).tolist()
            off, layout = 0, []
            for c in counts:
                layout.append(slice(off, off + c))
                off += c
        self.layout = layout
        self._scale = tuple(scale)
        self._spatial_cache = spatial_cache if spatial_cache is not None else {}

    @property
    def device(self):
        return self.feats.device

    @property
    def dtype(self):
        return self.feats.dtype

    def replace(self, feats, coords=None):
        return SparseTensor(feats, self.coords if coords is None else coords, torch.Size([self.shape[0], *feats.shape[1:]]),
                            self.layout if coords is None else None, self._spatial_cache, self._scale)

    def register_spatial_cache(self, key, value):
        self._spatial_cache.setdefault(str(self._scale), {})[key] = value

    def get_spatial_cache(self, key=None):
        cur = self._spatial_cache.get(str(self._scale), {})
        return cur if key is None else cur.get(key)
