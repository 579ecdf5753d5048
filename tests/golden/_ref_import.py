"""Import the (read-only, untrusted) reference from /root/reference with stubs for
its missing third-party deps.  Used ONLY by tests/golden/make_golden.py, in the
build container; nothing on the GPU box imports this (the reference is absent there).

Stubs (none of them is touched by the code paths we execute -- SURVEY.md section 8c):
  spconv.pytorch.SparseConvTensor   (sparse/basic.py:6, import-time only)
  torch_cluster.fps, pytorch3d.ops.knn_points, timm DropPath/trunc_normal_
  vox2seq, flash_attn is present but we force ATTN_BACKEND=sdpa.
"""
import os
import sys
import types

REF = "/root/reference"


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def install():
    if not os.path.isdir(REF):
        raise RuntimeError("reference tree not present (expected in the build container only)")
    os.environ.setdefault("ATTN_BACKEND", "sdpa")
    os.environ.setdefault("SPARSE_ATTN_BACKEND", "flash_attn")
    import torch

    class _SparseConvTensor:  # never instantiated on our path
        pass

    sp = _stub("spconv")
    spp = _stub("spconv.pytorch", SparseConvTensor=_SparseConvTensor)
    sp.pytorch = spp
    _stub("vox2seq")
    _stub("torch_cluster", fps=lambda *a, **k: (_ for _ in ()).throw(RuntimeError("stub")))
    p3 = _stub("pytorch3d")
    p3o = _stub("pytorch3d.ops", knn_points=None, knn_gather=None, sample_farthest_points=None)
    p3.ops = p3o

    class _DropPath(torch.nn.Module):
        def __init__(self, p=0.0):
            super().__init__()
        def forward(self, x):
            return x

    timm = _stub("timm")
    tm = _stub("timm.models")
    tl = _stub("timm.models.layers", DropPath=_DropPath,
               trunc_normal_=torch.nn.init.trunc_normal_)
    timm.models = tm
    tm.layers = tl
    tl2 = _stub("timm.layers", DropPath=_DropPath, trunc_normal_=torch.nn.init.trunc_normal_)
    timm.layers = tl2
    if REF not in sys.path:
        sys.path.insert(0, REF)
