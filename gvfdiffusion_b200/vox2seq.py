"""`vox2seq.encode / decode`: call-compatible with the reference's first-party extension
(model/sparse_voxel_diffusion/vox2seq/vox2seq/__init__.py:7-50) on the sm_100a kernels."""
import ctypes as C

import torch

from . import _lib


def _perm(permute):
    p = [int(x) for x in permute]
    assert sorted(p) == [0, 1, 2], "permute must be a permutation of [0, 1, 2]"
    return (C.c_int * 3)(*p)


@torch.no_grad()
def encode(coords, permute=(0, 1, 2), mode="z_order"):
    """coords [N,3] integer tensor on CUDA -> int32 codes [N] (30 bits)."""
    assert coords.shape[-1] == 3 and coords.ndim == 2, "Input coordinates must be of shape [N, 3]"
    if mode not in ("z_order", "hilbert"):
        raise ValueError(f"Unknown encoding mode: {mode}")
    c = coords.to(torch.int32).contiguous()
    out = torch.empty(c.shape[0], dtype=torch.int32, device=c.device)
    if c.shape[0] == 0:
        return out
    _lib.check(_lib.lib().gvf_vox2seq_encode(_lib.ptr(c), c.shape[0], _perm(permute), int(mode == "hilbert"),
                                             _lib.ptr(out), _lib.current_stream()), "gvf_vox2seq_encode")
    return out


@torch.no_grad()
def decode(code, permute=(0, 1, 2), mode="z_order"):
    """int32 codes [N] -> coords [N,3] int32."""
    assert code.ndim == 1, "Input code must be of shape [N]"
    if mode not in ("z_order", "hilbert"):
        raise ValueError(f"Unknown decoding mode: {mode}")
    c = code.to(torch.int32).contiguous()
    out = torch.empty((c.shape[0], 3), dtype=torch.int32, device=c.device)
    if c.shape[0] == 0:
        return out
    _lib.check(_lib.lib().gvf_vox2seq_decode(_lib.ptr(c), c.shape[0], _perm(permute), int(mode == "hilbert"),
                                             _lib.ptr(out), _lib.current_stream()), "gvf_vox2seq_decode")
    return out
