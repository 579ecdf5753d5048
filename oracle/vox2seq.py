"""CPU oracle (numpy) for vox2seq (TEST INFRASTRUCTURE, see oracle/__init__.py): Morton and Hilbert
(Skilling 2004) codes, 10 bits per axis, restating reference
model/sparse_voxel_diffusion/vox2seq/src/z_order.cu:13-65 and hilbert.cu:35-135 and the call
conventions of vox2seq/__init__.py:7-50.  Pinned to the reference's own pure-PyTorch twin
(vox2seq/pytorch/, the oracle of the reference's only known-answer test, vox2seq/test.py)."""
import numpy as np


def _spread3(v):
    v = v.astype(np.uint32)
    v = (v * np.uint32(0x00010001)) & np.uint32(0xFF0000FF)
    v = (v * np.uint32(0x00000101)) & np.uint32(0x0F00F00F)
    v = (v * np.uint32(0x00000011)) & np.uint32(0xC30C30C3)
    v = (v * np.uint32(0x00000005)) & np.uint32(0x49249249)
    return v


def _gather3(v):
    v = v.astype(np.uint32) & np.uint32(0x49249249)
    v = (v ^ (v >> np.uint32(2))) & np.uint32(0x030C30C3)
    v = (v ^ (v >> np.uint32(4))) & np.uint32(0x0300F00F)
    v = (v ^ (v >> np.uint32(8))) & np.uint32(0x030000FF)
    v = (v ^ (v >> np.uint32(16))) & np.uint32(0x000003FF)
    return v


def encode(coords, permute=(0, 1, 2), mode="z_order"):
    X = [coords[:, p].astype(np.uint32).copy() for p in permute]
    if mode == "hilbert":
        Q = np.uint32(1 << 9)
        while Q > 1:
            P = np.uint32(Q - 1)
            for d in range(3):
                hit = (X[d] & Q) != 0
                t = (X[0] ^ X[d]) & P
                x0 = np.where(hit, X[0] ^ P, X[0] ^ t)
                xd = np.where(hit, X[d], X[d] ^ t)
                if d == 0:
                    X[0] = np.where(hit, X[0] ^ P, X[0])      # t == 0 when d == 0
                else:
                    X[0], X[d] = x0, xd
            Q = np.uint32(Q >> 1)
        X[1] ^= X[0]
        X[2] ^= X[1]
        t = np.zeros_like(X[0])
        Q = np.uint32(1 << 9)
        while Q > 1:
            t = np.where((X[2] & Q) != 0, t ^ np.uint32(Q - 1), t)
            Q = np.uint32(Q >> 1)
        X = [x ^ t for x in X]
    elif mode != "z_order":
        raise ValueError(mode)
    return (_spread3(X[0]) * np.uint32(4) + _spread3(X[1]) * np.uint32(2) + _spread3(X[2])).astype(np.int32)


def decode(codes, permute=(0, 1, 2), mode="z_order"):
    c = codes.astype(np.uint32)
    X = [_gather3(c >> np.uint32(2)), _gather3(c >> np.uint32(1)), _gather3(c)]
    if mode == "hilbert":
        t = X[2] >> np.uint32(1)
        X[2] = X[2] ^ X[1]
        X[1] = X[1] ^ X[0]
        X[0] = X[0] ^ t
        Q = np.uint32(2)
        while Q != np.uint32(2 << 9):
            P = np.uint32(Q - 1)
            for d in (2, 1, 0):
                hit = (X[d] & Q) != 0
                t = (X[0] ^ X[d]) & P
                if d == 0:
                    X[0] = np.where(hit, X[0] ^ P, X[0])
                else:
                    x0 = np.where(hit, X[0] ^ P, X[0] ^ t)
                    X[d] = np.where(hit, X[d], X[d] ^ t)
                    X[0] = x0
            Q = np.uint32(Q << 1)
    out = np.zeros((c.shape[0], 3), np.int32)
    for k, p in enumerate(permute):
        out[:, p] = X[k].astype(np.int32)
    return out
