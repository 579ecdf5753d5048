"""vox2seq: oracle pinned to the reference's pure-PyTorch twin (CPU); CUDA kernels vs the oracle and the
size-independent properties of the reference's own test (vox2seq/test.py: full 256^3 grid)."""
import os

import numpy as np
import pytest
import torch

from oracle import vox2seq as O

G = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vox2seq.pt"), weights_only=False)


def test_oracle_matches_reference_twin():
    c, k = G["coords"].numpy(), G["codes"].numpy()
    for mode in ("z_order", "hilbert"):
        for perm in ([0, 1, 2], [2, 0, 1], [2, 1, 0]):
            key = "".join(map(str, perm))
            assert np.array_equal(O.encode(c, perm, mode), G[f"enc_{mode}_{key}"].numpy())
            if f"dec_{mode}_{key}" in G:
                assert np.array_equal(O.decode(k, perm, mode), G[f"dec_{mode}_{key}"].numpy())
            assert np.array_equal(O.decode(O.encode(c, perm, mode), perm, mode), c)      # true inverse


def test_oracle_edge_cases():
    e = np.zeros((0, 3), np.int32)
    assert O.encode(e).shape == (0,) and O.decode(np.zeros(0, np.int32)).shape == (0, 3)
    top = np.array([[1023, 1023, 1023], [0, 0, 0], [1023, 0, 0]], np.int32)
    assert O.encode(top, mode="z_order").tolist() == [2 ** 30 - 1, 0, int("100" * 10, 2)]


@pytest.mark.gpu
def test_cuda_matches_oracle_and_golden():
    from gvfdiffusion_b200 import vox2seq as V
    c, k = G["coords"].cuda(), G["codes"].cuda()
    for mode in ("z_order", "hilbert"):
        for perm in ([0, 1, 2], [2, 0, 1], [2, 1, 0]):
            key = "".join(map(str, perm))
            assert torch.equal(V.encode(c, perm, mode).cpu(), G[f"enc_{mode}_{key}"])
            assert np.array_equal(V.decode(k, perm, mode).cpu().numpy(), O.decode(G["codes"].numpy(), perm, mode))
    assert V.encode(torch.zeros((0, 3), dtype=torch.int32, device="cuda")).shape == (0,)


@pytest.mark.gpu
def test_cuda_full_256_grid_properties():
    from gvfdiffusion_b200 import vox2seq as V
    R = 256
    ax = torch.arange(R, device="cuda", dtype=torch.int32)
    coords = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3)
    for mode in ("z_order", "hilbert"):
        code = V.encode(coords, mode=mode)
        assert int(code.min()) == 0 and int(code.max()) == R ** 3 - 1
        assert torch.equal(torch.sort(code.long())[0], torch.arange(R ** 3, device="cuda"))     # a bijection
        assert torch.equal(V.decode(code, mode=mode), coords)                                  # round trip
    walk = V.decode(torch.arange(R ** 3, device="cuda", dtype=torch.int32), mode="hilbert")
    step = (walk[1:] - walk[:-1]).abs().sum(-1)
    assert int(step.max()) == 1 and int(step.min()) == 1          # Hilbert curve: consecutive codes are face neighbours
