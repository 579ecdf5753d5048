"""ORACLE (test infrastructure, never on the product path): CPU restatement of

  * SparseVAE.to_representation / _build_perturbation / _calc_layout (reference
    model/sparse_voxel_diffusion/sparse_vae.py:104-112,114-180,202-227) -- pinned by
    tests/golden/to_representation.pt, produced by the reference's own method
    (tests/golden/make_golden.py gen_to_representation);
  * submanifold sparse convolution (reference sparse/conv/conv_spconv.py:6-15 -> spconv.SubMConv3d).
    spconv is a third-party dependency absent from /root/reference and from this image (setup.sh installs
    `spconv-cu120`, un-pinned): "parity unpinned".  Restated from its published semantics -- the
    output at an active site is the dense cross-correlation (torch conv3d, zero padding ks // 2, weight
    [Cout, kx, ky, kz, Cin]) of the zero-filled grid evaluated at that site; inactive sites stay inactive.
"""
import torch

PRIMES = [2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53]
ORDER = (("_xyz", 3), ("_features_dc", 3), ("_scaling", 3), ("_rotation", 4), ("_opacity", 1))


def radical_inverse(base, n):
    val, inv_base = 0.0, 1.0 / base
    inv_base_n = inv_base
    while n > 0:
        val += (n % base) * inv_base_n
        n //= base
        inv_base_n *= inv_base
    return val


def build_perturbation(num_gaussians, reg_mode, mipgs_voxel_size):
    """:104-112"""
    off = torch.tensor([[n / num_gaussians] + [radical_inverse(PRIMES[d], n) for d in range(2)]
                        for n in range(num_gaussians)]).float() - 0.5
    if reg_mode == "soft_invoxel":
        off = off / 0.5 / mipgs_voxel_size
    return torch.atanh(off)


def to_representation(feats, coords, cfg, resolution, perturbation, kind="MipGS", start=0):
    """feats [N, C] fp32, coords [N, 4] int -> dict of raw GaussianModel tensors for all N voxels (:143-158 /
    :165-180; the per-batch-entry split is a row slice)."""
    G = cfg["num_gaussians"]
    xyz = (coords[:, 1:].float() + 0.5) / resolution
    out = {}
    for name, w in ORDER:
        f = feats[:, start:start + G * w]
        start += G * w
        if name == "_xyz":
            offset = f.reshape(-1, G, 3) * cfg["lr"][name]
            if cfg["perturb_offset"]:
                offset = offset + perturbation
            if cfg["reg_mode"] == "invoxel":
                offset = torch.tanh(offset) / resolution
            elif cfg["reg_mode"] == "soft_invoxel":
                offset = torch.tanh(offset) / resolution * 0.5 * (1.25 if kind == "GS" else cfg["voxel_size"])
            out[name] = (xyz.unsqueeze(1) + offset).flatten(0, 1)
        else:
            shape = (G, 1, 3) if name == "_features_dc" else (G, w)
            out[name] = f.reshape(-1, *shape).flatten(0, 1) * cfg["lr"][name]
    return out


def neighbor_map(coords, ksize=3, dilation=1):
    """coords [N,4] int -> [N, ksize^3] int64 row index of the neighbour (k = (kx*ks + ky)*ks + kz) or -1."""
    table = {tuple(c): i for i, c in enumerate(coords.tolist())}
    half = ksize // 2
    out = torch.full((coords.shape[0], ksize ** 3), -1, dtype=torch.int64)
    for i, (b, x, y, z) in enumerate(coords.tolist()):
        for k in range(ksize ** 3):
            kx, ky, kz = k // (ksize * ksize), (k // ksize) % ksize, k % ksize
            out[i, k] = table.get((b, x + dilation * (kx - half), y + dilation * (ky - half), z + dilation * (kz - half)), -1)
    return out


def subm_conv3d(feats, coords, weight, bias, batch_size, grid_size, dilation=1):
    """feats [N, Cin], coords [N,4] int, weight [Cout, k, k, k, Cin] -> [N, Cout] through a dense conv3d."""
    N, Cin = feats.shape
    ks = weight.shape[1]
    c = coords.long()
    dense = torch.zeros(batch_size, Cin, grid_size, grid_size, grid_size, dtype=feats.dtype)
    dense[c[:, 0], :, c[:, 1], c[:, 2], c[:, 3]] = feats
    y = torch.nn.functional.conv3d(dense, weight.permute(0, 4, 1, 2, 3).contiguous(), bias,
                                   padding=dilation * (ks // 2), dilation=dilation)
    return y[c[:, 0], :, c[:, 1], c[:, 2], c[:, 3]]
