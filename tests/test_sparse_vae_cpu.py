"""Pins oracle/sparse_vae.py: to_representation against the fixture produced by the reference's own
SparseVAE.to_representation (tests/golden/to_representation.pt), the neighbour map / submanifold convolution
restatement against each other (gather form == dense conv3d form), and the host-side layout mirror."""
import os

import torch

from oracle import sparse_vae as OSV

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ("_xyz", "_features_dc", "_scaling", "_rotation", "_opacity")


def test_to_representation_oracle_matches_reference_fixture():
    g = torch.load(os.path.join(G, "to_representation.pt"), weights_only=False)
    cfg = g["cfg"]
    pert = OSV.build_perturbation(cfg["num_gaussians"], cfg["reg_mode"], cfg["voxel_size"])
    assert torch.equal(pert, g["perturbation"])
    out = OSV.to_representation(g["feats"], g["coords"], cfg, g["resolution"], pert)
    off = 0
    for n, rep in zip(g["counts"], g["reps"]):
        for name in NAMES:
            assert torch.equal(out[name][off * 8:(off + n) * 8], rep[name]), name      # same torch ops: bit-exact
        off += n
    # layout table of _calc_layout
    start = 0
    for name, w in OSV.ORDER:
        assert g["layout"][name] == (start, start + 8 * w)
        start += 8 * w


def test_host_mirror_layout_and_perturbation_match_fixture():
    from gvfdiffusion_b200.model.sparse_voxel_diffusion.sparse_vae import SparseVAE
    g = torch.load(os.path.join(G, "to_representation.pt"), weights_only=False)
    sv = SparseVAE(resolution=64, representation_config={"MipGS": g["cfg"]}, device="cpu")
    assert sv.out_channels == 112
    assert {k: tuple(v["range"]) for k, v in sv.layouts["MipGS"].items()} == g["layout"]
    assert torch.equal(sv.perturbation["MipGS"], g["perturbation"])


def test_subm_conv_gather_form_equals_dense_form():
    g = torch.Generator().manual_seed(3)
    res, n, cin, cout = 10, 120, 8, 16
    coords = []
    for b in range(2):
        lin = torch.randperm(res ** 3, generator=g)[:n]
        coords.append(torch.stack([torch.full((n,), b), lin // (res * res), (lin // res) % res, lin % res], 1))
    coords = torch.cat(coords).int()
    x = torch.randn(2 * n, cin, generator=g)
    w = torch.randn(cout, 3, 3, 3, cin, generator=g) * 0.1
    bias = torch.randn(cout, generator=g)
    ref = OSV.subm_conv3d(x, coords, w, bias, 2, res)
    nbr = OSV.neighbor_map(coords, 3)
    xz = torch.cat([x, torch.zeros(1, cin)])
    cols = xz[nbr.clamp_min(-1)].reshape(2 * n, 27 * cin)          # index -1 = the appended zero row
    out = cols @ w.reshape(cout, -1).t() + bias
    assert (out - ref).abs().max() < 1e-4
    assert int((nbr[:, 13] == torch.arange(2 * n)).all()) == 1      # centre tap is the voxel itself


def test_sparse_tensor_container_layout_and_cache():
    from gvfdiffusion_b200.sparse.basic import SparseTensor
    coords = torch.tensor([[0, 1, 2, 3], [0, 4, 5, 6], [1, 0, 0, 0], [2, 7, 7, 7], [2, 1, 1, 1]])
    x = SparseTensor(torch.arange(10.0).reshape(5, 2), coords)
    assert x.shape == torch.Size([3, 2]) and x.coords.dtype == torch.int32
    assert [(s.start, s.stop) for s in x.layout] == [(0, 2), (2, 3), (3, 5)]
    y = x.replace(torch.zeros(5, 7))
    assert y.shape == torch.Size([3, 7]) and y.layout == x.layout and y.coords is x.coords
    x.register_spatial_cache("k", 5)
    assert y.get_spatial_cache("k") == 5            # the cache is shared by tensors over the same coordinates


def test_sparse_conv_mirror_rejects_what_is_not_on_the_path():
    import pytest
    from gvfdiffusion_b200.sparse.basic import SparseTensor
    from gvfdiffusion_b200.sparse.conv import SparseConv3d
    with pytest.raises(NotImplementedError):
        SparseConv3d(8, 8, 3, stride=2, device="cpu")
    with pytest.raises(ValueError):
        SparseConv3d(6, 8, 3, device="cpu")
    conv = SparseConv3d(8, 16, 3, device="cpu").load_state_dict(
        {"p.conv.weight": torch.randn(16, 3, 3, 3, 8), "p.conv.bias": torch.randn(16)}, prefix="p.")
    assert conv.weight.shape == (16, 27 * 8) and conv.weight.dtype == torch.float16
    with pytest.raises(RuntimeError, match="device only"):
        conv(SparseTensor(torch.zeros(2, 8), torch.tensor([[0, 0, 0, 0], [0, 1, 1, 1]])))


def test_sparse_transformer_vae_oracle_matches_reference_fixture():
    """oracle/sparse_window.py (window partition + per-window attention + swin blocks + SparseTransformerVAE
    encode / decode) against the reference's own SparseTransformerVAE run on the CPU in fp32
    (tests/golden/sparse_vae_tiny.pt, tests/golden/make_golden.py gen_sparse_vae)."""
    import torch.nn.functional as F
    from oracle import sparse_window as OSW
    from oracle.dit import absolute_position_embedding
    g = torch.load(os.path.join(G, "sparse_vae_tiny.pt"), weights_only=False)
    cfg, sd, coords = g["cfg"], g["state_dict"], g["coords"]
    H = cfg["model_channels"] // cfg["num_head_channels"]
    dec = OSW.vae_decode(sd, cfg["num_blocks"], H, g["latent"], coords, cfg["window_size"], precision="fp32",
                         use_fp16=False, norm_output=True)
    assert dec.shape == g["decode"].shape
    assert float((dec - g["decode"]).norm() / g["decode"].norm()) < 2e-5
    # encode (:151-176): input_layer + APE -> encoder blocks -> layer_norm -> to_latent -> (mean, logvar)
    C = cfg["model_channels"]
    h = F.linear(g["feats"], sd["input_layer.weight"], sd["input_layer.bias"])
    h = h + absolute_position_embedding(coords[:, 1:].float()[None], C)[0]
    h = OSW.transformer_blocks(sd, "encoder.", cfg["num_blocks"], H, h, coords, cfg["window_size"], "fp32")
    h = F.linear(F.layer_norm(h, (C,)), sd["to_latent.weight"], sd["to_latent.bias"])
    mean, logvar = h.chunk(2, dim=-1)
    assert float((mean - g["mean"]).norm() / g["mean"].norm()) < 2e-5
    assert float((logvar - g["logvar"]).norm() / g["logvar"].norm()) < 2e-5


def test_old_attn_impl_layout_oracle_and_weight_permutation():
    """use_old_attn_impl=True ([H][3][d] qkv channels, the class default; the shipped configs use false): the oracle
    restates it, and the device mirror's load-time row permutation of to_qkv yields the same q, k, v."""
    import torch.nn.functional as F
    from gvfdiffusion_b200.sparse.transformer import qkv_rows_from_old_attn_impl
    from oracle import sparse_window as OSW
    g = torch.load(os.path.join(G, "sparse_vae_tiny.pt"), weights_only=False)
    cfg, sd, coords = g["cfg"], g["state_dict"], g["coords"]
    H = cfg["model_channels"] // cfg["num_head_channels"]
    dec = OSW.vae_decode(sd, cfg["num_blocks"], H, g["latent"], coords, cfg["window_size"], precision="fp32",
                         use_fp16=False, norm_output=True, old_attn_impl=True)
    ref = g["decode_old_attn_impl"]
    assert float((dec - ref).norm() / ref.norm()) < 2e-5
    assert float((g["decode"] - ref).norm() / ref.norm()) > 0.1          # the two layouts are different functions
    # permuted rows + the [3][H][d] reading == original rows + the [H][3][d] reading, bit for bit
    w, b = sd["decoder.0.attn.to_qkv.weight"], sd["decoder.0.attn.to_qkv.bias"]
    wp, bp = qkv_rows_from_old_attn_impl(w, b, H)
    x = torch.randn(37, w.shape[1], generator=torch.Generator().manual_seed(4))
    d = w.shape[0] // (3 * H)
    old = F.linear(x, w, b).reshape(-1, H, 3, d).permute(0, 2, 1, 3)
    new = F.linear(x, wp, bp).reshape(-1, 3, H, d)
    assert torch.equal(old, new)
    # and through the whole oracle trunk: permuted state dict in the default layout == old layout on the original
    sd2 = dict(sd)
    for i in range(cfg["num_blocks"]):
        k = f"decoder.{i}.attn.to_qkv."
        sd2[k + "weight"], sd2[k + "bias"] = qkv_rows_from_old_attn_impl(sd[k + "weight"], sd[k + "bias"], H)
    dec2 = OSW.vae_decode(sd2, cfg["num_blocks"], H, g["latent"], coords, cfg["window_size"], precision="fp32",
                          use_fp16=False, norm_output=True)
    assert float((dec2 - ref).norm() / ref.norm()) < 2e-5
