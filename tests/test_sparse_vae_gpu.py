"""GPU parity of the voxel-side operators (include/gvf_b200.h section 7) through the C ABI:
gvf_to_representation against the fixture of the reference's own SparseVAE.to_representation and against the
oracle at the benchmark size; the submanifold convolution (neighbour map bit-exact, im2col bit-exact, fp16
tcgen05 GEMM within fp16 tolerance) against oracle/sparse_vae.py."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ("_xyz", "_features_dc", "_scaling", "_rotation", "_opacity")


def _voxels(n_per_batch, res, batches, seed):
    g = torch.Generator().manual_seed(seed)
    out = []
    for b in range(batches):
        lin = torch.randperm(res ** 3, generator=g)[:n_per_batch]
        xyz = torch.stack([lin // (res * res), (lin // res) % res, lin % res], 1)
        out.append(torch.cat([torch.full((n_per_batch, 1), b), xyz], 1))
    return torch.cat(out).int()


def test_to_representation_matches_reference_fixture():
    from gvfdiffusion_b200.model.sparse_voxel_diffusion import SparseVAE
    from gvfdiffusion_b200.sparse.basic import SparseTensor
    g = torch.load(os.path.join(G, "to_representation.pt"), weights_only=False)
    sv = SparseVAE(resolution=g["resolution"], representation_config={"MipGS": g["cfg"]}, device=DEV)
    x = SparseTensor(g["feats"].to(DEV), g["coords"].to(DEV))
    assert [(s.start, s.stop) for s in x.layout] == [(0, 37), (37, 57)]
    reps = sv.to_representation(x)["MipGS"]
    assert len(reps) == 2
    for rep, ref, act in zip(reps, g["reps"], g["activated"]):
        for name in NAMES:
            got = getattr(rep, name).cpu()
            assert got.shape == ref[name].shape, name
            if name == "_xyz":           # tanhf (libdevice) vs the host tanh: a few ulp of an offset <= 0.012
                assert (got - ref[name]).abs().max() < 2e-7
            else:                        # one IEEE multiply: bit-exact
                assert torch.equal(got, ref[name]), name
        # the activated view the renderer consumes (GaussianModel getters of the reference)
        gt = rep.gaussian_tensor().cpu()
        assert torch.allclose(gt[:, 0:3], act["xyz"], rtol=3e-6, atol=1e-6)
        assert torch.allclose(gt[:, 7:10], act["scaling"], rtol=3e-6, atol=1e-7)
        assert torch.allclose(gt[:, 6:7], act["opacity"], rtol=3e-6, atol=1e-7)
        assert torch.allclose(gt[:, 10:14], act["rotation"], rtol=3e-6, atol=1e-6)


@pytest.mark.parametrize("kind,reg,perturb", [("MipGS", "soft_invoxel", True), ("GS", "invoxel", False),
                                              ("GS", "soft_invoxel", True)])
def test_to_representation_matches_oracle_at_bench_size(kind, reg, perturb):
    from gvfdiffusion_b200.model.sparse_voxel_diffusion import SparseVAE
    from gvfdiffusion_b200.sparse.basic import SparseTensor
    from oracle import sparse_vae as OSV
    cfg = {"lr": {"_xyz": 1.0, "_features_dc": 1.0, "_opacity": 1.0, "_scaling": 1.0, "_rotation": 0.1},
           "perturb_offset": perturb, "reg_mode": reg, "voxel_size": 1.5, "num_gaussians": 8}
    rc = {kind: cfg} if kind == "MipGS" else {kind: cfg, "MipGS": dict(cfg)}
    sv = SparseVAE(resolution=64, representation_config=rc, device=DEV)
    coords = _voxels(2048, 64, 2, seed=5)
    C = sv.out_channels
    feats = torch.randn(coords.shape[0], C, generator=torch.Generator().manual_seed(9))
    reps = sv.to_representation(SparseTensor(feats.to(DEV), coords.to(DEV)))
    full = dict(sv.rep_config[kind])
    pert = OSV.build_perturbation(8, reg, 1.5) if perturb else None
    ref = OSV.to_representation(feats, coords, full, 64, pert, kind=kind, start=sv.layouts[kind]["_xyz"]["range"][0])
    for b in range(2):
        rep = reps[kind][b]
        assert rep._xyz.shape == (2048 * 8, 3) and rep._features_dc.shape == (2048 * 8, 1, 3)
        sl = slice(b * 2048 * 8, (b + 1) * 2048 * 8)
        assert (rep._xyz.cpu() - ref["_xyz"][sl]).abs().max() < 2e-7
        for name in NAMES[1:]:
            assert torch.equal(getattr(rep, name).cpu(), ref[name][sl]), name


@pytest.mark.parametrize("n,res,ks,dil", [(300, 12, 3, 1), (2000, 64, 3, 1), (150, 9, 3, 2), (64, 4, 3, 1), (40, 8, 1, 1),
                                          (200, 10, 5, 1)])
def test_neighbor_map_bit_exact(n, res, ks, dil):
    from gvfdiffusion_b200 import ops
    from oracle import sparse_vae as OSV
    coords = _voxels(n, res, 2, seed=n)
    nbr = ops.sparse_neighbor_map(coords.to(DEV), 2, res, ks, dil).cpu()
    assert torch.equal(nbr.long(), OSV.neighbor_map(coords, ks, dil))


def test_neighbor_map_flags_bad_coordinates():
    from gvfdiffusion_b200.sparse.basic import SparseTensor
    from gvfdiffusion_b200.sparse.conv import SparseConv3d
    conv = SparseConv3d(8, 8, 3, device=DEV)
    coords = torch.tensor([[0, 1, 1, 1], [0, 1, 1, 1], [0, 2, 2, 2]], dtype=torch.int32, device=DEV)
    with pytest.raises(ValueError, match="duplicate"):
        conv(SparseTensor(torch.zeros(3, 8, device=DEV), coords), grid_size=4)
    coords = torch.tensor([[0, 1, 1, 1], [0, 9, 1, 1]], dtype=torch.int32, device=DEV)
    with pytest.raises(ValueError, match="outside"):
        conv(SparseTensor(torch.zeros(2, 8, device=DEV), coords), grid_size=4)


@pytest.mark.parametrize("n,res,cin,cout,dtype", [(500, 16, 32, 64, torch.float32), (3000, 64, 128, 128, torch.float16),
                                                  (77, 6, 8, 8, torch.float32), (1200, 32, 64, 136, torch.float16)])
def test_subm_conv_matches_oracle(n, res, cin, cout, dtype):
    from gvfdiffusion_b200 import ops
    from gvfdiffusion_b200.sparse.basic import SparseTensor
    from gvfdiffusion_b200.sparse.conv import SparseConv3d
    from oracle import sparse_vae as OSV
    g = torch.Generator().manual_seed(n + cin)
    coords = _voxels(n, res, 2, seed=n)
    x = torch.randn(2 * n, cin, generator=g).to(dtype)
    w = torch.randn(cout, 3, 3, 3, cin, generator=g) * (1.0 / (27 * cin) ** 0.5)
    bias = torch.randn(cout, generator=g) * 0.1
    conv = SparseConv3d(cin, cout, 3, indice_key="t", device=DEV).load_state_dict({"conv.weight": w, "conv.bias": bias})
    xs = SparseTensor(x.to(DEV), coords.to(DEV))
    y = conv(xs, grid_size=res)
    assert y.feats.dtype == torch.float16 and y.feats.shape == (2 * n, cout) and y.layout == xs.layout
    # im2col operand is an exact gather of the fp16-rounded features
    nbr = conv.neighbor_map(xs)
    cols = ops.sparse_im2col(xs.feats, nbr).cpu()
    onbr = OSV.neighbor_map(coords, 3)
    xz = torch.cat([x.half(), torch.zeros(1, cin, dtype=torch.float16)])
    assert torch.equal(cols, xz[onbr].reshape(2 * n, 27 * cin))
    # the contraction: fp16 operands, fp32 accumulation, one fp16 rounding
    ref = OSV.subm_conv3d(x.half().float(), coords, w.half().float(), bias, 2, res)
    err = float((y.feats.float().cpu() - ref).norm() / ref.norm())
    assert err < 1e-3, err
    # second call reuses the cached neighbour map (same object)
    assert conv.neighbor_map(xs) is nbr


@pytest.mark.parametrize("n,res,cin,cout", [(3000, 64, 128, 128), (1200, 32, 64, 136), (5000, 64, 64, 256), (130, 8, 192, 64)])
def test_subm_conv_gather_fused_is_bit_identical_to_im2col(n, res, cin, cout):
    """The GEMM whose TMA producer gathers the neighbour rows itself (tile::gather4, absent voxels zero-filled) against the
    materialised im2col operand + plain GEMM: same k order, same tiles -> identical bits; and against the oracle."""
    from gvfdiffusion_b200 import ops
    from gvfdiffusion_b200.sparse.basic import SparseTensor
    from gvfdiffusion_b200.sparse.conv import SparseConv3d
    from oracle import sparse_vae as OSV
    g = torch.Generator().manual_seed(n + cin + 1)
    coords = _voxels(n, res, 2, seed=n + 1)
    x = torch.randn(2 * n, cin, generator=g).half()
    w = torch.randn(cout, 3, 3, 3, cin, generator=g) * (1.0 / (27 * cin) ** 0.5)
    bias = torch.randn(cout, generator=g) * 0.1
    conv = SparseConv3d(cin, cout, 3, indice_key="g", device=DEV).load_state_dict({"conv.weight": w, "conv.bias": bias})
    xs = SparseTensor(x.to(DEV), coords.to(DEV))
    fused = conv(xs, grid_size=res).feats
    conv.fused_gather = False
    two = conv(xs, grid_size=res).feats
    assert torch.equal(fused, two)
    f32 = ops.sparse_conv_gemm(xs.feats, conv.neighbor_map(xs), conv.weight, conv.bias, out_f32=True)
    ref = OSV.subm_conv3d(x.float(), coords, w.half().float(), bias, 2, res)
    assert float((f32.cpu() - ref).norm() / ref.norm()) < 1e-4


@pytest.mark.parametrize("kind,reg,perturb", [("MipGS", "soft_invoxel", True), ("GS", "invoxel", False), ("GS", "none", False)])
def test_to_representation_backward_matches_autograd(kind, reg, perturb):
    """gvf_to_representation_bwd (through SparseVAE.to_representation under autograd) against torch autograd of the oracle
    restatement: fp32, 1e-5 relative."""
    from gvfdiffusion_b200.model.sparse_voxel_diffusion import SparseVAE
    from gvfdiffusion_b200.sparse.basic import SparseTensor
    from oracle import sparse_vae as OSV
    cfg = {"lr": {"_xyz": 1.0, "_features_dc": 0.5, "_opacity": 0.05, "_scaling": 0.25, "_rotation": 0.1},
           "perturb_offset": perturb, "reg_mode": reg, "voxel_size": 1.5, "num_gaussians": 8}
    rc = {kind: cfg} if kind == "MipGS" else {kind: cfg, "MipGS": dict(cfg)}
    sv = SparseVAE(resolution=64, representation_config=rc, device=DEV)
    coords = _voxels(700, 64, 2, seed=6)
    gen = torch.Generator().manual_seed(10)
    feats = torch.randn(coords.shape[0], sv.out_channels, generator=gen)
    fd = feats.to(DEV).requires_grad_(True)
    reps = sv.to_representation(SparseTensor(fd, coords.to(DEV)))[kind]
    P = coords.shape[0] * 8
    w = {n: torch.randn(P, k, generator=gen) for n, k in (("_xyz", 3), ("_features_dc", 3), ("_scaling", 3), ("_rotation", 4), ("_opacity", 1))}
    loss = 0
    for b, rep in enumerate(reps):
        sl = slice(b * 700 * 8, (b + 1) * 700 * 8)
        for n in NAMES:
            loss = loss + (getattr(rep, n).reshape(700 * 8, -1) * w[n][sl].to(DEV)).sum()
    loss.backward()
    fr = feats.clone().requires_grad_(True)
    full = dict(sv.rep_config[kind])
    pert = OSV.build_perturbation(8, reg, 1.5) if perturb else None
    lo = sv.layouts[kind]["_xyz"]["range"][0]
    ref = OSV.to_representation(fr, coords, full, 64, pert, kind=kind, start=lo)
    sum((ref[n].reshape(P, -1) * w[n]).sum() for n in NAMES).backward()
    got = fd.grad.cpu()
    assert (got[:, :lo] == 0).all() and (got[:, lo + 112:] == 0).all()
    err = (got - fr.grad).abs().max() / fr.grad.abs().max()
    assert err < 1e-5, float(err)


def test_sparse_vae_training_losses_terms_and_gradients():
    """SparseVAE.training_losses (sparse_vae.py:303-362) over the nn.Module backbone: the terms add up the way the reference
    adds them, every backbone parameter receives a finite gradient through render -> to_representation -> decoder ->
    posterior -> encoder, and the LPIPS term (seeded-random VGG16) changes the loss and the gradients."""
    from gvfdiffusion_b200 import synthetic as S
    from gvfdiffusion_b200.model.sparse_voxel_diffusion import SparseTransformerVAE, SparseVAE
    from gvfdiffusion_b200.sparse.basic import SparseTensor
    torch.manual_seed(4)
    rep = {"MipGS": {"lr": {"_xyz": 1.0, "_features_dc": 1.0, "_opacity": 1.0, "_scaling": 1.0, "_rotation": 0.1},
                     "perturb_offset": True, "reg_mode": "soft_invoxel", "voxel_size": 1.5, "num_gaussians": 8,
                     "2d_filter_kernel_size": 0.1, "3d_filter_kernel_size": 0.0009, "scaling_bias": 0.004, "opacity_bias": 0.1,
                     "scaling_activation": "softplus"}}
    reg = {"MipGS": {"lambda_vol": 10000.0, "lambda_opacity": 0.001}}
    m = SparseTransformerVAE(32, 64, 128, 112, 8, 2, window_size=8, use_fp16=True, use_old_attn_impl=False, norm_output=True).to(DEV)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.startswith(("to_latent", "out_layer")) and n.endswith("weight"):
                p.normal_(0, 0.05)
    coords = _voxels(300, 32, 2, seed=8).to(DEV)
    g = torch.Generator().manual_seed(9)
    x = SparseTensor(torch.randn(600, 64, generator=g).to(DEV), coords)
    noise = torch.randn(600, 8, generator=g).to(DEV)
    ext, intr = S.orbit_extrinsics(2, radius=1.2), S.intrinsics(40.0)[None].repeat(2, 1, 1)
    image = torch.rand(2, 3, 64, 64, generator=g).to(DEV)
    out = {}
    for lam in (0.0, 0.2):
        fw = SparseVAE({"vae": m}, resolution=32, representation_config=rep, device=DEV, lambda_ssim=0.2, lambda_lpips=lam,
                       lamda_kl=1e-6, regularizations=reg)
        for p in m.parameters():
            p.grad = None
        terms, reps = fw.training_losses(x, image, ext, intr, noise=noise)
        assert len(reps["MipGS"]) == 2 and reps["MipGS"][0]._xyz.shape == (300 * 8, 3)
        rec = terms["MipGS_l1"] + 0.2 * terms["MipGS_ssim"] + (lam * terms["MipGS_lpips"] if lam else 0.0)
        total = rec + 1e-6 * terms["kl"] + 10000.0 * terms["reg_MipGS_vol"] + 0.001 * terms["reg_MipGS_opacity"]
        assert abs(float(terms["loss"].detach()) - float(total.detach())) < 1e-6 * abs(float(total.detach())) + 1e-7
        (terms["loss"] * 65536.0).backward()
        grads = {n: p.grad.clone() for n, p in m.named_parameters()}
        assert all(v is not None and torch.isfinite(v).all() for v in grads.values())
        assert float(grads["encoder.0.attn.to_qkv.weight"].abs().max()) > 0 and float(grads["out_layer.weight"].abs().max()) > 0
        out[lam] = (float(terms["loss"].detach()), grads)
    assert out[0.2][0] > out[0.0][0]
    assert float((out[0.2][1]["out_layer.weight"] - out[0.0][1]["out_layer.weight"]).abs().max()) > 0
