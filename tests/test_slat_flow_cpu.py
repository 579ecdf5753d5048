"""The oracle restatement of the TRELLIS structured-latent flow model (oracle/slat_flow.py) against the
output of the reference's own SLatFlowModel class recorded in tests/golden/slat_flow_tiny.pt
(make_golden.py::gen_slat_flow: CPU fp32, two batch entries, ResBlocks with down / upsampling, two transformer blocks,
q / k RMS-norm), the SparseDownsample quirk (mean over count + 1), and two Euler steps of the reference's sampler."""
import os

import torch

from oracle import slat_flow as O

G = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "slat_flow_tiny.pt"), weights_only=False)


def test_forward_matches_reference_class():
    out = O.slat_flow_forward(G["state_dict"], G["cfg"], G["x"], G["coords"], G["t"], G["cond"])
    err = float((out - G["out"]).norm() / G["out"].norm())
    assert err < 1e-5, err


def test_downsample_counts_its_zero_initial_value():
    f, c, idx = O.downsample(G["x"], G["coords"])
    assert torch.equal(c, G["down_coords"])
    assert float((f - G["down_feats"]).abs().max()) < 1e-6
    # the quirk, spelt out: a cell with n fine rows is sum / (n + 1), not the plain mean
    n = torch.bincount(idx)
    plain = torch.zeros_like(f).index_add_(0, idx, G["x"]) / n[:, None]
    assert float((plain * (n / (n + 1.0))[:, None] - G["down_feats"]).abs().max()) < 1e-6


def test_euler_steps_match_reference_sampler():
    B = int(G["coords"][:, 0].max()) + 1
    fn = lambda x, t: O.slat_flow_forward(G["state_dict"], G["cfg"], x, G["coords"], torch.full((B,), t), G["cond"])
    s = O.flow_euler_sample(fn, G["x"], **G["euler_args"])
    err = float((s - G["euler_samples"]).norm() / G["euler_samples"].norm())
    assert err < 1e-5, err


def test_product_mirror_builds_the_reference_block_lists():
    """Constructor bookkeeping without a GPU: block counts / widths of structured_latent_flow.py:128-183."""
    from gvfdiffusion_b200.trellis.models import SLatFlowModel
    m = SLatFlowModel(**G["cfg"], device="cpu")
    sd = G["state_dict"]
    n_in = len({k.split(".")[1] for k in sd if k.startswith("input_blocks.")})
    n_out = len({k.split(".")[1] for k in sd if k.startswith("out_blocks.")})
    assert len(m.input_blocks) == n_in and len(m.out_blocks) == n_out
    for i, b in enumerate(m.input_blocks):
        assert tuple(sd[f"input_blocks.{i}.conv1.conv.weight"].shape) == (b.out_channels, 3, 3, 3, b.channels)
        assert (f"input_blocks.{i}.skip_connection.weight" in sd) == (b.channels != b.out_channels)
    for i, b in enumerate(m.out_blocks):
        assert tuple(sd[f"out_blocks.{i}.conv1.conv.weight"].shape) == (b.out_channels, 3, 3, 3, b.channels)
    assert m.input_blocks[-1].downsample and m.out_blocks[0].upsample


# ---------------------------------------------------------------------------------------------- SLatGaussianDecoder
GD = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "slat_decoder_gs_tiny.pt"), weights_only=False)


def _gvf_names(sd):
    """The reference's TRELLIS decoder under the names of GVF's own static-VAE decoder (same architecture)."""
    out = {}
    for k, v in sd.items():
        if k.startswith("blocks."):
            out["decoder." + k[7:]] = v
        elif k.startswith("input_layer."):
            out["from_latent." + k[12:]] = v
        elif k.startswith("out_layer."):
            out[k] = v
    return out


def test_gaussian_decoder_oracle_matches_reference_class():
    """oracle.sparse_window.vae_decode / oracle.sparse_vae.to_representation (the static-VAE restatements) reproduce the
    reference's own SLatGaussianDecoder: out_layer rows and the raw tensors of every batch entry's Gaussian model."""
    from oracle import sparse_vae as OSV
    from oracle import sparse_window as OSW
    cfg = GD["cfg"]
    rows = OSW.vae_decode(_gvf_names(GD["state_dict"]), cfg["num_blocks"], cfg["model_channels"] // cfg["num_head_channels"],
                          GD["latent"], GD["coords"], cfg["window_size"], precision="fp32", use_fp16=False, norm_output=True)
    assert float((rows - GD["rows"]).norm() / GD["rows"].norm()) < 1e-5
    rc = dict(cfg["representation_config"], reg_mode="soft_invoxel")
    pert = OSV.build_perturbation(rc["num_gaussians"], "soft_invoxel", rc["voxel_size"])
    assert float((pert - GD["state_dict"]["offset_perturbation"]).abs().max()) < 1e-6
    raw = OSV.to_representation(GD["rows"], GD["coords"], rc, cfg["resolution"], pert)
    G, off = rc["num_gaussians"], 0
    for b, rep in enumerate(GD["reps"]):
        n = int((GD["coords"][:, 0] == b).sum()) * G
        for name, want in rep.items():
            assert float((raw[name][off:off + n] - want).abs().max()) < 1e-6, (b, name)
        off += n


# ---------------------------------------------------------------------------------------------- SparseStructureFlowModel
GS = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sparse_structure_flow_tiny.pt"), weights_only=False)


def test_sparse_structure_flow_oracle_matches_reference_class():
    for name, c in GS.items():
        out = O.sparse_structure_flow_forward(c["state_dict"], c["cfg"], c["x"], c["t"], c["cond"])
        assert out.shape == c["out"].shape
        assert float((out - c["out"]).norm() / c["out"].norm()) < 1e-5, name
        # patchify / unpatchify are inverse permutations
        ps = c["cfg"]["patch_size"]
        assert torch.equal(O.unpatchify(O.patchify(c["x"], ps), ps), c["x"])


def test_sparse_structure_decoder_oracle_matches_reference_class():
    c = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sparse_structure_decoder_tiny.pt"), weights_only=False)
    out = O.sparse_structure_decoder_forward(c["state_dict"], c["cfg"], c["z"])
    assert out.shape == c["out"].shape
    assert float((out - c["out"]).norm() / c["out"].norm()) < 1e-5


def test_downsample_plan_bookkeeping_matches_oracle_on_the_host():
    """The integer part of SparseDownsample (cell codes, unique, fine rows grouped by cell, coarse layout) is torch index work
    and runs anywhere: it must reproduce the reference's coarse coordinates and cell index (through the oracle, which is
    pinned to the reference's own SparseDownsample by the fixture)."""
    from gvfdiffusion_b200.sparse.basic import SparseTensor
    from gvfdiffusion_b200.sparse.spatial import downsample_plan
    coords = G["coords"]
    st = SparseTensor(torch.zeros(coords.shape[0], 8), coords)
    plan = downsample_plan(st, 2)
    _, c_ref, idx_ref = O.downsample(G["x"], coords)
    assert torch.equal(plan["coords"], c_ref) and torch.equal(plan["idx"].long(), idx_ref)
    assert torch.equal(plan["coords"], G["down_coords"])
    # order / offsets: the fine rows of cell p, ascending
    off = plan["offsets"].tolist()
    order = plan["order"].long()
    for p in (0, 1, len(off) - 2):
        rows = order[off[p]:off[p + 1]]
        assert torch.equal(rows, (idx_ref == p).nonzero().flatten())
    assert [s.stop - s.start for s in plan["layout"]] == torch.bincount(c_ref[:, 0].long()).tolist()
    assert downsample_plan(st, 2) is plan                              # cached on the tensor's spatial cache
