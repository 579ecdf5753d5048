"""Serialized and full sparse attention (reference sparse/attention/serialized_attn.py:38-192,
sparse/attention/full_attn.py:90-215) on the general varlen kernel (gvf_sparse_varlen_attn_f16) against a torch fp32
restatement of what the reference computes: gather `qkv.feats[fwd_indices]`, softmax attention inside every sequence,
`out[bwd_indices]`.  Tolerance 2e-3 of the output's max (fp16 outputs)."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "serialization.pt")


def _sdpa(q, k, v):
    # [L, H, C] fp32
    o = F.scaled_dot_product_attention(q.transpose(0, 1)[None], k.transpose(0, 1)[None], v.transpose(0, 1)[None])
    return o[0].transpose(0, 1)


def _close(a, b, tol=2e-3):
    err = (a.float() - b.float()).abs().max().item()
    assert err <= tol * max(b.float().abs().max().item(), 1e-6), err


def test_serialized_attention_matches_reference_semantics_and_partition():
    from gvfdiffusion_b200.sparse.attention import SerializeMode, calc_serialization, \
        sparse_serialized_scaled_dot_product_self_attention
    from gvfdiffusion_b200.sparse.basic import SparseTensor
    G = torch.load(GOLD, weights_only=False)
    gen = torch.Generator().manual_seed(3)
    for case in G:
        coords = case["coords"].to(DEV)
        T, H, C = coords.shape[0], 3, 64
        feats = (torch.randn(T, 3, H, C, generator=gen) * 0.7).half().to(DEV)
        st = SparseTensor(feats, coords)
        mode = SerializeMode[case["mode"]]
        # the device partition (sm_100a vox2seq codes) is the reference's, index for index
        fwd, bwd, seq_lens, seq_batch = calc_serialization(st, case["window"], mode, case["shift_sequence"], case["shift_window"])
        assert torch.equal(fwd.cpu(), case["fwd"]) and torch.equal(bwd.cpu(), case["bwd"])
        assert list(seq_lens) == case["seq_lens"] and list(seq_batch) == case["seq_batch_indices"]
        out = sparse_serialized_scaled_dot_product_self_attention(st, case["window"], mode, case["shift_sequence"],
                                                                  case["shift_window"])
        g = feats.float()[fwd]                                   # [M, 3, H, C]
        ref, s0 = [], 0
        for n in seq_lens:
            ref.append(_sdpa(g[s0:s0 + n, 0], g[s0:s0 + n, 1], g[s0:s0 + n, 2]))
            s0 += n
        ref = torch.cat(ref)[bwd]
        assert out.feats.shape == (T, H, C)
        _close(out.feats, ref)


def test_full_sparse_self_and_cross_attention():
    from gvfdiffusion_b200.sparse.attention import sparse_scaled_dot_product_attention
    from gvfdiffusion_b200.sparse.basic import SparseTensor
    gen = torch.Generator().manual_seed(5)
    counts = (700, 1, 333, 1500)
    coords = torch.cat([torch.cat([torch.full((n, 1), b), torch.randint(0, 64, (n, 3), generator=gen)], 1)
                        for b, n in enumerate(counts)]).int().to(DEV)
    T, H, C = coords.shape[0], 4, 64
    qkv = (torch.randn(T, 3, H, C, generator=gen) * 0.7).half().to(DEV)
    out = sparse_scaled_dot_product_attention(SparseTensor(qkv, coords))
    ref, s0 = [], 0
    for n in counts:
        f = qkv[s0:s0 + n].float()
        ref.append(_sdpa(f[:, 0], f[:, 1], f[:, 2]))
        s0 += n
    _close(out.feats, torch.cat(ref))
    # voxels against a dense context (TRELLIS flow blocks' cross-attention), packed kv and separate k / v
    L = 77
    q = (torch.randn(T, H, C, generator=gen) * 0.7).half().to(DEV)
    kv = (torch.randn(len(counts), L, 2, H, C, generator=gen) * 0.7).half().to(DEV)
    o2 = sparse_scaled_dot_product_attention(SparseTensor(q, coords), kv)
    o3 = sparse_scaled_dot_product_attention(SparseTensor(q, coords), kv[:, :, 0], kv[:, :, 1])
    ref, s0 = [], 0
    for b, n in enumerate(counts):
        ref.append(_sdpa(q[s0:s0 + n].float(), kv[b, :, 0].float(), kv[b, :, 1].float()))
        s0 += n
    _close(o2.feats, torch.cat(ref))
    assert torch.equal(o2.feats, o3.feats)
    with pytest.raises(NotImplementedError):
        sparse_scaled_dot_product_attention(q[None], SparseTensor(qkv[:, :2], coords))
