"""Host partition of the serialized sparse attention (gvfdiffusion_b200/sparse/attention/serialized_attn.py) against
the reference's own `calc_serialization` (sparse/attention/serialized_attn.py:38-119) recorded by
tests/golden/make_golden.py::gen_serialization (curve codes from the reference's PyTorch vox2seq twin).  Runs on CPU
tensors with the oracle's numpy curve encoder -- no device code involved."""
import os
import types

import numpy as np
import torch

from gvfdiffusion_b200.sparse.attention.serialized_attn import SerializeMode, calc_serialization
from oracle import vox2seq as OV

G = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "serialization.pt"), weights_only=False)


def _enc(coords, permute, mode):
    return torch.from_numpy(OV.encode(coords.numpy(), tuple(permute), mode).astype(np.int64))


def _tensor(case):
    layout, off = [], 0
    for n in case["counts"]:
        layout.append(slice(off, off + n))
        off += n
    return types.SimpleNamespace(coords=case["coords"], layout=layout, device=case["coords"].device)


def test_calc_serialization_matches_reference():
    assert len(G) >= 6
    for case in G:
        fwd, bwd, seq_lens, seq_batch = calc_serialization(_tensor(case), case["window"], SerializeMode[case["mode"]],
                                                           case["shift_sequence"], case["shift_window"], encode=_enc)
        assert torch.equal(fwd, case["fwd"]), case["mode"]
        assert torch.equal(bwd, case["bwd"]), case["mode"]
        assert list(seq_lens) == case["seq_lens"] and list(seq_batch) == case["seq_batch_indices"]
        # every voxel is kept exactly once, from a position that holds it
        assert torch.equal(fwd[bwd], torch.arange(case["coords"].shape[0]))
