"""Structural check of the built library (no GPU needed): the hot kernels are Blackwell-native -- tcgen05 tensor-core
MMAs with TMEM accumulators fed by TMA -- not recompiled mma.sync code.  SASS mnemonics per the profiling recipe:
tcgen05.mma -> UTC*MMA, tcgen05.ld / st -> LDTM / STTM, cp.async.bulk.tensor -> UTMALDG / UTMASTG; the legacy
mma.sync path shows as HMMA and is allowed only in the two small-tile kernels that use it on purpose."""
import os
import re
import shutil
import subprocess

import pytest

from gvfdiffusion_b200 import _lib

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
LEGACY = re.compile(r"(?<![A-Z])HMMA")      # mma.sync; not the HMMA inside UTCHMMA


@pytest.fixture(scope="module")
def sass_by_kernel():
    if not os.path.exists(CUOBJDUMP) or not os.path.exists(_lib.LIB_PATH):
        pytest.skip("cuobjdump or the built library is not available")
    txt = subprocess.run([CUOBJDUMP, "-sass", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    out, cur = {}, None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = []
        elif cur is not None:
            out[cur].append(line)
    return {k: "\n".join(v) for k, v in out.items()}


def _kernels(sass, needle):
    return {k: v for k, v in sass.items() if needle in k}


def test_gemm_and_attention_use_tcgen05_tmem_and_tma(sass_by_kernel):
    for needle in ("gemm_ws_kernel", "attn_fwd6_kernel", "gemm_ln_kernel"):
        ks = _kernels(sass_by_kernel, needle)
        assert ks, needle
        for name, body in ks.items():
            assert re.search(r"UTC\w*MMA", body), f"{name}: no tcgen05.mma"
            assert "UTMALDG" in body, f"{name}: operands are not loaded by TMA"
            assert "LDTM" in body, f"{name}: accumulators are not read from TMEM"
            assert not LEGACY.search(body), f"{name}: legacy mma.sync in a tcgen05 kernel"
    # generation-2 GEMMs leave the SM through TMA stores
    assert all("UTMASTG" in b for b in _kernels(sass_by_kernel, "gemm_ws_kernel").values())
    # the fp32-residual epilogue (MODE 2) brings the residual tile in by TMA as well: more than the two operand loads
    resid = [b for k, b in sass_by_kernel.items() if "gemm_ws_kernelILi128ELi4ELi2ELi1" in k]
    assert resid and resid[0].count("UTMALDG") >= 3


def test_legacy_mma_only_where_intended(sass_by_kernel):
    with_hmma = {k for k, b in sass_by_kernel.items() if LEGACY.search(b)}
    assert with_hmma, "the small-tile kernels are expected to use mma.sync"
    for k in with_hmma:
        assert "attn_small_mma_" in k or "sparse_window_attn_kernel" in k or "sparse_attn_bwd_" in k, k


def test_rasteriser_has_no_tensor_core_instructions(sass_by_kernel):
    for needle in ("sort_blend_kernel", "preprocess_kernel", "blend_backward2_kernel", "scatter_kernel"):
        for name, body in _kernels(sass_by_kernel, needle).items():
            assert not re.search(r"UTC\w*MMA", body) and not LEGACY.search(body), name
    assert "MUFU.EX2" in next(iter(_kernels(sass_by_kernel, "sort_blend_kernel").values()))


def test_round2_kernels_are_blackwell_native(sass_by_kernel):
    """attention v8 and the training-step kernels: tcgen05 + TMEM + TMA; the split-K epilogue reduces through the TMA unit
    (UTMAREDG), the sparse-convolution GEMM gathers its A rows with tile::gather4 (UTMALDG ... GATHER4 / .G4), the attention
    backward reads its LSE / D rows with bulk copies (UBLKCP)."""
    for needle in ("attn_fwd8_kernel", "attn_bwd_dkdv_kernel", "attn_bwd_dq_kernel"):
        ks = {k: b for k, b in _kernels(sass_by_kernel, needle).items() if "sparse_attn_bwd" not in k}   # mma.sync, below
        assert ks, needle
        for name, body in ks.items():
            assert re.search(r"UTC\w*MMA", body) and "UTMALDG" in body and "LDTM" in body and "STTM" in body, name
            assert not LEGACY.search(body), name
    assert any("UBLKCP" in b for b in _kernels(sass_by_kernel, "attn_bwd_dq_kernel").values())
    # fp32-store epilogue (MODE 4): TMA reduce-add for the split-K partial tiles
    m4 = {k: b for k, b in _kernels(sass_by_kernel, "gemm_ws_kernel").items() if re.search(r"ELi4ELi[12]ELb", k)}
    assert m4 and all("UTMAREDG" in b for b in m4.values()), list(m4)[:2]
    # MN-major (TRANS) and gather instantiations exist and are tcgen05 kernels
    # template tail: <..., TRANS, GATHER, TRANSB>
    ws = _kernels(sass_by_kernel, "gemm_ws_kernel")
    trans = [k for k in ws if "ELb1ELb0ELb0EEEv" in k]
    gath = {k: b for k, b in ws.items() if "ELb0ELb1ELb0EEEv" in k}
    transb = {k: b for k, b in ws.items() if "ELb0ELb0ELb1EEEv" in k}
    assert trans and gath and transb
    assert all(re.search(r"UTC\w*MMA", b) and "UTMALDG" in b for b in transb.values())
    assert all(re.search(r"UTMALDG\S*(GATHER4|G4)", b) or "GATHER" in b for b in gath.values()), "no gather4 TMA load in the sparse-conv GEMM"


def test_packed_window_attention_has_a_tma_gather_variant(sass_by_kernel):
    """sparse_window_attn_kernel<PACKED, TMA = true>: rows staged by `cp.async.bulk.tensor ... tile::gather4` (UTMALDG) on an
    mbarrier; the cp.async variant (the measured default) has none."""
    ks = _kernels(sass_by_kernel, "sparse_window_attn_kernel")
    tma = [b for k, b in ks.items() if "Lb1ELb1E" in k]
    plain = [b for k, b in ks.items() if "Lb1ELb0E" in k]
    assert tma and plain
    assert all("UTMALDG" in b and "SYNCS" in b for b in tma)
    assert all("UTMALDG" not in b and "LDGSTS" in b for b in plain)
