"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every
symbol include/gvf_b200.h declares; argument validation works without touching the device; the host
mirrors load reference state dicts; the product never imports the oracle."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from gvfdiffusion_b200 import _lib
    L = _lib.lib()
    hdr = open(os.path.join(ROOT, "include", "gvf_b200.h")).read()
    declared = set(re.findall(r"GVF_API[^;(]*?\b(gvf_\w+)\s*\(", hdr))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(L, name), f"{name} declared in gvf_b200.h but not exported"
    assert declared == set(_lib.declared_symbols()), declared ^ set(_lib.declared_symbols())
    assert L.gvf_abi_version() == 1
    assert L.gvf_status_string(-2) == b"workspace too small"


def test_argument_validation_without_device():
    from gvfdiffusion_b200 import _lib
    L = _lib.lib()
    assert L.gvf_raster_workspace_bytes(0, 1, 1, 1, 1) == 0
    n = L.gvf_raster_workspace_bytes(24, 16384, 512, 512, 24 * 16384 * 8)
    assert 50e6 < n < 400e6
    offs = [L.gvf_raster_workspace_offset(i, 24, 16384, 512, 512, 1000) for i in range(10)]
    assert offs == sorted(offs) and all(o % 256 == 0 for o in offs)
    prm = _lib.RasterParams()
    assert L.gvf_raster_forward(C.byref(prm), 1, 1, 0, None, None, None, None, None, None, None, None, None, None,
                                None, 0, 1, None) == -1                       # null pointers -> GVF_ERR_INVALID
    assert L.gvf_gemm_f16(None, 0, None, 0, 1, 8, 8, 0, None, None, 0, None, 0, 0, None) == -1
    assert L.gvf_attn_fwd_f16(None, None, None, None, 1, 1, 1, 1, 32, None, None, None, None, 0, 0, 1.0, None) == -1


def test_host_mirrors_load_reference_state_dicts():
    from gvfdiffusion_b200.model.autoencoder import GSKLTemporalVariationalAutoEncoder
    from gvfdiffusion_b200.model.dit import DiT
    g = torch.load(os.path.join(ROOT, "tests", "golden", "dit_tiny.pt"), weights_only=False)
    m = DiT(**g["cfg"])
    assert set(m.state_dict()) == set(g["state_dict"])
    m.load_state_dict(g["state_dict"])
    with pytest.raises(RuntimeError):
        m(g["x"], g["t"], g["cond_images"], g["static_latent"], g["deformation_position_xyz"])   # CPU: no fallback
    v = torch.load(os.path.join(ROOT, "tests", "golden", "vae_tiny.pt"), weights_only=False)
    vv = GSKLTemporalVariationalAutoEncoder(**v["cfg"])
    vv.load_state_dict(v["state_dict"])
    with pytest.raises(NotImplementedError):
        DiT(**{**g["cfg"], "pe_mode": "learnable"})


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gvfdiffusion_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"
                assert "oracle/" not in src or f.endswith((".cu", ".h", ".cuh")) and "#include" not in [
                    l for l in src.splitlines() if "oracle/" in l][0]


def test_schedule_host_mirror_matches_oracle_on_cpu():
    from gvfdiffusion_b200.model import dpmsolver as D
    from oracle import dpm as O
    b = O.reference_betas(1000)
    ns, ons = D.NoiseScheduleVP("discrete", betas=torch.from_numpy(b)), O.NoiseScheduleVP(b)
    assert ns.total_N == ons.total_N == 996
    for t in torch.linspace(1.0, 1e-3, 33).tolist():
        tt = torch.tensor([t])
        assert abs(float(ns.marginal_lambda(t)) - float(ons.marginal_lambda(tt))) < 2e-5
        assert abs(float(ns.marginal_alpha(t)) - float(ons.marginal_alpha(tt))) < 1e-6
    fn = D.model_wrapper(lambda *a, **k: None, ns, model_type="v", guidance_type="classifier-free",
                         condition={"static_latent": torch.zeros(1)}, unconditional_condition=None)
    assert abs(float(fn.t_input(1.0)) - (1.0 - 1 / 996) * 1000) < 1e-3 and abs(float(fn.t_input(1e-3)) + 0.004) < 1e-3


def test_header_is_plain_c_and_cpp():
    """include/gvf_b200.h is the drop-in boundary: it must compile as C99 and as C++17 without torch or CUDA headers."""
    import shutil
    import subprocess
    hdr = os.path.join(ROOT, "include", "gvf_b200.h")
    if not shutil.which("gcc"):
        pytest.skip("gcc not available")
    for cmd in (["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr],
                ["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", hdr]):
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    txt = open(hdr).read()
    assert "#include <torch" not in txt and "#include <cuda" not in txt and "at::Tensor" not in txt
