"""GPU numerics of each sm_100a kernel (through the C ABI) against a plain PyTorch fp32
reference of the same op.  Tolerances: fp16 outputs -> 2e-3 relative to the tensor's max
(one fp16 ulp at the top of the range is 1e-3), stated per test."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _g(seed=0):
    return torch.Generator(device="cpu").manual_seed(seed)


def _rand(shape, g, scale=1.0):
    return (torch.randn(shape, generator=g) * scale).to(DEV)


def _close(a, b, tol, what=""):
    err = (a.float() - b.float()).abs().max().item()
    ref = b.float().abs().max().item()
    assert err <= tol * max(ref, 1e-6), f"{what}: max err {err:.3e} vs max |ref| {ref:.3e}"


@pytest.mark.parametrize("M,N,K", [(256, 128, 64), (1000, 512, 512), (12288, 1536, 512), (384, 2048, 512),
                                   (130, 72, 200), (12288, 512, 2048)])
def test_gemm_bias_f16(M, N, K):
    from gvfdiffusion_b200 import ops
    g = _g(M + N + K)
    a = _rand((M, K), g).half()
    w = _rand((N, K), g, 0.05).half()
    b = _rand((N,), g, 0.1)
    ref = a.float() @ w.float().T + b
    _close(ops.gemm(a, w, b, ops.EPI_F16), ref, 2e-3, "bias f16")
    _close(ops.gemm(a, w, None, ops.EPI_F32), a.float() @ w.float().T, 2e-4, "f32")
    _close(ops.gemm(a, w, b, ops.EPI_GELU_F16), F.gelu(ref.half().float(), approximate="tanh"), 2e-3, "gelu")


def test_gemm_residual_epilogues():
    from gvfdiffusion_b200 import ops
    g = _g(5)
    M, N, K, rpb = 768, 256, 128, 256
    a = _rand((M, K), g).half()
    w = _rand((N, K), g, 0.05).half()
    b = _rand((N,), g, 0.1)
    gate = _rand((3, 2 * N), g).half()
    x = _rand((M, N), g)
    lin = (a.float() @ w.float().T + b).half().float()
    ref = x + (lin * gate[:, N:].float().repeat_interleave(rpb, 0)).half().float()
    out = x.clone()
    ops.gemm(a, w, b, ops.EPI_RESID_F32, out=out, gate=gate[:, N:], gate_stride=2 * N, rows_per_batch=rpb)
    _close(out, ref, 1e-3, "gate+resid f32")
    out = x.clone()
    ops.gemm(a, w, b, ops.EPI_RESID_F32, out=out)
    _close(out, x + lin, 1e-3, "resid f32")
    xh = x.half()
    out = xh.clone()
    ops.gemm(a, w, b, ops.EPI_RESID_F16, out=out)
    _close(out, (lin + xh.float()).half(), 2e-3, "resid f16")


def _ref_attn(q, k, v, scale):
    qf, kf, vf = (t.float().permute(0, 2, 1, 3) for t in (q, k, v))
    s = (qf @ kf.transpose(-1, -2)) * scale
    return (torch.softmax(s, dim=-1) @ vf).permute(0, 2, 1, 3)


@pytest.mark.parametrize("Nb,Lq,Lk,H,D", [(2, 512, 512, 4, 32), (3, 512, 1370, 2, 32), (2, 256, 4096, 3, 32),
                                          (2, 100, 70, 2, 32), (5, 24, 24, 4, 32), (2, 512, 512, 3, 64),
                                          (2, 700, 512, 2, 64), (1, 128, 128, 1, 32), (1, 129, 257, 1, 64)])
def test_attention_matches_fp32(Nb, Lq, Lk, H, D):
    from gvfdiffusion_b200 import ops
    g = _g(Nb * Lq + Lk + D)
    q = _rand((Nb, Lq, H, D), g).half()
    k = _rand((Nb, Lk, H, D), g).half()
    v = _rand((Nb, Lk, H, D), g).half()
    scale = 1.0 / math.sqrt(D)
    o = ops.attention(q, k, v, scale)
    torch.cuda.synchronize()
    _close(o, _ref_attn(q, k, v, scale), 2e-3, "attention")


@pytest.mark.parametrize("dbg", [0x10, 0x30, 0x14, 0x81, 0x84, 0x88, 0x90])
def test_attention_kernel_generations(dbg):
    """every d = 32 kernel variant the dispatcher can pick (v4 plain / MUFU ping-pong / polynomial share, v6
    MUFU-only / polynomial share), on shapes with 1..4 query tiles, ragged key counts, and scores whose row
    maximum keeps growing along the keys so that v6's lazy rescale of O in TMEM runs many times"""
    from gvfdiffusion_b200 import _lib, ops
    L = _lib.lib()
    g = _g(dbg)
    scale = 1.0 / math.sqrt(32)
    try:
        L.gvf_attn_set_debug(dbg)
        for (Nb, Lq, Lk, H) in [(2, 512, 512, 3), (1, 1000, 1370, 2), (2, 300, 200, 2), (1, 130, 4096, 2), (1, 512, 70, 1)]:
            q = _rand((Nb, Lq, H, 32), g).half()
            k = _rand((Nb, Lk, H, 32), g).half()
            v = _rand((Nb, Lk, H, 32), g).half()
            _close(ops.attention(q, k, v, scale), _ref_attn(q, k, v, scale), 2e-3, f"variant {dbg:#x} {(Nb, Lq, Lk, H)}")
        # keys whose projection on q grows with the key index: the running maximum rises block after block
        Nb, Lq, Lk, H = 1, 512, 2048, 2
        q = (_rand((Nb, Lq, H, 32), g).abs() + 0.5).half()
        ramp = torch.linspace(-6, 6, Lk, device=DEV)[None, :, None, None]
        k = (ramp.expand(Nb, Lk, H, 32) + 0.1 * _rand((Nb, Lk, H, 32), g)).half()
        v = _rand((Nb, Lk, H, 32), g).half()
        _close(ops.attention(q, k, v, 1.0), _ref_attn(q, k, v, 1.0), 2e-3, f"variant {dbg:#x} growing max")
    finally:
        L.gvf_attn_set_debug(0)


def test_attention_packed_and_shared_views():
    from gvfdiffusion_b200 import ops
    g = _g(77)
    Nb, L, H, D = 3, 384, 4, 32
    qkv = _rand((Nb, L, 3, H, D), g).half()          # packed qkv, the DiT self-attention layout
    q, k, v = qkv.unbind(2)
    scale = 1.0 / math.sqrt(D)
    _close(ops.attention(q, k, v, scale), _ref_attn(q, k, v, scale), 2e-3, "packed qkv")
    # kv shared by all batch entries (static cross-attention), kv packed [L, 2, H, D]
    kv = _rand((640, 2, H, D), g).half()
    ks, vs = kv[:, 0], kv[:, 1]
    ref = _ref_attn(q, ks[None].expand(Nb, -1, -1, -1), vs[None].expand(Nb, -1, -1, -1), scale)
    _close(ops.attention(q, ks, vs, scale, kv_shared=True), ref, 2e-3, "kv shared")
    # q shared (VAE decoder queries), d = 64, kv chunked along the last dim like autoencoder.py:130
    Hh, Dd = 2, 64
    qs = _rand((300, Hh, Dd), g).half()
    kvc = _rand((Nb, 512, 2 * Hh * Dd), g).half()
    kk, vv = kvc[..., :Hh * Dd].unflatten(-1, (Hh, Dd)), kvc[..., Hh * Dd:].unflatten(-1, (Hh, Dd))
    ref = _ref_attn(qs[None].expand(Nb, -1, -1, -1), kk, vv, 0.125)
    _close(ops.attention(qs, kk, vv, 0.125, q_shared=True), ref, 2e-3, "q shared")
    # temporal view: sequences strided by N tokens (DiT temporal attention, T = 24)
    T, Nt = 24, 40
    x = _rand((T, Nt, 3, H, D), g).half()
    xt = x.permute(1, 0, 2, 3, 4)                    # [Nt, T, 3, H, D] strided view
    qt, kt, vt = xt.unbind(2)
    out = torch.empty((T, Nt, H, D), dtype=torch.float16, device=DEV)
    ops.attention(qt, kt, vt, scale, out=out.permute(1, 0, 2, 3))
    _close(out.permute(1, 0, 2, 3), _ref_attn(qt, kt, vt, scale), 2e-3, "temporal view")


@pytest.mark.parametrize("Lk,shared", [(4096, True), (2100, False), (1370, False)])
def test_attention_last_wave_key_split(Lk, shared):
    """384 (batch, head) units on 148 SMs: the units of the third wave are cut into three key ranges and merged
    (attn_merge_kernel).  Checked on the first and the last batch entries (whole / split units) against torch,
    and against the same launch with the scratch buffer unregistered (no split)."""
    from gvfdiffusion_b200 import _lib, ops
    g = _g(500 + Lk)
    Nb, Lq, H, D = 24, 512, 16, 32
    scale = 1.0 / math.sqrt(D)
    q = (_rand((Nb, Lq, H, D), g) * 1.5).half()
    if shared:
        k, v = (_rand((Lk, H, D), g) * 1.5).half(), _rand((Lk, H, D), g).half()
    else:
        k, v = (_rand((Nb, Lk, H, D), g) * 1.5).half(), _rand((Nb, Lk, H, D), g).half()
    out = ops.attention(q, k, v, scale, kv_shared=shared)
    for sl in (slice(0, 2), slice(Nb - 2, Nb)):
        kk = k[None].expand(2, -1, -1, -1) if shared else k[sl]
        vv = v[None].expand(2, -1, -1, -1) if shared else v[sl]
        _close(out[sl], _ref_attn(q[sl], kk, vv, scale), 2e-3, f"Lk {Lk} batches {sl}")
    L = _lib.lib()
    L.gvf_attn_set_workspace(None, 0)
    try:
        whole = ops.attention.__wrapped__(q, k, v, scale, kv_shared=shared) if hasattr(ops.attention, "__wrapped__") else None
        if whole is None:
            saved, ops._attn_ws = ops._attn_ws, torch.empty(1, device=DEV)     # keep _ensure_attn_ws from re-registering
            try:
                whole = ops.attention(q, k, v, scale, kv_shared=shared)
            finally:
                ops._attn_ws = saved
    finally:
        L.gvf_attn_set_workspace(_lib.ptr(ops._attn_ws), ops._attn_ws.numel() * 4)
    _close(out, whole.float(), 1e-3, f"split vs whole Lk {Lk}")


def test_attention_key_split_ragged_query_blocks():
    """Two query blocks per (batch, head), the second one partial (Lq = 1000), 320 units -> 24 split units."""
    from gvfdiffusion_b200 import ops
    g = _g(612)
    Nb, Lq, Lk, H, D = 10, 1000, 2100, 16, 32
    scale = 1.0 / math.sqrt(D)
    q = (_rand((Nb, Lq, H, D), g) * 1.5).half()
    k, v = (_rand((Nb, Lk, H, D), g) * 1.5).half(), _rand((Nb, Lk, H, D), g).half()
    out = ops.attention(q, k, v, scale)
    for sl in (slice(0, 1), slice(Nb - 1, Nb)):
        _close(out[sl], _ref_attn(q[sl], k[sl], v[sl], scale), 2e-3, f"batches {sl}")


@pytest.mark.parametrize("T,H", [(24, 16), (32, 8), (5, 8), (17, 16), (16, 8)])
def test_attention_short_sequences_mma(T, H):
    """L <= 32 with contiguous heads, H % 8 == 0: the warp-level mma.sync kernel (DiT temporal attention),
    on the strided (T, N, 3, H, d) -> (N, T, ...) view and on a plain contiguous batch; the CUDA-core kernel
    (debug switch 0x40) must agree with it."""
    from gvfdiffusion_b200 import _lib, ops
    g = _g(1000 + T + H)
    D, Nt = 32, 37
    scale = 1.0 / math.sqrt(D)
    x = (_rand((T, Nt, 3, H, D), g) * 1.5).half()
    qt, kt, vt = x.permute(1, 0, 2, 3, 4).unbind(2)
    out = torch.empty((T, Nt, H, D), dtype=torch.float16, device=DEV)
    ops.attention(qt, kt, vt, scale, out=out.permute(1, 0, 2, 3))
    ref = _ref_attn(qt, kt, vt, scale)
    _close(out.permute(1, 0, 2, 3), ref, 2e-3, "temporal view (mma)")
    qc, kc, vc = (t.contiguous() for t in (qt, kt, vt))
    _close(ops.attention(qc, kc, vc, scale), ref, 2e-3, "contiguous (mma)")
    L = _lib.lib()
    L.gvf_attn_set_debug(0x40)
    try:
        old = ops.attention(qc, kc, vc, scale)
    finally:
        L.gvf_attn_set_debug(0)
    _close(old, ref, 2e-3, "contiguous (cuda cores)")


@pytest.mark.parametrize("C,dt", [(512, torch.float32), (768, torch.float16), (64, torch.float32), (384, torch.float16)])
def test_ln_mod(C, dt):
    from gvfdiffusion_b200 import ops
    g = _g(C)
    M, rpb = 96, 32
    x = _rand((M, C), g).to(dt)
    mod = _rand((3, 2 * C), g, 0.3).half()
    w, b = _rand((C,), g), _rand((C,), g)
    ln = F.layer_norm(x.float(), (C,), None, None, 1e-6)
    ref = ln * (1 + mod[:, C:].float().repeat_interleave(rpb, 0)) + mod[:, :C].float().repeat_interleave(rpb, 0)
    _close(ops.ln_mod(x, shift=mod[:, :C], scale=mod[:, C:], mod_stride=2 * C, rows_per_batch=rpb), ref, 1e-3)
    _close(ops.ln_mod(x, w=w, b=b), ln * w + b, 1e-3)
    _close(ops.ln_mod(x, eps=1e-5), F.layer_norm(x.float(), (C,), None, None, 1e-5), 1e-3)


@pytest.mark.parametrize("rows,H,D", [(200, 4, 32), (203, 16, 64), (77, 2, 64)])
def test_rmsnorm_heads(rows, H, D):
    from gvfdiffusion_b200 import ops
    g = _g(9)
    qkv = _rand((rows, 3 * H * D), g).half()
    gq, gk = _rand((H, D), g) + 1, _rand((H, D), g) + 1
    q, k, v = qkv.float().reshape(rows, 3, H, D).unbind(1)
    rq = F.normalize(q, dim=-1) * gq * D ** 0.5
    rk = F.normalize(k, dim=-1) * gk * D ** 0.5
    buf = qkv.clone()
    ops.rmsnorm_heads_(buf, H, D, H * D, gq, gk)
    o = buf.float().reshape(rows, 3, H, D)
    _close(o[:, 0], rq, 1e-3)
    _close(o[:, 1], rk, 1e-3)
    assert torch.equal(o[:, 2], v)


def test_modulation_small_linear_ape_final():
    from gvfdiffusion_b200 import ops
    g = _g(21)
    B, C, Fq, R = 3, 512, 256, 1000
    t = torch.tensor([998.996, 500.25, -0.004], device=DEV)
    W0, W2, Wm = _rand((C, Fq), g, 0.02).half(), _rand((C, C), g, 0.02).half(), _rand((R, C), g, 0.05).half()
    b0, b2, bm = _rand((C,), g, 0.02), _rand((C,), g, 0.02), _rand((R,), g, 0.02)
    temb = torch.empty((B, C), dtype=torch.float16, device=DEV)
    st = torch.empty_like(temb)
    mod = torch.empty((B, R), dtype=torch.float16, device=DEV)
    ops.dit_modulation(t, W0, b0, W2, b2, Wm, bm, temb, st, mod)
    half = Fq // 2
    freqs = torch.exp(-math.log(10000) * torch.arange(half, device=DEV, dtype=torch.float32) / half)
    emb = torch.cat([torch.cos(t[:, None] * freqs), torch.sin(t[:, None] * freqs)], -1)
    te = F.linear(F.silu(F.linear(emb, W0.float(), b0)), W2.float(), b2)
    _close(temb, te, 3e-3, "t_emb")
    _close(mod, F.linear(F.silu(te), Wm.float(), bm), 4e-3, "modulation")
    # small linear + APE
    x = _rand((70, 16), g)
    W = _rand((512, 16), g, 0.2).half()
    bb = _rand((512,), g, 0.1)
    xyz = torch.rand((35, 3), generator=g).to(DEV) - 0.5
    pos = ops.ape(xyz, 512)
    fd = 512 // 3 // 2
    fr = 1.0 / (10000 ** (torch.arange(fd, device=DEV, dtype=torch.float32) / fd))
    o = torch.outer(xyz.reshape(-1), fr)
    pref = torch.cat([torch.sin(o), torch.cos(o)], -1).reshape(35, -1)
    pref = torch.cat([pref, torch.zeros(35, 512 - pref.shape[1], device=DEV)], -1)
    _close(pos, pref, 1e-5, "APE")
    y = ops.small_linear(x, W, bb, out_f16=False, add=pos, add_rows=35)
    _close(y, F.linear(x.half().float(), W.float(), bb) + pref.repeat(2, 1), 1e-3, "small linear + add")
    # final layer
    X = _rand((64, 512), g)
    modf = _rand((2, 1024), g, 0.3).half()
    Wf, bf = _rand((16, 512), g, 0.05).half(), _rand((16,), g, 0.1)
    ln = F.layer_norm(X, (512,), None, None, 1e-6)
    ref = F.linear(ln * (1 + modf[:, 512:].float().repeat_interleave(32, 0)) + modf[:, :512].float().repeat_interleave(32, 0),
                   Wf.float(), bf)
    _close(ops.dit_final_layer(X, modf[:, :512], modf[:, 512:], 1024, 32, Wf, bf), ref, 2e-3, "final layer")


def test_small_linear_and_final_layer_large_m():
    """The 64-rows-per-block small linear and the 4-rows-per-warp final layer at the benchmark row count and at ragged
    row counts, against torch and against the small-M code path (same per-output arithmetic: identical bits)."""
    from gvfdiffusion_b200 import ops
    g = _g(77)
    for M, add_rows in ((12288, 512), (4100, 205), (4097, 4097)):
        x = _rand((M, 16), g)
        W = _rand((512, 16), g, 0.2).half()
        bb = _rand((512,), g, 0.1)
        pos = _rand((add_rows, 512), g)
        y = ops.small_linear(x, W, bb, out_f16=False, add=pos, add_rows=add_rows)
        ref = F.linear(x.half().float(), W.float(), bb) + pos.repeat((M + add_rows - 1) // add_rows, 1)[:M]
        _close(y, ref, 1e-3, f"small linear M={M}")
        # the small-M kernel on row slices that keep the add phase: identical bits
        y_small = torch.cat([ops.small_linear(x[i:i + add_rows], W, bb, out_f16=False, add=pos, add_rows=add_rows)
                             for i in range(0, M, add_rows)], 0) if add_rows < 4096 else None
        if y_small is not None:
            assert torch.equal(y, y_small)
        y16 = ops.small_linear(x, W, bb, out_f16=True)
        _close(y16, F.linear(x.half().float(), W.float(), bb), 2e-3, f"small linear fp16 M={M}")
    for M, rpb in ((12288, 12288), (1003, 1003), (96, 32)):
        X = _rand((M, 512), g)
        nb = M // rpb
        modf = _rand((nb, 1024), g, 0.3).half()
        Wf, bf = _rand((16, 512), g, 0.05).half(), _rand((16,), g, 0.1)
        ln = F.layer_norm(X, (512,), None, None, 1e-6)
        ref = F.linear(ln * (1 + modf[:, 512:].float().repeat_interleave(rpb, 0)) + modf[:, :512].float().repeat_interleave(rpb, 0),
                       Wf.float(), bf)
        _close(ops.dit_final_layer(X, modf[:, :512], modf[:, 512:], 1024, rpb, Wf, bf), ref, 2e-3, f"final layer M={M}")


def test_ln_two_rows_per_warp_is_bit_identical():
    """ln_mod2_kernel (two rows per warp in flight: widths 512 / 768 / 1024, >= 2048 rows) against the one-row kernel."""
    from gvfdiffusion_b200 import _lib, ops
    g = _g(5)
    try:
        for M, C, rpb in ((12288, 512, 12288), (2049, 768, 683), (4097, 1024, 4097), (6144, 512, 512)):
            nb = (M + rpb - 1) // rpb
            for dt in (torch.float32, torch.float16):
                x = _rand((M, C), g).to(dt)
                w, b = _rand((C,), g) + 1, _rand((C,), g)
                mod = _rand((nb, 2 * C), g, 0.3).half()
                outs = []
                for two in (1, 0):
                    _lib.lib().gvf_ln_set_two_rows(two)
                    outs.append((ops.ln_mod(x), ops.ln_mod(x, w=w, b=b),
                                 ops.ln_mod(x, shift=mod[:, :C], scale=mod[:, C:], mod_stride=2 * C, rows_per_batch=rpb)))
                for a, r in zip(*outs):
                    assert torch.equal(a, r), (M, C, dt)
                _close(outs[0][1], F.layer_norm(x.float(), (C,), w, b, 1e-6), 1e-3)
    finally:
        _lib.lib().gvf_ln_set_two_rows(0)


def test_geglu_cast_dpm():
    from gvfdiffusion_b200 import ops
    g = _g(31)
    h = _rand((50, 256), g).half()
    a, gt = h.float().chunk(2, -1)
    _close(ops.geglu(h), a * F.gelu(gt), 2e-3, "geglu")
    x = _rand((1000,), g)
    assert torch.equal(ops.cast_f16(x), x.half())
    n = 3000
    xs, v = _rand((n,), g), _rand((3, n), g)
    al, sg, s1, s2 = 0.8, 0.6, 2.0, 1.5
    eps = [al * v[i] + sg * xs for i in range(3)]
    e = eps[0] + s1 * (eps[1] - eps[0]) + s2 * (eps[2] - eps[1])
    out = torch.empty_like(xs)
    _close(ops.dpm_x0(xs, v, 3, al, sg, s1, s2, out), (xs - sg * e) / al, 1e-5, "x0 cfg")
    _close(ops.dpm_x0(xs, v[2], 1, al, sg, 1.0, 1.0, out), (xs - sg * eps[2]) / al, 1e-5, "x0")
    m0, m1 = _rand((n,), g), _rand((n,), g)
    _close(ops.dpm_update(xs, m0, m1, 0.9, -0.3, 1.2, 2, torch.empty_like(xs)),
           0.9 * xs + 0.3 * m0 + 0.5 * 0.3 * (1.2 * (m0 - m1)), 1e-5, "update2")
    _close(ops.dpm_update(xs, m0, None, 0.9, -0.3, 0.0, 1, torch.empty_like(xs)), 0.9 * xs + 0.3 * m0, 1e-5, "update1")


def test_gemm_qkv_rmsnorm_fused():
    from gvfdiffusion_b200 import ops
    g = _g(41)
    M, C, H, D = 640, 256, 8, 32
    a = _rand((M, C), g).half()
    w = _rand((3 * C, C), g, 0.05).half()
    b = _rand((3 * C,), g, 0.1)
    gq, gk = _rand((H, D), g) + 1, _rand((H, D), g) + 1
    out = torch.empty((M, 3 * C), dtype=torch.float16, device=DEV)
    ops.gemm_qkv_rmsnorm(a, w, b, gq, gk, out)
    lin = (a.float() @ w.float().T + b).half().float().reshape(M, 3, H, D)
    rq = F.normalize(lin[:, 0], dim=-1) * gq * D ** 0.5
    rk = F.normalize(lin[:, 1], dim=-1) * gk * D ** 0.5
    o = out.float().reshape(M, 3, H, D)
    _close(o[:, 0], rq, 2e-3, "q")
    _close(o[:, 1], rk, 2e-3, "k")
    _close(o[:, 2], lin[:, 2], 1e-3, "v")


def test_gemm_ragged_rows_and_multi_batch_gate():
    from gvfdiffusion_b200 import ops
    g = _g(43)
    M, N, K, rpb = 200, 136, 72, 50          # tiles span several batches, ragged M / N / K
    a = _rand((M, K), g).half()
    w = _rand((N, K), g, 0.05).half()
    b = _rand((N,), g, 0.1)
    gate = _rand((4, N), g).half()
    x = _rand((M, N), g)
    lin = (a.float() @ w.float().T + b).half().float()
    out = x.clone()
    ops.gemm(a, w, b, ops.EPI_RESID_F32, out=out, gate=gate, gate_stride=N, rows_per_batch=rpb)
    _close(out, x + (lin * gate.float().repeat_interleave(rpb, 0)).half().float(), 1e-3, "ragged gate")
    o14 = torch.empty((M, 14), dtype=torch.float32, device=DEV)
    w16 = torch.cat([w[:14], torch.zeros(2, K, dtype=torch.float16, device=DEV)])
    ops.gemm(a, w16, torch.cat([b[:14], torch.zeros(2, device=DEV)]), ops.EPI_F32_COMPACT, out=o14)
    _close(o14, lin[:, :14], 1e-3, "compact")


@pytest.mark.parametrize("M,K,rpb,kind", [(12288, 512, 12288, "mod"), (300, 2048, 100, "affine"), (1000, 136, 250, "mod"),
                                          (257, 512, 257, "plain")])
def test_gemm_resid_ln_fused(M, K, rpb, kind):
    """Residual Linear + LayerNorm(+modulate / affine) in one kernel against the two-kernel path and torch."""
    from gvfdiffusion_b200 import ops
    g = _g(300 + M + K)
    N = 512
    a = _rand((M, K), g).half()
    w = _rand((N, K), g, 0.05).half()
    b = _rand((N,), g, 0.1)
    nb = (M + rpb - 1) // rpb
    gate = _rand((nb, N), g).half() if kind != "plain" else None
    x0 = _rand((M, N), g) * 2 + 0.3
    mod = _rand((nb, 2 * N), g, 0.3).half()
    lw, lb = _rand((N,), g) + 1, _rand((N,), g)
    kw = {}
    if kind == "mod":
        kw = dict(shift=mod[:, :N], scale=mod[:, N:], mod_stride=2 * N)
    elif kind == "affine":
        kw = dict(ln_w=lw, ln_b=lb)
    # fused
    x1 = x0.clone()
    y1 = torch.empty((M, N), dtype=torch.float16, device=DEV)
    ops.gemm_resid_ln(a, w, b, x1, y1, gate=gate, gate_stride=N, rows_per_batch=rpb, **kw)
    # two kernels
    x2 = x0.clone()
    ops.gemm(a, w, b, ops.EPI_RESID_F32, out=x2, gate=gate, gate_stride=N, rows_per_batch=rpb)
    if kind == "mod":
        y2 = ops.ln_mod(x2, shift=mod[:, :N], scale=mod[:, N:], mod_stride=2 * N, rows_per_batch=rpb)
    elif kind == "affine":
        y2 = ops.ln_mod(x2, w=lw, b=lb)
    else:
        y2 = ops.ln_mod(x2)
    assert torch.equal(x1, x2), f"residual stream differs: max {float((x1 - x2).abs().max())}"
    _close(y1, y2.float(), 1e-3, "fused LN vs ln_mod kernel")
    # torch
    lin = (a.float() @ w.float().T + b).half().float()
    if gate is not None:
        lin = (lin * gate.float().repeat_interleave(rpb, 0)[:M]).half().float()
    xr = x0 + lin
    ln = F.layer_norm(xr, (N,), None, None, 1e-6)
    if kind == "mod":
        ln = ln * (1 + mod[:, N:].float().repeat_interleave(rpb, 0)[:M]) + mod[:, :N].float().repeat_interleave(rpb, 0)[:M]
    elif kind == "affine":
        ln = ln * lw + lb
    _close(x1, xr, 1e-3, "x vs torch")
    _close(y1, ln, 2e-3, "y vs torch")


@pytest.mark.parametrize("variant", [4, 5, 6, 7])
def test_gemm_tma_epilogue_all_modes(variant):
    """Generation-2 kernels (eight epilogue warps, TMA stores): every fused epilogue, ragged M / N / K,
    gates spanning several batches inside one tile."""
    from gvfdiffusion_b200 import _lib, ops
    L = _lib.lib()
    g = _g(200 + variant)
    try:
        L.gvf_gemm_set_variant(variant)
        for (M, N, K, rpb) in [(12288, 512, 512, 12288), (1000, 1536, 512, 250), (300, 264, 136, 50), (640, 768, 3072, 640)]:
            a = _rand((M, K), g).half()
            w = _rand((N, K), g, 0.05).half()
            b = _rand((N,), g, 0.1)
            lin = a.float() @ w.float().T + b
            lin16 = lin.half().float()
            tag = f"variant {variant} {M}x{N}x{K}"
            _close(ops.gemm(a, w, b, ops.EPI_F16), lin, 2e-3, tag + " f16")
            _close(ops.gemm(a, w, b, ops.EPI_F32), lin, 1e-3, tag + " f32")
            _close(ops.gemm(a, w, b, ops.EPI_GELU_F16), F.gelu(lin16, approximate="tanh"), 2e-3, tag + " gelu")
            x = _rand((M, N), g)
            out = x.clone()
            ops.gemm(a, w, b, ops.EPI_RESID_F32, out=out)
            _close(out, x + lin16, 1e-3, tag + " resid")
            nb = (M + rpb - 1) // rpb
            gate = _rand((nb, N), g).half()
            out = x.clone()
            ops.gemm(a, w, b, ops.EPI_RESID_F32, out=out, gate=gate, gate_stride=N, rows_per_batch=rpb)
            _close(out, x + (lin16 * gate.float().repeat_interleave(rpb, 0)[:M]).half().float(), 1e-3, tag + " gate")
            x16 = _rand((M, N), g).half()
            out16 = x16.clone()
            ops.gemm(a, w, b, ops.EPI_RESID_F16, out=out16)
            _close(out16, x16.float() + lin16, 2e-3, tag + " resid16")
        # fused q/k RMS norm (DiT head layout)
        M, C, H, D = 700, 256, 8, 32
        a = _rand((M, C), g).half()
        w = _rand((3 * C, C), g, 0.05).half()
        b = _rand((3 * C,), g, 0.1)
        gq, gk = _rand((H, D), g) + 1, _rand((H, D), g) + 1
        out = torch.empty((M, 3 * C), dtype=torch.float16, device=DEV)
        ops.gemm_qkv_rmsnorm(a, w, b, gq, gk, out)
        lin = (a.float() @ w.float().T + b).half().float().reshape(M, 3, H, D)
        o = out.float().reshape(M, 3, H, D)
        _close(o[:, 0], F.normalize(lin[:, 0], dim=-1) * gq * D ** 0.5, 2e-3, "q")
        _close(o[:, 1], F.normalize(lin[:, 1], dim=-1) * gk * D ** 0.5, 2e-3, "k")
        _close(o[:, 2], lin[:, 2], 1e-3, "v")
    finally:
        L.gvf_gemm_set_variant(-1)


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
def test_gemm_scheduling_variants(variant):
    from gvfdiffusion_b200 import _lib, ops
    L = _lib.lib()
    g = _g(100 + variant)
    try:
        L.gvf_gemm_set_variant(variant)
        for (M, N, K) in [(12288, 512, 512), (1000, 1536, 512), (300, 264, 136), (12288, 2048, 512)]:
            a = _rand((M, K), g).half()
            w = _rand((N, K), g, 0.05).half()
            b = _rand((N,), g, 0.1)
            ref = a.float() @ w.float().T + b
            _close(ops.gemm(a, w, b, ops.EPI_F16), ref, 2e-3, f"variant {variant} {M}x{N}x{K}")
            x = _rand((M, N), g)
            out = x.clone()
            ops.gemm(a, w, b, ops.EPI_RESID_F32, out=out)
            _close(out, x + ref.half().float(), 1e-3, f"variant {variant} resid {M}x{N}x{K}")
    finally:
        L.gvf_gemm_set_variant(-1)


@pytest.mark.parametrize("M,F_,K", [(12288, 3072, 768), (1000, 384, 96), (130, 128, 64), (24576, 3072, 768)])
def test_gemm_geglu_fused_is_bit_identical_to_the_two_kernels(M, F_, K):
    """FeedForward.net[0] + GEGLU (reference model/autoencoder.py:90-107) in one kernel: interleaved weight rows, value
    and gate meet in one accumulator tile.  Same rounding points as gvf_gemm_f16 + gvf_geglu_f16 -> identical bits;
    and within fp16 rounding of the fp32 torch expression."""
    from gvfdiffusion_b200 import ops
    g = _g(M + F_)
    a = _rand((M, K), g).half()
    w = _rand((2 * F_, K), g, 0.05).half()
    b = _rand((2 * F_,), g, 0.1)
    two = ops.geglu(ops.gemm(a, w, b, ops.EPI_F16))
    wi, bi = ops.geglu_interleave(w, b)
    one = ops.gemm_geglu(a, wi, bi)
    assert torch.equal(one, two)
    h = (a.float() @ w.float().T + b).half().float()
    ref = h[:, :F_] * F.gelu(h[:, F_:])
    _close(one, ref, 2e-3, "geglu fused")


def test_lpips_cuda_path_matches_torch_formulation():
    """LPIPS on the device (fp16 channels-last VGG16 through cuDNN, one batch for both images, the tail fused into
    gvf_lpips_tap_fwd / _bwd) against the module's own plain-torch fp32 formulation run on the same device: loss and the
    gradient of the prediction.  fp16 convolutions: 2e-2."""
    from gvfdiffusion_b200.utils.lpips import LPIPS
    m = LPIPS(seed=3).to(DEV).eval()
    g = torch.Generator().manual_seed(5)
    x = (torch.rand(3, 3, 128, 96, generator=g) * 2 - 1).to(DEV).requires_grad_(True)
    y = (torch.rand(3, 3, 128, 96, generator=g) * 2 - 1).to(DEV)
    loss = m(x, y)
    (loss * 65536.0).backward()                        # fp16 activation gradients need the training step's loss scale
    xr = x.detach().clone().requires_grad_(True)
    fx = m.taps(xr)
    with torch.no_grad():
        fy = m.taps(y)
    ref = torch.sum(torch.cat([l((m._unit(a) - m._unit(b)) ** 2).mean((2, 3), True) for a, b, l in zip(fx, fy, m.lin)], 0)) / 3
    (ref * 65536.0).backward()
    assert abs(float(loss.detach()) - float(ref.detach())) < 2e-2 * abs(float(ref.detach())), (float(loss.detach()), float(ref.detach()))
    rel = float((x.grad - xr.grad).norm() / xr.grad.norm())
    assert rel < 5e-2, rel
    # the fused tail alone, on identical fp16 activations: tight
    from gvfdiffusion_b200.utils.lpips.lpips import _TapFn
    f = (torch.randn(4, 128, 20, 24, generator=g) * 0.7).half().to(DEV).contiguous(memory_format=torch.channels_last)
    a, b = f[:2].clone().requires_grad_(True), f[2:]
    w = torch.rand(128, generator=g).to(DEV)
    d = _TapFn.apply(a, b, w)
    gw = torch.tensor([0.7, -1.3], device=DEV)
    (d * gw).sum().backward()
    a32 = a.detach().float().requires_grad_(True)
    e = (m._unit(a32) - m._unit(b.float())) ** 2
    dr = (e * w[None, :, None, None]).sum(1).mean((1, 2))
    (dr * gw).sum().backward()
    assert torch.allclose(d, dr.detach(), rtol=1e-4, atol=1e-7)
    assert float((a.grad.float() - a32.grad).norm() / a32.grad.norm()) < 2e-3
