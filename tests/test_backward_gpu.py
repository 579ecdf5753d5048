"""Backward kernels of the motion-VAE training step (csrc/attn_bwd.cu, csrc/backward.cu) through the C ABI against
torch autograd of the same fp32 expression.  Tolerances: gradients are fp16 tensors (activation grads) -> relative L2
<= 5e-3 of the fp32 autograd gradient; fp32 reductions (bias / skinny weight grads) <= 1e-4."""
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


def _rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12))


def _ref_attn(q, k, v, scale):
    # q [Nb,Lq,H,D] ... -> [Nb,Lq,H,D], fp32
    s = torch.einsum("blhd,bmhd->bhlm", q, k) * scale
    return torch.einsum("bhlm,bmhd->blhd", s.softmax(-1), v)


@pytest.mark.parametrize("Nb,Lq,Lk,H,D", [(4, 512, 512, 3, 64), (3, 32, 32, 3, 32), (2, 300, 200, 2, 64),
                                           (2, 130, 520, 2, 32), (24, 512, 512, 12, 64)])
def test_attention_backward_matches_autograd(Nb, Lq, Lk, H, D):
    from gvfdiffusion_b200 import ops
    g = _g(Nb * Lq + Lk + D)
    q, k, v = (_rand((Nb, L, H, D), g, 0.8).half() for L in (Lq, Lk, Lk))
    do = _rand((Nb, Lq, H, D), g, 0.5).half()
    scale = D ** -0.5
    o, lse = ops.attention_fwd_lse(q, k, v, scale)
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    ref = _ref_attn(qf, kf, vf, scale)
    ref.backward(do.float())
    assert _rel(o, ref) < 2e-3
    # LSE2 rows: log2 sum exp(scale s)
    s = torch.einsum("blhd,bmhd->bhlm", q.float(), k.float()) * scale
    lse_ref = torch.logsumexp(s, -1) / math.log(2.0)
    assert torch.allclose(lse[:, :, :Lq], lse_ref, atol=2e-3, rtol=1e-4)
    assert torch.isinf(lse[:, :, Lq:]).all()
    dq, dk, dv = (torch.empty_like(t) for t in (q, k, v))
    ops.attention_bwd(q, k, v, o, do, lse, scale, dq, dk, dv)
    errs = (_rel(dq, qf.grad), _rel(dk, kf.grad), _rel(dv, vf.grad))
    print(f"attn bwd Nb={Nb} Lq={Lq} Lk={Lk} H={H} D={D}: rel L2 dq {errs[0]:.2e} dk {errs[1]:.2e} dv {errs[2]:.2e}")
    assert max(errs) < 5e-3, errs


@pytest.mark.parametrize("T,n,L,H,D", [(3, 300, 32, 3, 32), (4, 1000, 512, 3, 64), (24, 2048, 512, 12, 64)])
def test_attention_backward_shared_queries_packed_kv(T, n, L, H, D):
    """Decoder cross-attention of the motion VAE: q [n,H,D] shared by the T frames, k / v halves of one packed
    [T, L, 2, H, D] tensor; dq sums over frames, dk / dv land in the packed gradient tensor."""
    from gvfdiffusion_b200 import ops
    g = _g(T + n)
    q = _rand((n, H, D), g, 0.8).half()
    kv = _rand((T, L, 2, H, D), g, 0.8).half()
    do = _rand((T, n, H, D), g, 0.5).half()
    scale = D ** -0.5
    o, lse = ops.attention_fwd_lse(q, kv[:, :, 0], kv[:, :, 1], scale, q_shared=True)
    qf, kvf = q.float().requires_grad_(True), kv.float().requires_grad_(True)
    ref = _ref_attn(qf[None].expand(T, n, H, D), kvf[:, :, 0], kvf[:, :, 1], scale)
    ref.backward(do.float())
    assert _rel(o, ref) < 2e-3
    dq, dkv = torch.empty_like(q), torch.empty_like(kv)
    ops.attention_bwd(q, kv[:, :, 0], kv[:, :, 1], o, do, lse, scale, dq, dkv[:, :, 0], dkv[:, :, 1], q_shared=True)
    errs = (_rel(dq, qf.grad), _rel(dkv, kvf.grad))
    print(f"attn bwd shared-q T={T} n={n} L={L} D={D}: rel L2 dq {errs[0]:.2e} dkv {errs[1]:.2e}")
    assert max(errs) < 5e-3, errs


@pytest.mark.parametrize("M,C,dt", [(1000, 768, torch.float16), (96, 96, torch.float16), (520, 192, torch.float32)])
def test_ln_backward(M, C, dt):
    from gvfdiffusion_b200 import ops
    g = _g(M + C)
    x = _rand((M, C), g, 2.0).to(dt)
    dy = _rand((M, C), g).half()
    dres = _rand((M, C), g).half()
    xf = x.float().requires_grad_(True)
    F.layer_norm(xf, (C,), eps=1e-6).backward(dy.float())
    assert _rel(ops.ln_bwd(x, dy), xf.grad) < 2e-3
    assert _rel(ops.ln_bwd(x, dy, dres), xf.grad + dres.float()) < 2e-3


def test_geglu_backward_transpose_colsum():
    from gvfdiffusion_b200 import ops
    g = _g(7)
    M, Fh = 520, 384
    h = _rand((M, 2 * Fh), g, 1.5).half()
    dG = _rand((M, Fh), g).half()
    hf = h.float().requires_grad_(True)
    a, gt = hf.chunk(2, -1)
    (a * F.gelu(gt)).backward(dG.float())
    assert _rel(ops.geglu_bwd(h, dG), hf.grad) < 2e-3
    for R, C in ((520, 768), (97, 130), (12288, 96), (2, 8)):
        x = _rand((R, C), g).half()
        t = ops.transpose(x)
        R8 = (R + 7) // 8 * 8
        assert t.shape == (C, R8) and torch.equal(t[:, :R], x.T) and (t[:, R:] == 0).all()
    xs = _rand((1000, 2304), g).half()[:, 768:1536]                        # strided view
    assert torch.equal(ops.transpose(xs)[:, :1000], xs.T)
    for dt in (torch.float16, torch.float32):
        x = _rand((5000, 300), g).to(dt)
        ref = x.double().sum(0)
        got = ops.colsum(x)
        assert torch.allclose(got.double(), ref, rtol=1e-5, atol=1e-3)
        assert torch.allclose(ops.colsum(x, out=got.clone(), accumulate=True).double(), 2 * ref, rtol=1e-5, atol=2e-3)
    x14 = _rand((4097, 14), g)
    assert torch.allclose(ops.colsum(x14).double(), x14.double().sum(0), rtol=1e-5, atol=1e-3)


def test_skinny_linears_backward():
    from gvfdiffusion_b200 import ops
    g = _g(11)
    M, N, K = 3000, 768, 14
    x = _rand((M, K), g)
    w = _rand((N, K), g, 0.1).half()
    dy = _rand((M, N), g).half()
    dx = ops.small_linear_bwd_input(dy, w)
    assert _rel(dx, dy.float() @ w.float()) < 1e-5
    dx32 = ops.small_linear_bwd_input(dy.float(), w)
    assert _rel(dx32, dy.float() @ w.float()) < 1e-5
    dwt = ops.skinny_outer(x, dy)                                          # [K, N] = dW^T
    assert _rel(dwt, x.T @ dy.float()) < 1e-5
    acc = ops.skinny_outer(x, dy, out=dwt.clone(), accumulate=True)
    assert _rel(acc, 2 * (x.T @ dy.float())) < 1e-5
    y32 = _rand((M, 96), g)
    assert _rel(ops.skinny_outer(x, y32), x.T @ y32) < 1e-5
    # the vectorised path (fp16, N % 8 == 0, many rows): 8 columns per thread, row groups reduced through shared memory
    for Mb, Nb in ((20000, 768), (393216, 768), (4097, 96), (5000, 200)):
        xb, yb = _rand((Mb, K), g), _rand((Mb, Nb), g).half()
        assert _rel(ops.skinny_outer(xb, yb), xb.T @ yb.float()) < 2e-5, (Mb, Nb)


@pytest.mark.parametrize("C", [96, 192, 768])
def test_query_embed_backward(C):
    """gvf_vae_query_embed_bwd vs autograd of the fp32 expression (model/autoencoder.py:250-301,389-391,560); the forward's
    fp16 roundings are straight-through here as they are for autograd under autocast."""
    from gvfdiffusion_b200 import ops
    g = _g(C)
    Q, E = 257, C // 6
    queries = _rand((Q, 14), g, 0.3)
    gs = _rand((Q, C), g).half()
    dout = _rand((Q, C), g).half()
    om = torch.tensor([1.0 / 10000 ** (j / (E / 2.0)) for j in range(E)], dtype=torch.float64).half().float().to(DEV)
    xyz = queries[:, :3].half().float().requires_grad_(True)
    gsf = gs.float().requires_grad_(True)
    arg = xyz[:, :, None] * om                                              # [Q,3,E]
    pe = torch.cat([torch.sin(arg), torch.cos(arg)], -1).reshape(Q, C)
    u = F.layer_norm(gsf, (C,), eps=1e-5) + F.layer_norm(pe, (C,), eps=1e-5)
    F.layer_norm(u, (C,), eps=1e-6).backward(dout.float())
    dgs, dxyz = ops.vae_query_embed_bwd(queries, gs, dout)
    assert _rel(dgs, gsf.grad) < 3e-3, _rel(dgs, gsf.grad)
    assert _rel(dxyz, xyz.grad) < 2e-2, _rel(dxyz, xyz.grad)               # sin / cos arguments are fp16-rounded in the kernel


def test_skinny_expand():
    from gvfdiffusion_b200 import ops
    g = _g(13)
    for M, K, N in ((3000, 14, 768), (1, 14, 768), (777, 16, 96), (768, 14, 768)):
        x, wt = _rand((M, K), g), _rand((K, N), g, 0.2)
        ref = x @ wt
        assert _rel(ops.skinny_expand(x, wt, out_f16=False), ref) < 1e-6
        assert _rel(ops.skinny_expand(x, wt), ref) < 1e-3


@pytest.mark.parametrize("M,N,K,force", [(768, 768, 12288, 0), (768, 768, 12288, 5), (2304, 768, 12288, 0), (768, 3072, 12296, 0),
                                          (96, 96, 4096, 0), (1536, 768, 12288, 3), (6144, 768, 12288, 0)])
def test_gemm_split_k_fp32_store(M, N, K, force):
    """wgrad shapes of the training step: few output tiles, long reduction -> (tile, k range) work items whose partial
    tiles are summed by TMA reduce-add stores.  Against the fp32 matmul of the same fp16 operands."""
    from gvfdiffusion_b200 import _lib, ops
    g = _g(M + N + K + force)
    a = _rand((M, K), g, 0.5).half()
    w = _rand((N, K), g, 0.5).half()
    b = _rand((N,), g)
    ref = a.float() @ w.float().T
    try:
        _lib.lib().gvf_gemm_set_ksplit(force)
        out = ops.gemm(a, w, None, ops.EPI_F32)
        outb = ops.gemm(a, w, b, ops.EPI_F32)
        _lib.lib().gvf_gemm_set_ksplit(-1)
        plain = ops.gemm(a, w, None, ops.EPI_F32)
    finally:
        _lib.lib().gvf_gemm_set_ksplit(0)
    assert _rel(out, ref) < 5e-5 and _rel(plain, ref) < 5e-5, (_rel(out, ref), _rel(plain, ref))   # one long fp32 chain is the less exact one
    assert _rel(outb, ref + b) < 5e-5


@pytest.mark.parametrize("R,M,N", [(12288, 768, 768), (12288, 6144, 768), (12288, 768, 3072), (12288, 2304, 768), (1000, 96, 288),
                                   (393216, 768, 768), (150, 96, 96), (12290, 1536, 768)])
def test_gemm_tn_weight_gradient(R, M, N):
    """dW = dY^T X straight from the row-major activations: both operands MN-major in the tcgen05 GEMM (no transposed
    copies), split-K over the token dimension.  Also on strided column sub-blocks (the packed q | k | v gradient)."""
    from gvfdiffusion_b200 import ops
    g = _g(R + M + N)
    dy = _rand((R, M), g, 0.5).half()
    x = _rand((R, N), g, 0.5).half()
    ref = dy.float().T @ x.float()
    err = _rel(ops.gemm_tn(dy, x), ref)
    assert err < 5e-5, err
    if M >= 16:
        sub = dy[:, M // 2:]
        assert _rel(ops.gemm_tn(sub, x), sub.float().T @ x.float()) < 5e-5


def test_colsum_single_launch_form_and_gelu_epilogues():
    """The one-launch fp16 column sum (ticketed last-CTA reduction, N % 8 == 0) repeated on the same buffers, and the GELU
    epilogues of the MLP: forward with the pre-activation kept (epilogue 1 + side output), backward fused into the dgrad GEMM
    (epilogue 8) against the two-pass kernels and torch."""
    from gvfdiffusion_b200 import ops
    g = torch.Generator().manual_seed(77)
    for M, N in ((4096, 768), (5000, 2304), (257, 3072), (12288, 8)):
        x = _rand((M, N), g).half()
        ref = x.double().sum(0)
        for _ in range(3):                                        # the ticket counters reset themselves
            got = ops.colsum(x)
            assert torch.allclose(got.double(), ref, rtol=1e-5, atol=2e-3)
        assert torch.equal(ops.colsum(x), got)                    # deterministic
    xs = _rand((3000, 2304), g).half()[:, 768:1536]
    assert torch.allclose(ops.colsum(xs).double(), xs.double().sum(0), rtol=1e-5, atol=2e-3)
    M, K, N = 1000, 768, 3072
    a, w = (_rand((M, K), g) * 0.5).half(), (_rand((N, K), g) * 0.05).half()
    b = _rand((N,), g).half().float()
    h0 = torch.empty((M, N), dtype=torch.float16, device=DEV)
    hg = ops.gemm(a, w, b, ops.EPI_GELU_F16, gate=h0)
    assert torch.equal(h0, ops.gemm(a, w, b, ops.EPI_F16))
    assert torch.equal(hg, ops.gemm(a, w, b, ops.EPI_GELU_F16)) and torch.equal(hg, ops.gelu_tanh(h0))
    dy = _rand((M, K), g).half()
    # w [N, K] in the role of fc2's transposed weight: d hg [M, N] = dy [M, K] @ w^T, then GELU'(h0)
    fused = ops.gemm(dy, w, None, ops.EPI_GELU_BWD_F16, gate=h0)
    two = ops.gelu_tanh_bwd(h0, ops.gemm(dy, w, None, ops.EPI_F16))
    assert _rel(fused, two) < 1e-3
    hr = h0.float().requires_grad_(True)
    F.gelu(hr, approximate="tanh").backward(dy.float() @ w.float().t())
    assert _rel(fused, hr.grad) < 2e-3


@pytest.mark.parametrize("M,R,N", [(4096, 3072, 768), (1000, 768, 3072), (12288, 2304, 768), (333, 64, 96), (257, 16, 128),
                                   (4096, 768, 2304)])
def test_gemm_nn_input_gradient_without_transposed_weights(M, R, N):
    """dX = dY W straight from the row-major [out, in] weight (W as the MN-major B operand of the tcgen05 GEMM): equal to the
    GEMM over a transposed copy bit for bit, also with GELU' in the epilogue, and on a strided column sub-block of dY."""
    from gvfdiffusion_b200 import ops
    g = _g(M + R + N)
    dy = _rand((M, R), g, 0.5).half()
    w = _rand((R, N), g, 0.05).half()
    ref = dy.float() @ w.float()
    out = ops.gemm_nn(dy, w)
    assert _rel(out, ref) < 1e-3
    assert torch.equal(out, ops.gemm(dy, ops.transpose(w)[:, :R].contiguous() if R % 8 == 0 else ops.transpose(w), None, ops.EPI_F16))
    h0 = _rand((M, N), g, 1.5).half()
    fused = ops.gemm_nn(dy, w, gelu_bwd_gate=h0)
    assert _rel(fused, ops.gelu_tanh_bwd(h0, out)) < 1e-3
    if R >= 128:
        sub = dy[:, R // 2:]
        assert _rel(ops.gemm_nn(sub, w[R // 2:]), sub.float() @ w[R // 2:].float()) < 1e-3
