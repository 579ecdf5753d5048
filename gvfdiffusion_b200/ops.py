"""Thin torch-tensor front-ends of the C ABI (include/gvf_b200.h sections 2-5).

torch is used for device memory and the current stream; every computation happens in
libgvf_b200.so.  All functions enqueue on the current CUDA stream and never synchronise.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import check, current_stream, ptr

F16, F32 = torch.float16, torch.float32
EPI_F16, EPI_GELU_F16, EPI_RESID_F32, EPI_RESID_F16, EPI_F32, EPI_F32_COMPACT = 0, 1, 2, 3, 4, 5
EPI_GELU_BWD_F16 = 8          # out = fp16(acc) * gelu_tanh'(gate): gate = the saved fp16 pre-activation [M, N]


def _ll(vals):
    return (C.c_longlong * len(vals))(*[int(v) for v in vals])


def _req(t, dtype, name):
    if not (t.is_cuda and t.dtype == dtype):
        raise ValueError(f"{name}: expected a CUDA {dtype} tensor, got {t.dtype} on {t.device}")


def gemm(a, w, bias=None, epilogue=EPI_F16, out=None, gate=None, gate_stride=0, rows_per_batch=0):
    """out = epilogue(a[M,K] @ w[N,K]^T + bias).  a, w fp16 with contiguous rows."""
    _req(a, F16, "a")
    _req(w, F16, "w")
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K and a.stride(1) == 1 and w.stride(1) == 1
    if out is None:
        assert epilogue in (EPI_F16, EPI_GELU_F16, EPI_F32, EPI_GELU_BWD_F16)
        out = torch.empty((M, N), dtype=F32 if epilogue == EPI_F32 else F16, device=a.device)
    assert out.stride(1) == 1 and out.dtype == (F32 if epilogue in (EPI_RESID_F32, EPI_F32, EPI_F32_COMPACT) else F16)
    if epilogue in (EPI_GELU_F16, EPI_GELU_BWD_F16) and gate is not None:
        # GELU forward: `gate` RECEIVES the fp16 pre-activation; GELU backward: `gate` IS the saved pre-activation
        _req(gate, F16, "gate")
        assert gate.shape == (M, N) and gate.stride(1) == 1
        gate_stride = gate.stride(0)
    st = _lib.lib().gvf_gemm_f16(ptr(a), a.stride(0), ptr(w), w.stride(0), M, N, K, epilogue, ptr(bias),
                                 ptr(out), out.stride(0), ptr(gate), gate_stride, rows_per_batch,
                                 current_stream())
    check(st, "gvf_gemm_f16")
    return out


def gemm_qkv_rmsnorm(a, w, bias, gamma_q, gamma_k, out):
    """QKV Linear + per-head q/k RMS norm (head dim 32) in one kernel; out fp16 [M, 3C]."""
    M, K = a.shape
    N = w.shape[0]
    norm_cols = 2 * gamma_q.numel()
    assert gamma_q.shape[-1] == 32 and norm_cols <= N
    st = _lib.lib().gvf_gemm_qkv_rmsnorm_f16(ptr(a), a.stride(0), ptr(w), w.stride(0), M, N, K, ptr(bias), ptr(out),
                                             out.stride(0), ptr(gamma_q), ptr(gamma_k), norm_cols, current_stream())
    check(st, "gvf_gemm_qkv_rmsnorm_f16")
    return out


def gemm_resid_ln(a, w, bias, x, y, gate=None, gate_stride=0, rows_per_batch=0, ln_w=None, ln_b=None, shift=None,
                  scale=None, mod_stride=0, eps=1e-6):
    """x[M,512] += gate * fp16(a @ w^T + bias) in place (fp32), y = fp16(LN(x) modulated) -- one kernel."""
    _req(a, F16, "a")
    _req(w, F16, "w")
    _req(x, F32, "x")
    _req(y, F16, "y")
    M, K = a.shape
    N = w.shape[0]
    st = _lib.lib().gvf_gemm_resid_ln_f16(ptr(a), a.stride(0), ptr(w), w.stride(0), M, N, K, ptr(bias), ptr(x),
                                          x.stride(0), ptr(gate), gate_stride, rows_per_batch, ptr(ln_w), ptr(ln_b),
                                          ptr(shift), ptr(scale), mod_stride, eps, ptr(y), y.stride(0),
                                          current_stream())
    check(st, "gvf_gemm_resid_ln_f16")
    return y


def geglu_interleave(w, bias=None):
    """Rows of FeedForward.net[0] ([2F, K]: F value rows then F gate rows, model/autoencoder.py:90-101) re-ordered for
    `gemm_geglu`: per 256-row tile 128 value rows followed by their 128 gate rows.  F % 128 == 0."""
    F2 = w.shape[0]
    Fh = F2 // 2
    if Fh % 128:
        raise ValueError("geglu_interleave: hidden width must be a multiple of 128")
    idx = torch.arange(Fh, device=w.device).view(-1, 128)
    perm = torch.cat([idx, idx + Fh], dim=1).reshape(-1)
    return w[perm].contiguous(), (None if bias is None else bias[perm].contiguous())


def gemm_geglu(a, w_il, bias_il, out=None):
    """out[M, F] = fp16(value * gelu_erf(gate)) of (a @ w^T + bias) with w / bias from `geglu_interleave`: the FF's
    first Linear and GEGLU in one kernel (the [M, 2F] hidden tensor never exists)."""
    _req(a, F16, "a")
    _req(w_il, F16, "w")
    M, K = a.shape
    N = w_il.shape[0]
    assert w_il.shape[1] == K and a.stride(1) == 1 and w_il.stride(1) == 1 and N % 256 == 0
    if out is None:
        out = torch.empty((M, N // 2), dtype=F16, device=a.device)
    assert out.dtype == F16 and out.stride(1) == 1 and out.shape[1] == N // 2
    check(_lib.lib().gvf_gemm_geglu_f16(ptr(a), a.stride(0), ptr(w_il), w_il.stride(0), M, N, K, ptr(bias_il), ptr(out),
                                        out.stride(0), current_stream()), "gvf_gemm_geglu_f16")
    return out


_attn_ws = None


def _ensure_attn_ws(device):
    """Scratch for the key-range split of the last attention wave (gvf_attn_set_workspace), once per process."""
    global _attn_ws
    if _attn_ws is None:
        _attn_ws = torch.empty(147 * 3 * 512 * 34, dtype=F32, device=device)
        _lib.lib().gvf_attn_set_workspace(ptr(_attn_ws), _attn_ws.numel() * 4)


def attention(q, k, v, scale, out=None, q_shared=False, kv_shared=False):
    """q [Nb,Lq,H,D] (or [Lq,H,D] if q_shared), k/v [Nb,Lk,H,D] (or [Lk,H,D] if kv_shared): fp16
    views with contiguous last dim (any strides that are multiples of 8).  -> [Nb,Lq,H,D] fp16."""
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _req(t, F16, n)
        assert t.stride(-1) == 1
    if q_shared:
        Lq, H, D = q.shape
        qs = (0, q.stride(0), q.stride(1))
    else:
        _, Lq, H, D = q.shape
        qs = (q.stride(0), q.stride(1), q.stride(2))
    if kv_shared:
        Lk = k.shape[0]
        ks, vs = (0, k.stride(0), k.stride(1)), (0, v.stride(0), v.stride(1))
        Nb = q.shape[0]
    else:
        Nb, Lk = k.shape[0], k.shape[1]
        ks, vs = (k.stride(0), k.stride(1), k.stride(2)), (v.stride(0), v.stride(1), v.stride(2))
    if out is None:
        out = torch.empty((Nb, Lq, H, D), dtype=F16, device=q.device)
    os_ = (out.stride(0), out.stride(1), out.stride(2))
    _ensure_attn_ws(q.device)
    st = _lib.lib().gvf_attn_fwd_f16(ptr(q), ptr(k), ptr(v), ptr(out), Nb, Lq, Lk, H, D, _ll(qs), _ll(ks),
                                     _ll(vs), _ll(os_), int(q_shared), int(kv_shared), float(scale),
                                     current_stream())
    check(st, "gvf_attn_fwd_f16")
    return out


def small_linear(x, w, bias, out_f16, add=None, add_rows=0, out=None):
    _req(x, F32, "x")
    _req(w, F16, "w")
    M, K = x.shape
    N = w.shape[0]
    assert x.stride(1) == 1 and w.is_contiguous()
    if out is None:
        out = torch.empty((M, N), dtype=F16 if out_f16 else F32, device=x.device)
    st = _lib.lib().gvf_small_linear(ptr(x), x.stride(0), ptr(w), ptr(bias), M, N, K, ptr(add), add_rows,
                                     ptr(out), int(out_f16), current_stream())
    check(st, "gvf_small_linear")
    return out


def ln_mod(x, out=None, eps=1e-6, w=None, b=None, shift=None, scale=None, mod_stride=0, rows_per_batch=0):
    M, Cc = x.shape
    assert x.is_contiguous() and x.dtype in (F16, F32)
    if out is None:
        out = torch.empty((M, Cc), dtype=F16, device=x.device)
    st = _lib.lib().gvf_ln_mod_f16(ptr(x), int(x.dtype == F16), ptr(out), M, Cc, eps, ptr(w), ptr(b),
                                   ptr(shift), ptr(scale), mod_stride, rows_per_batch, current_stream())
    check(st, "gvf_ln_mod_f16")
    return out


def ln_mod_act(x, out=None, eps=1e-6, w=None, b=None, shift=None, scale=None, mod_stride=0, rows_per_batch=0, act=1):
    """ln_mod with an activation before the fp16 rounding (act 1 = SiLU): the norm -> SiLU pairs of SparseResBlock3d."""
    M, Cc = x.shape
    assert x.is_contiguous() and x.dtype in (F16, F32)
    if out is None:
        out = torch.empty((M, Cc), dtype=F16, device=x.device)
    st = _lib.lib().gvf_ln_mod_act_f16(ptr(x), int(x.dtype == F16), ptr(out), M, Cc, eps, ptr(w), ptr(b),
                                       ptr(shift), ptr(scale), mod_stride, rows_per_batch, act, current_stream())
    check(st, "gvf_ln_mod_act_f16")
    return out


def sparse_pool_mean(x, order, offsets, out=None):
    """SparseDownsample's pooling: x fp16 [N, C]; order int32 [N] (fine rows grouped by coarse cell), offsets int32
    [cells + 1] -> fp16 [cells, C] = sum / (count + 1) (the reference's include_self mean)."""
    _req(x, F16, "x")
    _req(order, torch.int32, "order")
    _req(offsets, torch.int32, "offsets")
    cells, Cc = offsets.shape[0] - 1, x.shape[1]
    assert x.stride(1) == 1 and order.is_contiguous() and offsets.is_contiguous() and order.shape[0] == x.shape[0]
    if out is None:
        out = torch.empty((cells, Cc), dtype=F16, device=x.device)
    check(_lib.lib().gvf_sparse_pool_mean_f16(ptr(x), x.stride(0), ptr(order), ptr(offsets), cells, Cc, ptr(out), out.stride(0),
                                              current_stream()), "gvf_sparse_pool_mean_f16")
    return out


def gather_concat(a=None, idx=None, b=None, out=None):
    """out[i] = [a[idx[i]] (or a[i]) | b[i]] (fp16): SparseUpsample's gather and / or the skip concatenation."""
    rows = idx.shape[0] if idx is not None else (a.shape[0] if a is not None else b.shape[0])
    Ca = 0 if a is None else a.shape[1]
    Cb = 0 if b is None else b.shape[1]
    for t, nm in ((a, "a"), (b, "b")):
        if t is not None:
            _req(t, F16, nm)
            assert t.stride(1) == 1
    if idx is not None:
        _req(idx, torch.int32, "idx")
        assert idx.is_contiguous()
    if b is not None:
        assert b.shape[0] == rows
    if out is None:
        out = torch.empty((rows, Ca + Cb), dtype=F16, device=(a if a is not None else b).device)
    check(_lib.lib().gvf_gather_concat_f16(ptr(a), a.stride(0) if a is not None else 8, Ca, ptr(idx), ptr(b),
                                           b.stride(0) if b is not None else 8, Cb, rows, ptr(out), out.stride(0),
                                           current_stream()), "gvf_gather_concat_f16")
    return out


def sparse_tap_gather_sum(P, nbr, idx=None, bias=None, out=None):
    """P fp32 [Nc, K3 * Cout] (per-tap products of the coarse rows), nbr int32 [N, K3], idx int32 [N] or None ->
    fp16 [N, Cout] = bias + sum_k P[idx[nbr[:, k]], k * Cout : (k + 1) * Cout]."""
    _req(P, F32, "P")
    _req(nbr, torch.int32, "nbr")
    N, K3 = nbr.shape
    Cout = P.shape[1] // K3
    assert P.stride(1) == 1 and nbr.is_contiguous() and P.shape[1] == K3 * Cout
    if idx is not None:
        _req(idx, torch.int32, "idx")
        assert idx.is_contiguous()
    if out is None:
        out = torch.empty((N, Cout), dtype=F16, device=P.device)
    check(_lib.lib().gvf_sparse_tap_gather_sum_f16(ptr(P), P.stride(0), ptr(nbr), ptr(idx), N, K3, Cout, ptr(bias), ptr(out),
                                                   out.stride(0), current_stream()), "gvf_sparse_tap_gather_sum_f16")
    return out


def rmsnorm_heads_(buf, H, D, k_off, gamma_q, gamma_k):
    rows, ld = buf.shape[0], buf.stride(0)
    st = _lib.lib().gvf_rmsnorm_heads_f16(ptr(buf), rows, ld, H, D, k_off, ptr(gamma_q), ptr(gamma_k),
                                          current_stream())
    check(st, "gvf_rmsnorm_heads_f16")
    return buf


def dit_modulation(t, W0, b0, W2, b2, Wmod, bmod, temb, silu_temb, mod_out):
    B = t.shape[0]
    Cc, Fq = W0.shape
    R = Wmod.shape[0]
    st = _lib.lib().gvf_dit_modulation(ptr(t), B, Cc, Fq, ptr(W0), ptr(b0), ptr(W2), ptr(b2), ptr(Wmod),
                                       ptr(bmod), R, ptr(temb), ptr(silu_temb), ptr(mod_out), current_stream())
    check(st, "gvf_dit_modulation")
    return mod_out


def ape(xyz, Cc, out=None):
    R = xyz.shape[0]
    if out is None:
        out = torch.empty((R, Cc), dtype=F32, device=xyz.device)
    check(_lib.lib().gvf_ape(ptr(xyz), R, Cc, ptr(out), current_stream()), "gvf_ape")
    return out


def vae_query_embed(queries, gs):
    Q, Cc = gs.shape
    out = torch.empty((Q, Cc), dtype=F16, device=gs.device)
    check(_lib.lib().gvf_vae_query_embed(ptr(queries), queries.stride(0), ptr(gs), Q, Cc, ptr(out),
                                         current_stream()), "gvf_vae_query_embed")
    return out


def geglu(h, out=None):
    M, F2 = h.shape
    if out is None:
        out = torch.empty((M, F2 // 2), dtype=F16, device=h.device)
    check(_lib.lib().gvf_geglu_f16(ptr(h), M, F2 // 2, ptr(out), current_stream()), "gvf_geglu_f16")
    return out


def cast_f16(x, out=None):
    x = x.contiguous()
    if out is None:
        out = torch.empty(x.shape, dtype=F16, device=x.device)
    check(_lib.lib().gvf_cast_f32_f16(ptr(x), x.numel(), ptr(out), current_stream()), "gvf_cast_f32_f16")
    return out


def dit_final_layer(x, shift, scale, mod_stride, rows_per_batch, W, bias, out=None):
    M, Cc = x.shape
    O = W.shape[0]
    if out is None:
        out = torch.empty((M, O), dtype=F32, device=x.device)
    check(_lib.lib().gvf_dit_final_layer(ptr(x), M, Cc, O, ptr(shift), ptr(scale), mod_stride, rows_per_batch,
                                         ptr(W), ptr(bias), ptr(out), current_stream()), "gvf_dit_final_layer")
    return out


def dpm_x0(x, v, branches, alpha, sigma, s1, s2, out, model_type=1):
    check(_lib.lib().gvf_dpm_x0(ptr(x), ptr(v), x.numel(), branches, model_type, alpha, sigma, s1, s2, ptr(out),
                                current_stream()), "gvf_dpm_x0")
    return out


def dpm_error_sq(x_higher, x_lower, x_prev, atol, rtol, E2):
    B = x_higher.shape[0]
    check(_lib.lib().gvf_dpm_error_sq(ptr(x_higher), ptr(x_lower), ptr(x_prev), B, x_higher.numel() // B, atol,
                                      rtol, ptr(E2), current_stream()), "gvf_dpm_error_sq")
    return E2


def dpm_update(x, m0, m1, cx, cm, inv_r0, order, out):
    check(_lib.lib().gvf_dpm_update(ptr(x), ptr(m0), ptr(m1), x.numel(), cx, cm, inv_r0, order, ptr(out),
                                    current_stream()), "gvf_dpm_update")
    return out


def affine_lastdim(x, a=None, b=None, a_scalar=1.0, b_scalar=0.0, out=None):
    if out is None:
        out = torch.empty_like(x)
    check(_lib.lib().gvf_affine_lastdim(ptr(x), x.numel(), x.shape[-1], ptr(a), ptr(b), a_scalar, b_scalar,
                                        ptr(out), current_stream()), "gvf_affine_lastdim")
    return out


def gaussian_tensor(prm, arrays):
    """raw canonical GaussianModel arrays -> activated [P,14] (train_vae.py:466-472)."""
    P = arrays[0].shape[0]
    out = torch.empty((P, 14), dtype=F32, device=arrays[0].device)
    check(_lib.lib().gvf_gaussian_tensor(C.byref(prm), P, *[ptr(a) for a in arrays], ptr(out),
                                         current_stream()), "gvf_gaussian_tensor")
    return out


class GaussianTensorFn(torch.autograd.Function):
    """`gaussian_tensor` under autograd (forward gvf_gaussian_tensor, backward gvf_gaussian_tensor_bwd)."""

    @staticmethod
    def forward(ctx, prm, xyz, dc, scaling, rotation, opacity):
        arrays = [t.detach().to(F32).contiguous() for t in (xyz, dc, scaling, rotation, opacity)]
        ctx.prm, ctx.shapes = prm, [t.shape for t in (xyz, dc, scaling, rotation, opacity)]
        ctx.save_for_backward(arrays[2], arrays[3], arrays[4])
        return gaussian_tensor(prm, arrays)

    @staticmethod
    def backward(ctx, g):
        scaling, rotation, opacity = ctx.saved_tensors
        P = scaling.shape[0]
        g = g.to(F32).contiguous()
        outs = [torch.empty(s, dtype=F32, device=g.device) for s in ctx.shapes]
        check(_lib.lib().gvf_gaussian_tensor_bwd(C.byref(ctx.prm), P, ptr(scaling), ptr(rotation), ptr(opacity), ptr(g),
                                                 *[ptr(o) for o in outs], current_stream()), "gvf_gaussian_tensor_bwd")
        return (None, *outs)


def fps(points, K, start=0, spatially_ordered=False):
    """points [P, >=3] fp32 (xyz first) -> int32 indices [K] of a farthest point sample.  spatially_ordered: the rows are
    voxel-major in lexicographic voxel order (what to_representation emits) -> the exactly pruned kernel, same indices."""
    _req(points, F32, "points")
    P = points.shape[0]
    assert points.stride(1) == 1
    ws = torch.empty(P, dtype=F32, device=points.device)
    idx = torch.empty(K, dtype=torch.int32, device=points.device)
    fn = _lib.lib().gvf_fps_ordered if spatially_ordered else _lib.lib().gvf_fps
    check(fn(ptr(points), points.stride(0), P, K, start, ptr(ws), ptr(idx), current_stream()), "gvf_fps")
    return idx


# ---------------------------------------------------------------- optional launch timing
class LaunchTimer:
    """Records CUDA events around selected launches (bench.py's live roofline numbers)."""

    def __init__(self):
        self.records = {}

    def wrap(self, name, fn):
        def timed(*a, **k):
            tag = k.pop("_tag", name)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a, **k)
            e1.record()
            self.records.setdefault(tag, []).append((e0, e1))
            return r
        return timed

    def summary(self):
        torch.cuda.synchronize()
        return {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in self.records.items()}


# ---- voxel-side operators (include/gvf_b200.h section 7) ----
def to_representation(feats, coords, num_gaussians, lr, resolution, reg_mode=0, voxel_size=1.0, perturbation=None):
    """SparseVAE.to_representation in one launch: feats fp32 [Nvox, 14 G], coords int32 [Nvox, 4] ->
    (_xyz [P,3], _features_dc [P,1,3], _scaling [P,3], _rotation [P,4], _opacity [P,1]), P = Nvox * G."""
    _req(feats, F32, "feats")
    _req(coords, torch.int32, "coords")
    assert feats.stride(1) == 1 and coords.is_contiguous() and coords.shape[1] == 4
    nvox, G = feats.shape[0], int(num_gaussians)
    assert feats.shape[1] >= 14 * G and coords.shape[0] == nvox
    P = nvox * G
    e = lambda *s: torch.empty(s, dtype=F32, device=feats.device)
    xyz, dc, sc, rot, op = e(P, 3), e(P, 1, 3), e(P, 3), e(P, 4), e(P, 1)
    if perturbation is not None:
        _req(perturbation, F32, "perturbation")
        assert perturbation.is_contiguous() and tuple(perturbation.shape) == (G, 3)
    lr5 = (C.c_float * 5)(*[float(v) for v in lr])
    check(_lib.lib().gvf_to_representation(ptr(feats), feats.stride(0), ptr(coords), nvox, G, ptr(perturbation), lr5,
                                           float(resolution), int(reg_mode), float(voxel_size), ptr(xyz), ptr(dc),
                                           ptr(sc), ptr(rot), ptr(op), current_stream()), "gvf_to_representation")
    return xyz, dc, sc, rot, op


def to_representation_bwd(feats, num_gaussians, lr, resolution, reg_mode=0, voxel_size=1.0, perturbation=None,
                          g_xyz=None, g_dc=None, g_scaling=None, g_rotation=None, g_opacity=None):
    """Backward of `to_representation`: gradients of the raw GaussianModel tensors (fp32, any may be None) -> d feats
    fp32 [Nvox, feats.shape[1]]."""
    _req(feats, F32, "feats")
    feats = feats.contiguous()                           # one row stride for feats and d feats
    nvox, G = feats.shape[0], int(num_gaussians)
    P = nvox * G
    gs = []
    for t, w in ((g_xyz, 3), (g_dc, 3), (g_scaling, 3), (g_rotation, 4), (g_opacity, 1)):
        if t is not None:
            t = t.to(F32).contiguous()
            assert t.is_cuda and t.numel() == P * w
        gs.append(t)
    out = torch.empty((nvox, feats.shape[1]), dtype=F32, device=feats.device)
    lr5 = (C.c_float * 5)(*[float(v) for v in lr])
    check(_lib.lib().gvf_to_representation_bwd(ptr(feats), feats.stride(0), nvox, G, ptr(perturbation), lr5, float(resolution),
                                               int(reg_mode), float(voxel_size), *[ptr(t) for t in gs], ptr(out),
                                               current_stream()), "gvf_to_representation_bwd")
    return out


def sparse_neighbor_map(coords, batch_size, grid_size, ksize=3, dilation=1, workspace=None, status=None):
    """coords int32 [N,4] (batch, x, y, z) -> nbr int32 [N, ksize^3] (-1 = no voxel there)."""
    _req(coords, torch.int32, "coords")
    assert coords.is_contiguous() and coords.shape[1] == 4
    N = coords.shape[0]
    need = _lib.lib().gvf_sparse_conv_workspace_bytes(int(batch_size), int(grid_size))
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=coords.device)
    nbr = torch.empty((N, ksize ** 3), dtype=torch.int32, device=coords.device)
    check(_lib.lib().gvf_sparse_neighbor_map(ptr(coords), N, int(batch_size), int(grid_size), int(ksize), int(dilation),
                                             ptr(workspace), workspace.numel() * workspace.element_size(), ptr(nbr),
                                             ptr(status), current_stream()), "gvf_sparse_neighbor_map")
    return nbr


def sparse_im2col(x, nbr, out=None):
    """x fp16 / fp32 [N, Cin], nbr int32 [N, K3] -> fp16 [N, K3 * Cin] (zero rows where nbr < 0)."""
    assert x.is_cuda and x.dtype in (F16, F32) and x.stride(1) == 1
    _req(nbr, torch.int32, "nbr")
    N, K3 = nbr.shape
    Cin = x.shape[1]
    if out is None:
        out = torch.empty((N, K3 * Cin), dtype=F16, device=x.device)
    assert out.is_contiguous() and out.dtype == F16
    check(_lib.lib().gvf_sparse_im2col_f16(ptr(x), int(x.dtype == F16), x.stride(0), ptr(nbr), N, K3, Cin, ptr(out),
                                           current_stream()), "gvf_sparse_im2col_f16")
    return out


# ---- training step of the motion VAE (include/gvf_b200.h section 8) ----
def _st3(t, shared=False):
    return (0, t.stride(0), t.stride(1)) if shared else (t.stride(0), t.stride(1), t.stride(2))


def attention_fwd_lse(q, k, v, scale, out=None, q_shared=False):
    """`attention` that also returns LSE2 [Nb, H, Lq rounded up to 128] fp32 for `attention_bwd`."""
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _req(t, F16, n)
        assert t.stride(-1) == 1
    Lq, H, D = (q.shape if q_shared else q.shape[1:])
    Nb, Lk = k.shape[0], k.shape[1]
    if out is None:
        out = torch.empty((Nb, Lq, H, D), dtype=F16, device=q.device)
    ld = (Lq + 127) // 128 * 128
    lse = torch.empty((Nb, H, ld), dtype=F32, device=q.device)
    st = _lib.lib().gvf_attn_fwd_lse_f16(ptr(q), ptr(k), ptr(v), ptr(out), ptr(lse), ld, Nb, Lq, Lk, H, D,
                                         _ll(_st3(q, q_shared)), _ll(_st3(k)), _ll(_st3(v)), _ll(_st3(out)),
                                         int(q_shared), 0, float(scale), current_stream())
    check(st, "gvf_attn_fwd_lse_f16")
    return out, lse


def attention_bwd(q, k, v, o, dout, lse, scale, dq, dk, dv, q_shared=False):
    """dq / dk / dv (fp16 views, written in place) of softmax(scale q k^T) v; q_shared: dq [Lq,H,D] sums over the batch."""
    for t, n in ((q, "q"), (k, "k"), (v, "v"), (o, "o"), (dout, "dout"), (dq, "dq"), (dk, "dk"), (dv, "dv")):
        _req(t, F16, n)
        assert t.stride(-1) == 1
    Lq, H, D = (q.shape if q_shared else q.shape[1:])
    Nb, Lk = k.shape[0], k.shape[1]
    dsum = torch.empty_like(lse)
    st = _lib.lib().gvf_attn_bwd_f16(ptr(q), ptr(k), ptr(v), ptr(o), ptr(dout), ptr(lse), ptr(dsum), ptr(dq), ptr(dk),
                                     ptr(dv), Nb, Lq, Lk, H, D, lse.shape[-1], _ll(_st3(q, q_shared)), _ll(_st3(k)),
                                     _ll(_st3(v)), _ll(_st3(o)), _ll(_st3(dout)), _ll(_st3(dq, q_shared)), _ll(_st3(dk)),
                                     _ll(_st3(dv)), int(q_shared), float(scale), current_stream())
    check(st, "gvf_attn_bwd_f16")
    return dq, dk, dv


def transpose(x, out=None):
    """x fp16 [R, C] (row stride any) -> [C, R8] with R8 = R rounded up to 8, pad columns zero; returns the padded
    tensor (a legal K-major GEMM operand whose reduction length is R8)."""
    _req(x, F16, "x")
    R, Cc = x.shape
    assert x.stride(1) == 1
    R8 = (R + 7) // 8 * 8
    if out is None:
        out = torch.empty((Cc, R8), dtype=F16, device=x.device)
    assert out.shape == (Cc, R8) and out.is_contiguous()
    check(_lib.lib().gvf_transpose_f16(ptr(x), R, Cc, x.stride(0), ptr(out), R8, current_stream()), "gvf_transpose_f16")
    return out


_red_ws = {}


def _reduce_ws(device, nbytes):
    ws = _red_ws.get(device)
    if ws is None or ws.numel() * 4 < nbytes:
        ws = torch.empty((nbytes + 3) // 4, dtype=F32, device=device)
        _red_ws[device] = ws
    return ws


def colsum(x, out=None, accumulate=False):
    """sum over rows of x [M, N] (fp16 / fp32) -> fp32 [N]: bias gradients."""
    M, N = x.shape
    assert x.stride(1) == 1 and x.dtype in (F16, F32)
    need = _lib.lib().gvf_colsum_workspace_bytes(M, N, 0)
    ws = _reduce_ws(x.device, need)
    if out is None:
        out = torch.empty(N, dtype=F32, device=x.device)
        accumulate = False
    check(_lib.lib().gvf_colsum(ptr(x), int(x.dtype == F16), M, N, x.stride(0), ptr(ws), ws.numel() * 4, ptr(out),
                                int(accumulate), current_stream()), "gvf_colsum")
    return out


def ln_bwd(x, dy, dres=None, eps=1e-6, out=None):
    M, Cc = x.shape
    assert x.is_contiguous() and dy.is_contiguous() and dy.dtype == F16 and (dres is None or dres.is_contiguous())
    if out is None:
        out = torch.empty((M, Cc), dtype=F16, device=x.device)
    check(_lib.lib().gvf_ln_bwd_f16(ptr(x), int(x.dtype == F16), ptr(dy), ptr(dres), ptr(out), M, Cc, eps,
                                    current_stream()), "gvf_ln_bwd_f16")
    return out


def geglu_bwd(h, dG, out=None):
    M, F2 = h.shape
    assert h.is_contiguous() and dG.is_contiguous() and h.dtype == F16 and dG.dtype == F16
    if out is None:
        out = torch.empty_like(h)
    check(_lib.lib().gvf_geglu_bwd_f16(ptr(h), ptr(dG), M, F2 // 2, ptr(out), current_stream()), "gvf_geglu_bwd_f16")
    return out


def gelu_tanh(h, out=None):
    """GELU(approximate="tanh") of an fp16 pre-activation (training forward of the sparse trunk's MLP)."""
    _req(h, F16, "h")
    assert h.is_contiguous()
    if out is None:
        out = torch.empty_like(h)
    check(_lib.lib().gvf_gelu_tanh_f16(ptr(h), h.numel(), ptr(out), current_stream()), "gvf_gelu_tanh_f16")
    return out


def gelu_tanh_bwd(h, dy, out=None):
    _req(h, F16, "h")
    _req(dy, F16, "dy")
    assert h.is_contiguous() and dy.is_contiguous() and dy.shape == h.shape
    if out is None:
        out = torch.empty_like(h)
    check(_lib.lib().gvf_gelu_tanh_bwd_f16(ptr(h), ptr(dy), h.numel(), ptr(out), current_stream()), "gvf_gelu_tanh_bwd_f16")
    return out


def small_linear_bwd_input(dy, w):
    """dx fp32 [M, K] = dy [M, N] (fp16 / fp32) @ w [N, K] (fp16), K <= 32."""
    M, N = dy.shape
    K = w.shape[1]
    assert dy.stride(1) == 1 and w.is_contiguous() and w.dtype == F16 and w.shape[0] == N
    dx = torch.empty((M, K), dtype=F32, device=dy.device)
    check(_lib.lib().gvf_small_linear_bwd_input(ptr(dy), int(dy.dtype == F16), dy.stride(0), ptr(w), M, N, K, ptr(dx), K,
                                                current_stream()), "gvf_small_linear_bwd_input")
    return dx


def skinny_outer(x, y, out=None, accumulate=False):
    """fp32 [K, N] = x[M, K]^T @ y[M, N] for K <= 16 (x fp32, y fp16 / fp32)."""
    _req(x, F32, "x")
    M, K = x.shape
    N = y.shape[1]
    assert x.stride(1) == 1 and y.stride(1) == 1 and y.shape[0] == M
    need = _lib.lib().gvf_colsum_workspace_bytes(M, N, K)
    ws = _reduce_ws(x.device, need)
    if out is None:
        out = torch.empty((K, N), dtype=F32, device=x.device)
        accumulate = False
    check(_lib.lib().gvf_skinny_outer(ptr(x), x.stride(0), K, ptr(y), int(y.dtype == F16), y.stride(0), M, N, ptr(ws),
                                      ws.numel() * 4, ptr(out), int(accumulate), current_stream()), "gvf_skinny_outer")
    return out


def vae_query_embed_bwd(queries, gs, dout, dxyz=None, accumulate=False):
    """-> (d gs fp16 [Q, C], d xyz): d xyz is written (or added) into the first three columns of `dxyz` (fp32 rows)."""
    Q, Cc = gs.shape
    dgs = torch.empty((Q, Cc), dtype=F16, device=gs.device)
    if dxyz is None:
        dxyz = torch.empty((Q, 3), dtype=F32, device=gs.device)
        accumulate = False
    assert dxyz.dtype == F32 and dxyz.stride(1) == 1 and dxyz.shape[0] == Q
    check(_lib.lib().gvf_vae_query_embed_bwd(ptr(queries), queries.stride(0), ptr(gs), ptr(dout), Q, Cc, ptr(dgs), ptr(dxyz),
                                             dxyz.stride(0), int(accumulate), current_stream()), "gvf_vae_query_embed_bwd")
    return dgs, dxyz


def skinny_expand(x, wt, out_f16=True):
    """x fp32 [M, K <= 16] @ wt fp32 [K, N] -> [M, N] fp16 / fp32."""
    _req(x, F32, "x")
    _req(wt, F32, "wt")
    M, K = x.shape
    N = wt.shape[1]
    assert x.stride(1) == 1 and wt.is_contiguous() and wt.shape[0] == K
    out = torch.empty((M, N), dtype=F16 if out_f16 else F32, device=x.device)
    check(_lib.lib().gvf_skinny_expand(ptr(x), x.stride(0), K, ptr(wt), M, N, ptr(out), int(out_f16), N, current_stream()),
          "gvf_skinny_expand")
    return out


def gemm_nn(a, w, out=None, gelu_bwd_gate=None):
    """fp16 [M, N] = a[M, R] @ w[R, N] (w row-major = a Linear's own [out, in] weight): the input gradient dX = dY W with no
    transposed weight copy.  gelu_bwd_gate: the saved fp16 pre-activation [M, N] -> the result is multiplied by gelu_tanh'."""
    _req(a, F16, "a")
    _req(w, F16, "w")
    M, R = a.shape
    N = w.shape[1]
    assert w.shape[0] == R and a.stride(1) == 1 and w.stride(1) == 1
    if out is None:
        out = torch.empty((M, N), dtype=F16, device=a.device)
    g = gelu_bwd_gate
    if g is not None:
        _req(g, F16, "gate")
        assert g.shape == (M, N) and g.stride(1) == 1
    check(_lib.lib().gvf_gemm_nn_f16(ptr(a), a.stride(0), ptr(w), w.stride(0), M, N, R, 8 if g is not None else 0, ptr(out),
                                     out.stride(0), ptr(g), g.stride(0) if g is not None else 0, current_stream()), "gvf_gemm_nn_f16")
    return out


def gemm_tn(a, w, out=None):
    """fp32 [M, N] = a[R, M]^T @ w[R, N] (fp16 row-major activations): the weight gradient dW = dY^T X, no transposes."""
    _req(a, F16, "a")
    _req(w, F16, "w")
    R, M = a.shape
    N = w.shape[1]
    assert w.shape[0] == R and a.stride(1) == 1 and w.stride(1) == 1
    if out is None:
        out = torch.empty((M, N), dtype=F32, device=a.device)
    check(_lib.lib().gvf_gemm_tn_f16(ptr(a), a.stride(0), ptr(w), w.stride(0), M, N, R, ptr(out), out.stride(0),
                                     current_stream()), "gvf_gemm_tn_f16")
    return out


def sparse_conv_gemm(x, nbr, w, bias=None, out_f32=False, out=None, residual=False):
    """Gather-fused submanifold convolution: x fp16 [N, Cin], nbr int32 [N, K3], w fp16 [Cout, K3 * Cin] -> [N, Cout].
    residual: `out` (fp16, given) += fp16(conv + bias) in place."""
    _req(x, F16, "x")
    _req(w, F16, "w")
    _req(nbr, torch.int32, "nbr")
    N, Cin = x.shape
    K3, Cout = nbr.shape[1], w.shape[0]
    assert x.stride(1) == 1 and w.stride(1) == 1 and nbr.is_contiguous() and w.shape[1] == K3 * Cin
    if out is None:
        assert not residual
        out = torch.empty((N, Cout), dtype=F32 if out_f32 else F16, device=x.device)
    assert not residual or out.dtype == F16
    check(_lib.lib().gvf_sparse_conv_gemm_f16(ptr(x), x.stride(0), ptr(nbr), N, K3, Cin, ptr(w), w.stride(0), Cout, ptr(bias),
                                              ptr(out), out.stride(0), 3 if residual else (4 if out.dtype == F32 else 0),
                                              current_stream()),
          "gvf_sparse_conv_gemm_f16")
    return out
