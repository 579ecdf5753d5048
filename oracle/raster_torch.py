"""CPU oracle for the rasteriser BACKWARD (TEST INFRASTRUCTURE, see oracle/__init__.py).

A differentiable torch restatement of the forward in oracle/raster.c (same stages, same
constants, same tile culling and thresholds); `torch.autograd` over it yields the reference
gradients the sm_100a backward kernels are checked against.
O(P * H * W) dense evaluation -- small scenes only.  PARITY UNPINNED like raster.c: the
algorithm is the third-party diff_gaussian_rasterization (mip-splatting fork); gradient
conventions that are a choice rather than mathematics follow upstream: the min(0.99, .)
clamp on alpha passes gradients straight through.
"""
import math

import torch
import torch.nn.functional as F

SH_C0 = 0.28209479177387814
TILE = 16


def activate(canon, delta, const):
    """GaussianModel.get_*_with_delta in differentiable torch (gaussian_model.py:98-114)."""
    aabb = torch.tensor(const["aabb"], dtype=torch.float32)
    z = lambda a, b: 0 if delta is None else delta[..., a:b]
    xyz = canon["_xyz"] * aabb[None, 3:] + aabb[None, :3] + z(0, 3)
    act = F.softplus if const["softplus"] else torch.exp
    s = act(canon["_scaling"] + const["scale_bias"] + z(3, 6))
    scales = torch.sqrt(torch.square(s) + const["min_kernel"] ** 2)
    rots = F.normalize(canon["_rotation"] + torch.tensor([1.0, 0, 0, 0])[None] + z(6, 10))
    shs = canon["_features_dc"].reshape(-1, 3) + z(10, 13)
    opac = torch.sigmoid(canon["_opacity"].reshape(-1, 1) + const["opacity_bias"] + z(13, 14)).reshape(-1)
    return xyz, scales, rots, shs, opac


def project(means3D, scales, rots, shs, opac, view_t, proj_t, H, W, tanfovx, tanfovy, kernel_size=0.1,
            scale_modifier=1.0, means2D=None, mip_filter=True):
    """preprocess stage; view_t / proj_t are the transposed matrices handed to the rasteriser.
    means2D: optional zero tensor [P,>=2] in NDC units added to the projected centre, so that its autograd
    gradient is upstream's `dL_dmean2D` (blend-stage gradient of the screen position, d pix / d ndc = W/2, H/2)."""
    V, Pm = view_t.T, proj_t.T                      # V p = view-space point
    p = means3D
    t = p @ V[:3, :3].T + V[:3, 3]
    hom = torch.cat([p, torch.ones_like(p[:, :1])], 1) @ Pm.T
    pw = 1.0 / (hom[:, 3] + 1e-7)
    ndc = hom[:, :2] * pw[:, None]
    pix = torch.stack([((ndc[:, 0] + 1) * W - 1) * 0.5, ((ndc[:, 1] + 1) * H - 1) * 0.5], 1)
    if means2D is not None:
        pix = pix + means2D[:, :2] * torch.tensor([0.5 * W, 0.5 * H])
    s = scale_modifier * scales
    r, x, y, z = rots.unbind(1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                     2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                     2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], 1).reshape(-1, 3, 3)
    L = R * s[:, None, :]
    Sigma = L @ L.transpose(1, 2)
    fx, fy = W / (2 * tanfovx), H / (2 * tanfovy)
    tz = t[:, 2]
    limx, limy = 1.3 * tanfovx, 1.3 * tanfovy
    cx = torch.clamp(t[:, 0] / tz, -limx, limx) * tz
    cy = torch.clamp(t[:, 1] / tz, -limy, limy) * tz
    zero = torch.zeros_like(tz)
    J = torch.stack([fx / tz, zero, -(fx * cx) / (tz * tz), zero, fy / tz, -(fy * cy) / (tz * tz)], 1).reshape(-1, 2, 3)
    T = J @ V[:3, :3]
    cov = T @ Sigma @ T.transpose(1, 2)
    a0, b, c0 = cov[:, 0, 0], cov[:, 0, 1], cov[:, 1, 1]
    ks = kernel_size
    det0 = torch.clamp(a0 * c0 - b * b, min=1e-6)
    det1 = torch.clamp((a0 + ks) * (c0 + ks) - b * b, min=1e-6)
    coef = torch.sqrt(det0 / (det1 + 1e-6) + 1e-6)
    coef = torch.where((det0 <= 1e-6) | (det1 <= 1e-6), torch.zeros_like(coef), coef)
    if not mip_filter:                              # plain 3DGS dilation (diff_gauss): no opacity compensation
        coef = torch.ones_like(coef)
    a, c = a0 + ks, c0 + ks
    det = a * c - b * b
    conic = torch.stack([c / det, -b / det, a / det], 1)
    mid = 0.5 * (a + c)
    lam = mid + torch.sqrt(torch.clamp(mid * mid - det, min=0.1))
    radius = torch.ceil(3.0 * torch.sqrt(lam)).detach()
    rgb = torch.clamp(SH_C0 * shs + 0.5, min=0.0)
    gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    pd = pix.detach()
    trunc = lambda v: torch.clamp(v, -1e6, 1e6).to(torch.int64)     # (int) truncation toward zero
    x0 = trunc((pd[:, 0] - radius) / TILE).clamp(0, gx)
    y0 = trunc((pd[:, 1] - radius) / TILE).clamp(0, gy)
    x1 = trunc((pd[:, 0] + radius + TILE - 1) / TILE).clamp(0, gx)
    y1 = trunc((pd[:, 1] + radius + TILE - 1) / TILE).clamp(0, gy)
    visible = (tz > 0.2) & (det != 0) & ((x1 - x0) * (y1 - y0) > 0)
    return dict(depth=tz, pix=pix, conic=conic, op=opac * coef, rgb=rgb, rect=(x0, y0, x1, y1), visible=visible)


def blend(sp, H, W, bg):
    """front-to-back blend with tile culling, alpha < 1/255 skip, T < 1e-4 stop -> [4,H,W]."""
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    tyi, txi = (ys / TILE).long(), (xs / TILE).long()
    vis = sp["visible"].nonzero().flatten()
    dbits = sp["depth"].detach()[vis].view(torch.int32).long()
    order = vis[torch.argsort(dbits * (2 ** 20) + vis, stable=True)]     # (depth bits, id) ascending
    T = torch.ones(H, W)
    C = torch.zeros(3, H, W)
    done = torch.zeros(H, W, dtype=torch.bool)
    x0, y0, x1, y1 = sp["rect"]
    for g in order.tolist():
        in_tile = (txi >= x0[g]) & (txi < x1[g]) & (tyi >= y0[g]) & (tyi < y1[g])
        dx, dy = sp["pix"][g, 0] - xs, sp["pix"][g, 1] - ys
        A, B, Cc = sp["conic"][g]
        power = -0.5 * (A * dx * dx + Cc * dy * dy) - B * dx * dy
        G = torch.exp(power)
        a_raw = sp["op"][g] * G
        alpha = a_raw - torch.clamp(a_raw - 0.99, min=0).detach()       # min(0.99, .) with pass-through gradient
        valid = in_tile & (power <= 0) & (alpha >= 1.0 / 255.0) & ~done
        test_T = T * (1 - alpha)
        stop = valid & (test_T < 1e-4)
        done = done | stop
        valid = valid & ~stop
        C = C + torch.where(valid, alpha * T, torch.zeros_like(T))[None] * sp["rgb"][g][:, None, None]
        T = torch.where(valid, test_T, T)
    bgt = torch.tensor(bg, dtype=torch.float32)
    return torch.cat([C + T[None] * bgt[:, None, None], (1 - T)[None]], 0)


def render(canon, delta, const, view_t, proj_t, H, W, tanfovx, tanfovy, bg=(1.0, 1.0, 1.0), kernel_size=0.1, mip_filter=True):
    m3, sc, rt, sh, op = activate(canon, delta, const)
    sp = project(m3, sc, rt, sh, op, view_t, proj_t, H, W, tanfovx, tanfovy, kernel_size, mip_filter=mip_filter)
    return blend(sp, H, W, bg)
