"""Host side of the frame-batched canonical+delta Gaussian rasteriser.

Mirrors what the reference does per render call in `renderers/gaussian_render.py:85-238`
(`render`) and `:269-353` (`GaussianRenderer.render`): camera matrices, GaussianModel
activations with delta, one `diff_gaussian_rasterization` call -- but for F frames at once
and through the C ABI (`gvf_raster_forward`, include/gvf_b200.h).  torch supplies device
memory and the current stream only.
"""
import ctypes as C

import torch

from . import _lib

TILE = 16
RB = {"splat": 0, "rect": 1, "tile_count": 2, "tile_start": 3, "keys": 4, "point_list": 5,
      "final_T": 6, "n_contrib": 7, "status": 8, "scan_tmp": 9, "dsplat": 10}


def intrinsics_to_projection(intrinsics, near, far):
    """reference renderers/gaussian_render.py:57-82, batched: (..., 3, 3) -> (..., 4, 4)."""
    ret = torch.zeros(intrinsics.shape[:-2] + (4, 4), dtype=intrinsics.dtype, device=intrinsics.device)
    ret[..., 0, 0] = 2 * intrinsics[..., 0, 0]
    ret[..., 1, 1] = 2 * intrinsics[..., 1, 1]
    ret[..., 0, 2] = 2 * intrinsics[..., 0, 2] - 1
    ret[..., 1, 2] = -2 * intrinsics[..., 1, 2] + 1
    ret[..., 2, 2] = far / (far - near)
    ret[..., 2, 3] = near * far / (near - far)
    ret[..., 3, 2] = 1.0
    return ret


def pack_cameras(extrinsics, intrinsics, near, far):
    """(F,4,4) world->camera, (3,3)|(F,3,3) normalised intrinsics -> cams (F,32) =
    [view^T | (persp @ view)^T] exactly as `GaussianRenderer.render` hands them over
    (renderers/gaussian_render.py:302-321), plus tanfovx/tanfovy = 0.5 / focal (host floats
    taken from frame 0: the reference uses one intrinsics matrix per object)."""
    ext = extrinsics.float()
    if ext.dim() == 2:
        ext = ext[None]
    intr = intrinsics.float()
    if intr.dim() == 2:
        intr = intr[None].expand(ext.shape[0], 3, 3)
    persp = intrinsics_to_projection(intr, near, far)
    full = persp @ ext
    cams = torch.cat([ext.transpose(1, 2).reshape(-1, 16), full.transpose(1, 2).reshape(-1, 16)], dim=1)
    i0 = intr[0].detach().cpu()
    # tan(0.5 * 2 * atan(0.5 / f)) == 0.5 / f up to rounding; the reference goes through
    # atan/tan in fp32 (renderers/gaussian_render.py:307-308, 102-103)
    import math
    tanfovx = math.tan(float(2 * torch.atan(0.5 / i0[0, 0])) * 0.5)
    tanfovy = math.tan(float(2 * torch.atan(0.5 / i0[1, 1])) * 0.5)
    return cams.contiguous(), tanfovx, tanfovy


class _PinnedRing:
    """A fixed ring of small pinned host buffers for asynchronous copies.  Asking torch for a fresh pinned tensor per
    render (`pin_memory()`, `torch.empty(pin_memory=True)`) goes to cudaHostAlloc whenever the caching host allocator has no
    block whose last use has completed -- a device-wide synchronisation of several milliseconds; with ~100 small copies
    in flight per training step that made the joint train step bimodal (90 / 200 ms).  A slot is reused only after the event
    recorded behind its last copy has completed."""

    def __init__(self, slots=512, nbytes=8192):
        self.slots, self.nbytes, self.buf, self.events, self.next = slots, nbytes, None, [None] * slots, 0

    def take(self, nbytes):
        if nbytes > self.nbytes:
            return None, -1
        if self.buf is None:
            self.buf = torch.empty((self.slots, self.nbytes), dtype=torch.uint8).pin_memory()
        i = self.next
        self.next = (i + 1) % self.slots
        ev = self.events[i]
        if ev is not None and not ev.query():
            ev.synchronize()
        return self.buf[i, :nbytes], i

    def mark(self, i):
        """Call after enqueueing the copy that uses slot i."""
        ev = self.events[i]
        if ev is None:
            ev = self.events[i] = torch.cuda.Event()
        ev.record()
        return ev


_RING = _PinnedRing()


def to_device_async(t, device):
    """Host tensor -> device through pinned memory without synchronising the stream (torch's blocking H2D copy waits for
    everything enqueued before it).  Device tensors pass through."""
    if t.is_cuda:
        return t
    t = t.contiguous()
    stage, slot = _RING.take(t.numel() * t.element_size())
    if stage is None:
        return t.pin_memory().to(device, non_blocking=True)
    host = stage.view(t.dtype).view(t.shape)
    host.copy_(t)
    out = host.to(device, non_blocking=True)
    _RING.mark(slot)
    return out


def make_params(H, W, tanfovx, tanfovy, const=None, kernel_size=0.1, scale_modifier=1.0,
                bg=(1.0, 1.0, 1.0), mip_filter=True):
    p = _lib.RasterParams()
    p.mip_filter = int(bool(mip_filter))
    p.H, p.W, p.tanfovx, p.tanfovy = int(H), int(W), float(tanfovx), float(tanfovy)
    p.kernel_size, p.scale_modifier = float(kernel_size), float(scale_modifier)
    p.bg = (C.c_float * 3)(*[float(b) for b in bg])
    if const is not None:
        p.aabb = (C.c_float * 6)(*const["aabb"])
        p.scale_bias, p.min_kernel = float(const["scale_bias"]), float(const["min_kernel"])
        p.opacity_bias, p.softplus = float(const["opacity_bias"]), int(const["softplus"])
    return p


class Rasterizer:
    """Owns the workspace for (F, P, H, W); grows `cap` when the tile-instance count needs it."""

    def __init__(self, device="cuda", tiles_per_gaussian=8):
        self.device = torch.device(device)
        self.tpg = tiles_per_gaussian
        self._key = None
        self._ws = None
        self.cap = 0
        self.hint = 0               # capacity learned by an earlier rasteriser of the same scene (see node())
        self._deferred = None       # (event, pinned int32[4]) of a forward(check_overflow="defer")

    def node(self):
        """A rasteriser with its OWN workspace for one autograd node: backward reads the splat records, sorted
        lists, final_T and n_contrib its forward left in the workspace (upstream returns them as the geom /
        binning / img buffers saved for backward), so a forward that will be differentiated must not share a
        workspace with the renders that follow it before `loss.backward()` (train_vae.py:313-334 renders every
        camera first)."""
        n = Rasterizer(self.device, self.tpg)
        n.hint = max(self.hint, self.cap)
        return n

    def _ensure(self, F, P, H, W, cap=None):
        cap = int(cap or max(self.cap if self._key == (F, P, H, W) else 0, self.hint, self.tpg * F * P, 1024))
        if self._key != (F, P, H, W) or cap != self.cap:
            nbytes = _lib.lib().gvf_raster_workspace_bytes(F, P, H, W, cap)
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._key, self.cap = (F, P, H, W), cap
        return self._ws

    def buffer(self, name, dtype, count=None):
        """View of a workspace sub-buffer (tests / backward)."""
        F, P, H, W = self._key
        L = _lib.lib()
        off = L.gvf_raster_workspace_offset(RB[name], F, P, H, W, self.cap)
        nxt = [L.gvf_raster_workspace_offset(i, F, P, H, W, self.cap) for i in range(len(RB))]
        ends = sorted(o for o in nxt if o > off) + [self._ws.numel()]
        v = self._ws[off:ends[0]].view(dtype)
        return v if count is None else v[:count]

    def status(self):
        """(num_rendered, overflow, longest global-sorted tile) -- synchronises."""
        s = self.buffer("status", torch.int32, 4).cpu()
        return int(s[0]) & 0xffffffff, bool(s[1]), int(s[2])

    def deferred_status(self):
        """(num_rendered, overflow) of the last forward(check_overflow="defer"), or None.  Waits for that
        forward's status copy only (an event), not for the stream."""
        if self._deferred is None:
            return None
        ev, host = self._deferred
        ev.synchronize()
        self._deferred = None
        R, overflow = int(host[0]) & 0xffffffff, bool(host[1])
        if overflow:                            # the NEXT forward gets a workspace that fits
            self.hint = max(self.hint, int(R * 1.25) + 1024)
        return R, overflow

    def forward(self, prm, arrays, delta, cams, activated=False, subpixel_offset=None,
                want_radii=True, out=None, check_overflow=True, views_per_delta=1):
        """arrays = (xyz, dc, scaling, rotation, opacity) fp32 contiguous device tensors.
        Returns rgba (F,4,H,W) fp32 and radii (F,P) int32 (or None).  views_per_delta = V > 1: the F frames
        are (timestep, camera) pairs ordered timestep-major and delta is [F / V, P, 14] (gvf_raster_forward_views).
        check_overflow: True = read the status word back (one host sync) and re-run with a larger workspace
        when the tile-instance capacity was exceeded (upstream sizes its buffers from the count, so it never
        drops a splat); "defer" = copy the status to pinned memory behind the kernels and let the caller ask
        `deferred_status()` after its own synchronisation point; False = unchecked (benchmarks of the kernels
        alone; the kernels clamp to `cap`, so memory stays safe but splats may be dropped)."""
        L = _lib.lib()
        F = cams.shape[0]
        P = arrays[0].shape[-2] if arrays[0].dim() >= 2 else arrays[0].shape[0]
        for t in tuple(arrays) + (cams,) + ((delta,) if delta is not None else ()):
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise ValueError("rasteriser inputs must be contiguous fp32 CUDA tensors")
        H, W = prm.H, prm.W
        while True:
            ws = self._ensure(F, P, H, W)
            rgba = out if out is not None else torch.empty((F, 4, H, W), dtype=torch.float32, device=self.device)
            radii = torch.empty((F, P), dtype=torch.int32, device=self.device) if want_radii else None
            if views_per_delta > 1:
                if delta is None or delta.shape[0] * views_per_delta != F or activated:
                    raise ValueError("views_per_delta: delta must be [F / views, P, 14] and inputs raw")
                st = L.gvf_raster_forward_views(C.byref(prm), F, P, int(views_per_delta), *[_lib.ptr(a) for a in arrays],
                                                _lib.ptr(delta), _lib.ptr(cams), _lib.ptr(subpixel_offset),
                                                _lib.ptr(rgba), _lib.ptr(radii), _lib.ptr(ws), ws.numel(),
                                                self.cap, _lib.current_stream())
            else:
                st = L.gvf_raster_forward(C.byref(prm), F, P, int(activated), *[_lib.ptr(a) for a in arrays],
                                          _lib.ptr(delta), _lib.ptr(cams), _lib.ptr(subpixel_offset),
                                          _lib.ptr(rgba), _lib.ptr(radii), _lib.ptr(ws), ws.numel(),
                                          self.cap, _lib.current_stream())
            _lib.check(st, "gvf_raster_forward")
            if not check_overflow:
                return rgba, radii
            if check_overflow == "defer":
                stage, slot = _RING.take(16)
                host = stage.view(torch.int32)
                host.copy_(self.buffer("status", torch.int32, 4), non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                _RING.mark(slot)
                self._deferred = (ev, host)
                return rgba, radii
            R, overflow, _ = self.status()
            if not overflow:
                return rgba, radii
            self._ensure(F, P, H, W, cap=int(R * 1.25) + 1024)


    def backward(self, prm, arrays, delta, cams, grad_rgba, activated=False, subpixel_offset=None,
                 want_delta_grad=True, want_means2D=False):
        """Backward of the last forward() with the same arguments.  Returns (five gradients of
        `arrays`, grad of delta or None, grad of means2D or None)."""
        L = _lib.lib()
        F, P, H, W = self._key
        dev = self.device
        g = grad_rgba.detach().to(dev, torch.float32).contiguous()
        lead = (F, P) if activated else (P,)
        outs = [torch.empty(lead + (n,), dtype=torch.float32, device=dev) for n in (3, 3, 3, 4)]
        outs.append(torch.empty(lead, dtype=torch.float32, device=dev))
        gdelta = torch.empty((F, P, 14), dtype=torch.float32, device=dev) if (want_delta_grad and not activated) else None
        gm2 = torch.empty((F, P, 2), dtype=torch.float32, device=dev) if want_means2D else None
        st = L.gvf_raster_backward(C.byref(prm), F, P, int(activated), *[_lib.ptr(a) for a in arrays],
                                   _lib.ptr(delta), _lib.ptr(cams), _lib.ptr(subpixel_offset), _lib.ptr(g),
                                   _lib.ptr(self._ws), self._ws.numel(), self.cap, *[_lib.ptr(o) for o in outs],
                                   _lib.ptr(gdelta), _lib.ptr(gm2), _lib.current_stream())
        _lib.check(st, "gvf_raster_backward")
        return outs, gdelta, gm2


def rgba_to_u8(rgba, out=None):
    """(clamp(rgb, 0, 1) * 255).astype(uint8): planar fp32 [F,4,H,W] -> [F,H,W,3] uint8 on the device
    (utils/inference_utils.py:278-283)."""
    if not (rgba.is_cuda and rgba.dtype == torch.float32 and rgba.is_contiguous() and rgba.dim() == 4 and rgba.shape[1] == 4):
        raise ValueError("rgba_to_u8: expected a contiguous CUDA fp32 [F,4,H,W] tensor")
    F, _, H, W = rgba.shape
    if out is None:
        out = torch.empty((F, H, W, 3), dtype=torch.uint8, device=rgba.device)
    _lib.check(_lib.lib().gvf_rgba_to_u8(_lib.ptr(rgba), F, H, W, _lib.ptr(out), _lib.current_stream()), "gvf_rgba_to_u8")
    return out


class RasterizeFrames(torch.autograd.Function):
    """autograd node: (raw canonical tensors, delta) -> RGBA for F frames (train_vae.py:313-334).

    The tile-instance overflow check of the forward is DEFERRED to the backward (one event wait there instead of a blocking
    status read-back per render, which would stop the host from enqueueing the rest of the step): if the forward did
    overflow its workspace, the backward re-renders into a workspace that fits -- its gradients are those of the complete
    image -- and warns that the image handed to the loss in this step was missing splats; the next forward starts from the
    grown capacity."""

    @staticmethod
    def forward(ctx, rz, prm, cams, xyz, dc, scaling, rotation, opacity, delta):
        arrays = tuple(t.detach().contiguous() for t in (xyz, dc, scaling, rotation, opacity))
        d = None if delta is None else delta.detach().contiguous()
        parent, rz = rz, rz.node()                 # per-node workspace, kept alive by ctx until backward
        rgba, radii = rz.forward(prm, arrays, d, cams, check_overflow="defer")
        ctx.parent = parent
        parent.hint = max(parent.hint, rz.cap)
        ctx.rz, ctx.prm, ctx.cams, ctx.arrays, ctx.delta = rz, prm, cams, arrays, d
        ctx.shapes = [t.shape for t in (xyz, dc, scaling, rotation, opacity)]
        ctx.mark_non_differentiable(radii)
        return rgba, radii

    @staticmethod
    def backward(ctx, g_rgba, _g_radii):
        st = ctx.rz.deferred_status()
        if st is not None and st[1]:
            import warnings
            warnings.warn(f"rasteriser workspace overflow in the forward of this step ({st[0]} tile instances > capacity "
                          f"{ctx.rz.cap}): re-rendered for the backward; the loss saw an image with splats missing")
            ctx.rz.forward(ctx.prm, ctx.arrays, ctx.delta, ctx.cams, check_overflow=True)
            ctx.parent.hint = max(ctx.parent.hint, ctx.rz.cap)
        outs, gdelta, _ = ctx.rz.backward(ctx.prm, ctx.arrays, ctx.delta, ctx.cams, g_rgba,
                                          want_delta_grad=ctx.delta is not None)
        outs = [o.reshape(s) for o, s in zip(outs, ctx.shapes)]
        return (None, None, None, *outs, gdelta)


class RasterizeActivated(torch.autograd.Function):
    """autograd node of the diff_gaussian_rasterization calling convention (one frame, ACTIVATED inputs):
    (means3D, means2D, dc, opacities, scales, rotations) -> (color [3,H,W], radii [P]).  means2D only receives
    the screen-space gradient (renderers/gaussian_render.py:96-100 reads `.grad` of a zero tensor)."""

    @staticmethod
    def forward(ctx, rz, prm, cams, sub, means3D, means2D, dc, opacities, scales, rotations):
        P = means3D.shape[0]
        f = lambda t, n: t.detach().to(torch.float32).reshape(1, P, n).contiguous()
        arrays = (f(means3D, 3), f(dc, 3), f(scales, 3), f(rotations, 4), f(opacities, 1).reshape(1, P))
        parent, rz = rz, rz.node()
        rgba, radii = rz.forward(prm, arrays, None, cams, activated=True, subpixel_offset=sub)
        parent.hint = max(parent.hint, rz.cap)
        ctx.rz, ctx.prm, ctx.cams, ctx.sub, ctx.arrays = rz, prm, cams, sub, arrays
        ctx.shapes = [t.shape for t in (means3D, dc, opacities, scales, rotations)]
        ctx.m2_shape = None if means2D is None else means2D.shape
        color, radii = rgba[0, :3], radii[0]
        ctx.mark_non_differentiable(radii)
        return color, radii

    @staticmethod
    def backward(ctx, g_color, _g_radii):
        H, W = ctx.prm.H, ctx.prm.W
        g = torch.zeros((1, 4, H, W), dtype=torch.float32, device=g_color.device)
        g[0, :3] = g_color
        want_m2 = ctx.m2_shape is not None and ctx.needs_input_grad[5]
        outs, _, gm2 = ctx.rz.backward(ctx.prm, ctx.arrays, None, ctx.cams, g, activated=True,
                                       subpixel_offset=ctx.sub, want_means2D=want_m2)
        g_m3, g_dc, g_sc, g_rot, g_op = (o[0] for o in outs)
        s = ctx.shapes
        g_m2 = None
        if want_m2:
            g_m2 = torch.zeros(ctx.m2_shape, dtype=torch.float32, device=g_color.device)
            g_m2[:, :2] = gm2[0]
        return (None, None, None, None, g_m3.reshape(s[0]), g_m2, g_dc.reshape(s[1]), g_op.reshape(s[2]),
                g_sc.reshape(s[3]), g_rot.reshape(s[4]))


def canon_arrays(canon, device):
    """GaussianModel raw tensors -> the five contiguous fp32 device arrays of the C ABI."""
    P = canon["_xyz"].shape[0]
    f = lambda t, n: t.detach().to(device=device, dtype=torch.float32).reshape(P, n).contiguous()
    return (f(canon["_xyz"], 3), f(canon["_features_dc"], 3), f(canon["_scaling"], 3),
            f(canon["_rotation"], 4), f(canon["_opacity"], 1).reshape(P).contiguous())
