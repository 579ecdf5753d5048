"""ctypes front-end of oracle/raster.c (TEST INFRASTRUCTURE, see oracle/__init__.py)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_raster.so")
_lib = None


class Params(C.Structure):
    _fields_ = [("H", C.c_int32), ("W", C.c_int32), ("tanfovx", C.c_float), ("tanfovy", C.c_float),
                ("kernel_size", C.c_float), ("scale_modifier", C.c_float), ("bg", C.c_float * 3),
                ("aabb", C.c_float * 6), ("scale_bias", C.c_float), ("min_kernel", C.c_float),
                ("opacity_bias", C.c_float), ("softplus", C.c_int32), ("mip_filter", C.c_int32)]


def build(force=False):
    src = os.path.join(_HERE, "raster.c")
    hdr = os.path.join(_HERE, "..", "include", "gvf_math.h")
    if (not force and os.path.exists(_SO)
            and os.path.getmtime(_SO) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return _SO
    subprocess.check_call(["make", "-C", _HERE, "-B"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.gvf_oracle_forward.restype = C.c_int64
        for n in ("expf", "logf", "log1pf", "softplusf", "sigmoidf"):
            f = getattr(_lib, "gvf_oracle_" + n)
            f.restype = C.c_float
            f.argtypes = [C.c_float]
    return _lib


def make_params(H, W, tanfovx, tanfovy, const, kernel_size=0.1, scale_modifier=1.0, bg=(1.0, 1.0, 1.0), mip_filter=True):
    p = Params()
    p.mip_filter = int(bool(mip_filter))
    p.H, p.W, p.tanfovx, p.tanfovy = H, W, tanfovx, tanfovy
    p.kernel_size, p.scale_modifier = kernel_size, scale_modifier
    p.bg = (C.c_float * 3)(*bg)
    p.aabb = (C.c_float * 6)(*const["aabb"])
    p.scale_bias, p.min_kernel, p.opacity_bias = const["scale_bias"], const["min_kernel"], const["opacity_bias"]
    p.softplus = int(const["softplus"])
    return p


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a, t=C.c_float):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def activate(prm, canon, delta):
    """canon: dict of numpy raw params; delta (P,14) or None."""
    P = canon["_xyz"].shape[0]
    arrs = [_f(canon[k]).reshape(P, -1) for k in ("_xyz", "_features_dc", "_scaling", "_rotation", "_opacity")]
    d = None if delta is None else _f(delta)
    out = [np.empty((P, n), np.float32) for n in (3, 3, 4, 3, 1)]
    lib().gvf_oracle_activate(C.byref(prm), P, *[_ptr(a) for a in arrs], _ptr(d), *[_ptr(o) for o in out])
    return out  # means3D, scales, rots, shs, opac


def forward(prm, means3D, scales, rots, shs, opac, view_t, proj_t, subpixel_offset=None, cap=None):
    """One frame.  view_t / proj_t are the transposed matrices the rasteriser receives."""
    P = means3D.shape[0]
    H, W = prm.H, prm.W
    nt = ((H + 15) // 16) * ((W + 15) // 16)
    cap = cap or max(1, 64 * P)
    o = {"rgba": np.empty((4, H, W), np.float32), "radii": np.empty(P, np.int32),
         "tiles_touched": np.empty(P, np.uint32), "splat": np.zeros((P, 10), np.float32),
         "rects": np.zeros((P, 4), np.int32), "ranges": np.zeros((nt, 2), np.uint32),
         "point_list": np.zeros(cap, np.uint32), "keys": np.zeros(cap, np.uint64),
         "n_contrib": np.zeros((H, W), np.uint32), "final_T": np.zeros((H, W), np.float32)}
    a = [_f(x) for x in (means3D, scales, rots, shs.reshape(P, 3), opac.reshape(P), view_t, proj_t)]
    so = None if subpixel_offset is None else _f(subpixel_offset)
    R = lib().gvf_oracle_forward(
        C.byref(prm), P, *[_ptr(x) for x in a], _ptr(so), _ptr(o["rgba"]), _ptr(o["radii"], C.c_int32),
        _ptr(o["tiles_touched"], C.c_uint32), _ptr(o["splat"]), _ptr(o["rects"], C.c_int32),
        _ptr(o["ranges"], C.c_uint32), _ptr(o["point_list"], C.c_uint32), _ptr(o["keys"], C.c_uint64),
        C.c_int64(cap), _ptr(o["n_contrib"], C.c_uint32), _ptr(o["final_T"]))
    if R < 0:   # capacity too small: retry with a larger list
        return forward(prm, means3D, scales, rots, shs, opac, view_t, proj_t, subpixel_offset, cap * 8)
    o["num_rendered"] = int(R)
    o["point_list"] = o["point_list"][:R]
    o["keys"] = o["keys"][:R]
    return o


def render_frames(prm, canon, delta, views_t, projs_t, want_radii=False):
    """F frames, canonical + per-frame delta (F,P,14) or None.  -> rgba (F,4,H,W), R (F,)"""
    F = views_t.shape[0]
    P = canon["_xyz"].shape[0]
    arrs = [_f(canon[k]).reshape(P, -1) for k in ("_xyz", "_features_dc", "_scaling", "_rotation", "_opacity")]
    d = None if delta is None else _f(delta)
    v, p = _f(views_t), _f(projs_t)
    out = np.empty((F, 4, prm.H, prm.W), np.float32)
    radii = np.empty((F, P), np.int32) if want_radii else None
    nr = np.zeros(F, np.int64)
    err = lib().gvf_oracle_render_frames(C.byref(prm), F, P, *[_ptr(a) for a in arrs], _ptr(d), _ptr(v),
                                         _ptr(p), _ptr(out), _ptr(radii, C.c_int32), _ptr(nr, C.c_int64))
    if err:
        raise RuntimeError("oracle raster failed")
    return (out, nr, radii) if want_radii else (out, nr)
