"""ctypes binding of libgvf_b200.so (the C ABI declared in include/gvf_b200.h).

There is deliberately NO fallback: if the shared object is missing or a call fails the
product raises.  PyTorch is used only for device memory and streams.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgvf_b200.so")
_lib = None


class GvfError(RuntimeError):
    pass


class RasterParams(C.Structure):
    _fields_ = [("H", C.c_int32), ("W", C.c_int32), ("tanfovx", C.c_float), ("tanfovy", C.c_float),
                ("kernel_size", C.c_float), ("scale_modifier", C.c_float), ("bg", C.c_float * 3),
                ("aabb", C.c_float * 6), ("scale_bias", C.c_float), ("min_kernel", C.c_float),
                ("opacity_bias", C.c_float), ("softplus", C.c_int32), ("mip_filter", C.c_int32)]


_P = C.c_void_p


class SparseBlock(C.Structure):
    """gvf_sparse_block (include/gvf_b200.h): weights, transposes and gradient outputs of one SparseTransformerBlock."""
    _fields_ = [(n, C.c_void_p) for n in (
        "w_qkv", "w_out", "w1", "w2", "b_qkv", "b_out", "b1", "b2", "w_qkv_t", "w_out_t", "w1_t", "w2_t",
        "g_w_qkv", "g_b_qkv", "g_w_out", "g_b_out", "g_w1", "g_b1", "g_w2", "g_b2")]


class WindowPartition(C.Structure):
    _fields_ = [("fwd_idx", C.c_void_p), ("cu_seqlens", C.c_void_p), ("num_windows", C.c_int), ("max_seqlen", C.c_int),
                ("seq_of_pos", C.c_void_p)]
_SIGS = {
    # name: (restype, argtypes)
    "gvf_status_string": (C.c_char_p, [C.c_int]),
    "gvf_abi_version": (C.c_int, []),
    "gvf_raster_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64]),
    "gvf_raster_workspace_offset": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64]),
    "gvf_raster_forward": (C.c_int, [C.POINTER(RasterParams), C.c_int, C.c_int, C.c_int,
                                     _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_size_t, C.c_int64, _P]),
    "gvf_raster_forward_views": (C.c_int, [C.POINTER(RasterParams), C.c_int, C.c_int, C.c_int,
                                           _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_size_t, C.c_int64, _P]),
    "gvf_rgba_to_u8": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "gvf_resample_u8": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, C.c_int, _P, _P, C.c_int, _P, _P, _P]),
    "gvf_pad_crop_u8": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "gvf_raster_backward": (C.c_int, [C.POINTER(RasterParams), C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, _P,
                                      _P, _P, _P, C.c_size_t, C.c_int64, _P, _P, _P, _P, _P, _P, _P, _P]),
    "gvf_gaussian_tensor": (C.c_int, [C.POINTER(RasterParams), C.c_int, _P, _P, _P, _P, _P, _P, _P]),
    "gvf_fps": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "gvf_fps_ordered": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "gvf_vox2seq_encode": (C.c_int, [_P, C.c_longlong, C.POINTER(C.c_int), C.c_int, _P, _P]),
    "gvf_vox2seq_decode": (C.c_int, [_P, C.c_longlong, C.POINTER(C.c_int), C.c_int, _P, _P]),
    "gvf_attn_fwd_f16": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong),
                                   C.POINTER(C.c_longlong), C.c_int, C.c_int, C.c_float, _P]),
    "gvf_attn_fwd_lse_f16": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong),
                                       C.POINTER(C.c_longlong), C.c_int, C.c_int, C.c_float, _P]),
    "gvf_attn_bwd_f16": (C.c_int, [_P] * 10 + [C.c_int] * 6 + [C.POINTER(C.c_longlong)] * 8 + [C.c_int, C.c_float, _P]),
    "gvf_transpose_f16": (C.c_int, [_P, C.c_int, C.c_int, C.c_longlong, _P, C.c_longlong, _P]),
    "gvf_colsum_workspace_bytes": (C.c_size_t, [C.c_longlong, C.c_int, C.c_int]),
    "gvf_colsum": (C.c_int, [_P, C.c_int, C.c_longlong, C.c_int, C.c_longlong, _P, C.c_size_t, _P, C.c_int, _P]),
    "gvf_ln_bwd_f16": (C.c_int, [_P, C.c_int, _P, _P, _P, C.c_int, C.c_int, C.c_float, _P]),
    "gvf_geglu_bwd_f16": (C.c_int, [_P, _P, C.c_longlong, C.c_int, _P, _P]),
    "gvf_sparse_attn_set_tma": (None, [C.c_int]),
    "gvf_sparse_packed_attn_f16": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_float, _P]),
    "gvf_sparse_packed_attn_bwd_f16": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_longlong, C.c_int, C.c_int,
                                                 C.c_float, _P]),
    "gvf_sparse_trunk_arena_bytes": (C.c_size_t, [C.c_int] * 6),
    "gvf_sparse_trunk_scratch_bytes": (C.c_size_t, [C.c_int] * 4),
    "gvf_sparse_trunk_forward": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, C.c_size_t,
                                           _P, C.c_size_t, _P, _P]),
    "gvf_sparse_trunk_backward": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, C.c_size_t, _P,
                                            _P, C.c_size_t, _P, C.c_size_t, _P, _P]),
    "gvf_lpips_tap_blocks": (C.c_int, [C.c_int]),
    "gvf_bias_relu_nhwc_f16": (C.c_int, [_P, _P, C.c_longlong, C.c_int, _P]),
    "gvf_maxpool2_nhwc_f16": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "gvf_maxpool2_nhwc_bwd_f16": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "gvf_lpips_tap_fwd": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "gvf_lpips_tap_bwd": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "gvf_gaussian_tensor_bwd": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "gvf_gemm_nn_f16": (C.c_int, [_P, C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int, _P, C.c_int, _P]),
    "gvf_gelu_tanh_f16": (C.c_int, [_P, C.c_longlong, _P, _P]),
    "gvf_gelu_tanh_bwd_f16": (C.c_int, [_P, _P, C.c_longlong, _P, _P]),
    "gvf_to_representation_bwd": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, C.POINTER(C.c_float), C.c_float, C.c_int,
                                            C.c_float, _P, _P, _P, _P, _P, _P, _P]),
    "gvf_small_linear_bwd_input": (C.c_int, [_P, C.c_int, C.c_longlong, _P, C.c_longlong, C.c_int, C.c_int, _P, C.c_int, _P]),
    "gvf_skinny_outer": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_int, C.c_longlong, C.c_longlong, C.c_int, _P, C.c_size_t,
                                   _P, C.c_int, _P]),
    "gvf_skinny_expand": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_longlong, C.c_int, _P, C.c_int, C.c_longlong, _P]),
    "gvf_vae_query_embed_bwd": (C.c_int, [_P, C.c_int, _P, _P, C.c_int, C.c_int, _P, _P, C.c_int, C.c_int, _P]),
    "gvf_gemm_f16": (C.c_int, [_P, C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, C.c_int,
                               _P, C.c_int, C.c_int, _P]),
    "gvf_gemm_geglu_f16": (C.c_int, [_P, C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, C.c_int, _P]),
    "gvf_gemm_qkv_rmsnorm_f16": (C.c_int, [_P, C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, C.c_int,
                                           _P, _P, C.c_int, _P]),
    "gvf_gemm_resid_ln_f16": (C.c_int, [_P, C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, C.c_int, _P, C.c_int,
                                        C.c_int, _P, _P, _P, _P, C.c_int, C.c_float, _P, C.c_int, _P]),
    "gvf_sparse_window_attn_f16": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _P]),
    "gvf_sparse_varlen_attn_f16": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _P]),
    "gvf_sparse_varlen_attn_lse_f16": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _P]),
    "gvf_sparse_varlen_attn_bwd_f16": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_longlong, C.c_int,
                                                 C.c_int, C.c_float, _P]),
    "gvf_attn_set_debug": (None, [C.c_int]),
    "gvf_attn_set_workspace": (None, [_P, C.c_size_t]),
    "gvf_attn_set_trace": (None, [_P]),
    "gvf_gemm_tn_f16": (C.c_int, [_P, C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int, _P]),
    "gvf_attn_bwd_set_serial": (None, [C.c_int]),
    "gvf_gemm_set_variant": (None, [C.c_int]),
    "gvf_gemm_set_ksplit": (None, [C.c_int]),
    "gvf_set_pdl": (None, [C.c_int]),
    "gvf_small_linear": (C.c_int, [_P, C.c_int, _P, _P, C.c_int, C.c_int, C.c_int, _P, C.c_int, _P, C.c_int, _P]),
    "gvf_ln_mod_f16": (C.c_int, [_P, C.c_int, _P, C.c_int, C.c_int, C.c_float, _P, _P, _P, _P, C.c_int, C.c_int, _P]),
    "gvf_ln_set_two_rows": (None, [C.c_int]),
    "gvf_ln_mod_act_f16": (C.c_int, [_P, C.c_int, _P, C.c_int, C.c_int, C.c_float, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "gvf_sparse_pool_mean_f16": (C.c_int, [_P, C.c_int, _P, _P, C.c_int, C.c_int, _P, C.c_int, _P]),
    "gvf_gather_concat_f16": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, C.c_int, C.c_int, C.c_int, _P, C.c_int, _P]),
    "gvf_sparse_tap_gather_sum_f16": (C.c_int, [_P, C.c_longlong, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P, C.c_int, _P]),
    "gvf_rmsnorm_heads_f16": (C.c_int, [_P, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "gvf_dit_modulation": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, C.c_int, _P, _P, _P, _P]),
    "gvf_ape": (C.c_int, [_P, C.c_int, C.c_int, _P, _P]),
    "gvf_vae_query_embed": (C.c_int, [_P, C.c_int, _P, C.c_int, C.c_int, _P, _P]),
    "gvf_vae_embed_sum": (C.c_int, [_P, C.c_int, _P, _P, C.c_int, C.c_int, _P, _P]),
    "gvf_diag_gaussian": (C.c_int, [_P, _P, _P, C.c_int, C.c_longlong, _P, _P, _P]),
    "gvf_diag_gaussian_bwd": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_longlong, _P, _P, _P]),
    "gvf_geglu_f16": (C.c_int, [_P, C.c_longlong, C.c_int, _P, _P]),
    "gvf_cast_f32_f16": (C.c_int, [_P, C.c_longlong, _P, _P]),
    "gvf_dit_final_layer": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P, C.c_int, C.c_int, _P, _P, _P, _P]),
    "gvf_dpm_x0": (C.c_int, [_P, _P, C.c_longlong, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                             _P, _P]),
    "gvf_dpm_error_sq": (C.c_int, [_P, _P, _P, C.c_int, C.c_longlong, C.c_float, C.c_float, _P, _P]),
    "gvf_dpm_update": (C.c_int, [_P, _P, _P, C.c_longlong, C.c_float, C.c_float, C.c_float, C.c_int, _P, _P]),
    "gvf_ssim_l1_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "gvf_ssim_l1_fwd": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, C.c_size_t, _P, _P, _P]),
    "gvf_ssim_l1_bwd": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P]),
    "gvf_knn": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P, C.c_int, _P, _P, _P]),
    "gvf_knn_interp_deltas": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_float, _P, _P]),
    "gvf_to_representation": (C.c_int, [_P, C.c_int, _P, C.c_int, C.c_int, _P, C.POINTER(C.c_float), C.c_float, C.c_int,
                                        C.c_float, _P, _P, _P, _P, _P, _P]),
    "gvf_sparse_conv_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "gvf_sparse_neighbor_map": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_size_t, _P, _P, _P]),
    "gvf_sparse_im2col_f16": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "gvf_sparse_conv_gemm_f16": (C.c_int, [_P, C.c_int, _P, C.c_int, C.c_int, C.c_int, _P, C.c_int, C.c_int, _P, _P, C.c_int,
                                           C.c_int, _P]),
    "gvf_raster_set_sort": (None, [C.c_int]),
    "gvf_flow_euler_step": (C.c_int, [_P, _P, _P, C.c_longlong, C.c_double, C.c_double, C.c_double, C.c_double, _P, _P, _P]),
    "gvf_affine_lastdim": (C.c_int, [_P, C.c_longlong, C.c_int, _P, _P, C.c_float, C.c_float, _P, _P]),
}


def declared_symbols():
    return sorted(_SIGS)


def lib():
    """Load the library (building it first when nvcc is available and sources changed)."""
    global _lib
    if _lib is not None:
        return _lib
    # under torchrun several ranks import at once: never rebuild concurrently when a library is already there
    multi_rank = int(os.environ.get("WORLD_SIZE", "1") or 1) > 1 and os.path.exists(LIB_PATH)
    if os.environ.get("GVF_NO_AUTOBUILD") != "1" and not multi_rank:
        try:
            from . import build as _b
            _b.build()
        except Exception as e:  # stale/missing toolchain: fall through to the existence check
            if not os.path.exists(LIB_PATH):
                raise GvfError(f"libgvf_b200.so is missing and could not be built: {e}") from e
    if not os.path.exists(LIB_PATH):
        raise GvfError(f"{LIB_PATH} not found: run `python -m gvfdiffusion_b200.build` "
                       "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(status, what=""):
    if status != 0:
        msg = lib().gvf_status_string(status).decode()
        raise GvfError(f"{what} failed: {msg} ({status})")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


_raw_stream = None


def current_stream():
    """cudaStream_t of torch's current stream on the current device.  Through torch's raw accessor when it exists: building a
    torch.cuda.Stream object per launch (torch.cuda.current_stream()) costs ~6 us, a tenth of a launch-heavy training step."""
    global _raw_stream
    import torch
    if _raw_stream is None:
        _raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", False)
    if _raw_stream:
        return C.c_void_p(_raw_stream(torch.cuda.current_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
