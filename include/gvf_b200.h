/* gvf_b200.h -- C ABI of libgvf_b200.so, the B200-native (sm_100a) kernels behind the
 * GVFDiffusion sampling + 4D Gaussian rendering hot path.
 *
 * The reference (ForeverFancy/GVFDiffusion) has no C ABI of its own: its "plugin
 * interface" for this path is a set of Python call signatures that end in third-party
 * CUDA extensions.  Each entry point below names the reference call site it replaces.
 * All pointers are DEVICE pointers unless marked host; all functions enqueue work on the
 * given CUDA stream (cudaStream_t passed as void*) and return without synchronising;
 * they allocate nothing (caller owns outputs and workspace) and keep no global state.
 * Return value: 0 on success, negative gvf_status on error (see gvf_status_string).
 *
 * Reference-side bindings (ctypes) are shown in INTEGRATION.md.
 */
#ifndef GVF_B200_H_
#define GVF_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define GVF_API __attribute__((visibility("default")))
#else
#define GVF_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

enum gvf_status {
  GVF_OK = 0,
  GVF_ERR_INVALID = -1,      /* bad argument (null pointer, unsupported shape) */
  GVF_ERR_WORKSPACE = -2,    /* workspace too small */
  GVF_ERR_CUDA = -3,         /* a CUDA runtime call / launch failed (cudaGetLastError) */
  GVF_ERR_UNSUPPORTED = -4   /* configuration not implemented on this path */
};
GVF_API const char* gvf_status_string(int status);
/* ABI version of this header; bumps on any signature change. */
GVF_API int gvf_abi_version(void);

/* ------------------------------------------------------------------------------------
 * 1. Rasteriser -- replaces diff_gaussian_rasterization.GaussianRasterizer.forward as
 *    called at reference renderers/gaussian_render.py:105-125,198-206 (and the identical
 *    renderers/gaussian_render_all_delta.py:93-125,195-203), fused with
 *    GaussianModel.get_*_with_delta (representations/gaussian/gaussian_model.py:98-114)
 *    and batched over F frames (one camera per frame: the loop of
 *    utils/inference_utils.py:256-269).
 * ---------------------------------------------------------------------------------- */
typedef struct gvf_raster_params {
  int32_t H, W;               /* image size; tiles are 16x16 */
  float tanfovx, tanfovy;     /* = 0.5 / focal (normalised intrinsics) */
  float kernel_size;          /* mip 2-D filter (pipe.kernel_size, 0.1) */
  float scale_modifier;
  float bg[3];
  /* GaussianModel constants, used only when activated == 0 */
  float aabb[6];              /* xyz = _xyz * aabb[3:6] + aabb[0:3] */
  float scale_bias;           /* inverse_softplus(scaling_bias) */
  float min_kernel;           /* mininum_kernel_size (3-D filter) */
  float opacity_bias;         /* logit(opacity_bias) */
  int32_t softplus;           /* 1 softplus, 0 exp scaling activation */
} gvf_raster_params;

/* Names of the sub-buffers inside the rasteriser workspace (for tests / backward). */
enum gvf_raster_buf {
  GVF_RB_SPLAT = 0,      /* float4[F*P*3]: (px,py,ca,cb) (cc,op,r,g) (b,depth,radius,tiles) */
  GVF_RB_RECT,           /* uint16[F*P*4]: tile rect x0,y0,x1,y1 */
  GVF_RB_TILE_COUNT,     /* uint32[F*T]   */
  GVF_RB_TILE_START,     /* uint32[F*T+1] exclusive scan; [F*T] = num_rendered */
  GVF_RB_KEYS,           /* uint64[cap]   (depth_bits<<32 | gaussian id), per-tile segments */
  GVF_RB_POINT_LIST,     /* uint32[cap]   gaussian ids in (frame, tile, depth, id) order */
  GVF_RB_FINAL_T,        /* float[F*H*W]  */
  GVF_RB_N_CONTRIB,      /* uint32[F*H*W] */
  GVF_RB_STATUS,         /* uint32[4]: num_rendered, overflow flag, max tile length, 0 */
  GVF_RB_SCAN_TMP,       /* uint32[...]   */
  GVF_RB_COUNT_
};

/* Bytes of workspace needed for F frames x P Gaussians with room for `cap` tile
 * instances (sum over frames of tiles touched). */
GVF_API size_t gvf_raster_workspace_bytes(int F, int P, int H, int W, int64_t cap);
/* Byte offset of a sub-buffer inside the workspace (same arguments). */
GVF_API size_t gvf_raster_workspace_offset(int which, int F, int P, int H, int W, int64_t cap);

/* Forward.
 *  activated == 0: xyz/dc/scaling/rotation/opacity are the RAW canonical GaussianModel
 *     tensors (_xyz[P,3], _features_dc[P,3], _scaling[P,3], _rotation[P,4], _opacity[P]),
 *     shared by all frames; delta is [F,P,14] = [xyz3|scale3|rot4|rgb3|opacity1] or NULL.
 *  activated == 1: the five arrays are per-frame ACTIVATED rasteriser inputs
 *     (means3D[F,P,3], shs[F,P,3], scales[F,P,3], rotations[F,P,4], opacities[F,P]); delta
 *     must be NULL.  This is the diff_gaussian_rasterization calling convention.
 *  cams: [F,32] = viewmatrix (view^T, 16 floats) then projmatrix ((P view)^T, 16 floats).
 *  subpixel_offset: [H,W,2] or NULL (zeros).
 *  out_rgba: [F,4,H,W] fp32, RGB composited over bg, A = 1 - T_final.
 *  radii: [F,P] int32 or NULL.
 * The number of tile instances is data dependent: status[0] holds it after the call and
 * status[1] != 0 reports that `cap` was too small (outputs are then incomplete). */
GVF_API int gvf_raster_forward(const gvf_raster_params* prm, int F, int P, int activated,
                       const float* xyz, const float* dc, const float* scaling,
                       const float* rotation, const float* opacity, const float* delta,
                       const float* cams, const float* subpixel_offset, float* out_rgba,
                       int32_t* radii, void* workspace, size_t workspace_bytes, int64_t cap,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GVF_B200_H_ */
