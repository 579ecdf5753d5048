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
  int32_t mip_filter;         /* 1: mip-splatting 2-D filter (cov2D += kernel_size I, opacity *= sqrt(det0 / det1)), the
                                 `pipe.use_mip_gaussian = True` rasteriser of renderers/gaussian_render.py:103-125;
                                 0: plain 3DGS dilation (cov2D += kernel_size I, usually 0.3, no opacity compensation) --
                                 the `diff_gauss` rasteriser of :126-141, used by the alignment pre-step
                                 (utils/inference_utils.py:50) */
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
  GVF_RB_DSPLAT,         /* float[F*P*12] backward accumulator: d(px,py,conic a,b,c,opacity',r,g,b) */
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

/* The reference's visualisation loop renders every timestep from 128 orbit cameras
 * (utils/inference_utils.py:243-269: 32 x 128 calls of renderer.render per object).  Same as
 * gvf_raster_forward(activated = 0) for F = timesteps x views_per_delta frames ordered timestep-major,
 * with delta [F / views_per_delta, P, 14]: the views of one timestep read the same delta rows (no
 * replicated copies), cams stays per frame [F,32]. */
GVF_API int gvf_raster_forward_views(const gvf_raster_params* prm, int F, int P, int views_per_delta,
                       const float* xyz, const float* dc, const float* scaling,
                       const float* rotation, const float* opacity, const float* delta,
                       const float* cams, const float* subpixel_offset, float* out_rgba,
                       int32_t* radii, void* workspace, size_t workspace_bytes, int64_t cap,
                       void* stream);
/* Output stage of the same loop (utils/inference_utils.py:278-283): (clamp(rgb, 0, 1) * 255).astype(uint8),
 * planar fp32 rgba [F,4,H,W] -> interleaved uint8 [F,H,W,3] on the device (a quarter of the bytes to copy back). */
GVF_API int gvf_rgba_to_u8(const float* rgba, int F, int H, int W, uint8_t* out, void* stream);

/* Rest of the output stage (utils/inference_utils.py:284-297): `Image.fromarray(rgb).resize((t, t), LANCZOS)` followed by the
 * centre pad (white) / centre crop back to 512^2, on interleaved uint8 frames [F, H, W, 3] that stay on the device.
 * gvf_resample_u8 is Pillow's separable 8-bit resampler: bounds_* int32 [out, 2] = (first input index, tap count), coef_*
 * int32 [out, ksize] = taps in fixed point with 22 fractional bits (gvfdiffusion_b200/utils/inference_utils.py:
 * pil_resample_coeffs computes them in double like Pillow's precompute_coeffs); horizontal pass into tmp [F, Hin, Wout, 3],
 * then vertical into out [F, Hout, Wout, 3].  Byte-identical to PIL. */
GVF_API int gvf_resample_u8(const uint8_t* in, int F, int Hin, int Win, int Hout, int Wout, const int* bounds_h,
                            const int* coef_h, int ksize_h, const int* bounds_v, const int* coef_v, int ksize_v,
                            uint8_t* tmp, uint8_t* out, void* stream);
GVF_API int gvf_pad_crop_u8(const uint8_t* in, int F, int Hin, int Win, int S, int fill, uint8_t* out, void* stream);

/* Backward -- replaces GaussianRasterizer.backward (reached through autograd from reference
 * train_vae.py:313-352) fused with the backward of GaussianModel.get_*_with_delta.  Must follow a
 * gvf_raster_forward call with the same arguments and workspace (it reads the splat records, sorted
 * lists, final_T and n_contrib left there).  dL_drgba: [F,4,H,W].
 *  activated == 0: g_xyz[P,3], g_dc[P,3], g_scaling[P,3], g_rotation[P,4], g_opacity[P] receive the
 *     gradients of the RAW canonical tensors summed over frames; g_delta [F,P,14] (or NULL) those of delta.
 *  activated == 1: the five outputs are per-frame gradients of the activated inputs ([F,P,..]).
 *  g_means2D: [F,P,2] or NULL, screen-space gradient in upstream's convention (d/d ndc). */
GVF_API int gvf_raster_backward(const gvf_raster_params* prm, int F, int P, int activated,
                                const float* xyz, const float* dc, const float* scaling,
                                const float* rotation, const float* opacity, const float* delta,
                                const float* cams, const float* subpixel_offset, const float* dL_drgba,
                                void* workspace, size_t workspace_bytes, int64_t cap, float* g_xyz,
                                float* g_dc, float* g_scaling, float* g_rotation, float* g_opacity,
                                float* g_delta, float* g_means2D, void* stream);

/* get_gaussian_tensor (reference train_vae.py:466-472): raw canonical GaussianModel tensors ->
 * activated [P,14] = [xyz3 | rgb3 | opacity1 | scale3 | rot4] (decoder queries, static_latent). */
GVF_API int gvf_gaussian_tensor(const gvf_raster_params* prm, int P, const float* xyz, const float* dc,
                                const float* scaling, const float* rotation, const float* opacity,
                                float* out, void* stream);
/* Backward of gvf_gaussian_tensor: g [P, 14] -> gradients of the raw tensors (train_vae.py:285-293 back-propagates the
 * deformation losses into the static VAE through the activated Gaussians it hands to the motion VAE as queries). */
GVF_API int gvf_gaussian_tensor_bwd(const gvf_raster_params* prm, int P, const float* scaling, const float* rotation,
                                    const float* opacity, const float* g, float* d_xyz, float* d_dc, float* d_scaling,
                                    float* d_rotation, float* d_opacity, void* stream);
/* Farthest point sampling of K of P points (rows of `ld` floats, xyz first) -- replaces
 * torch_cluster.fps in sample_gs (reference utils/inference_utils.py:180-198).  Deterministic
 * (starts at `start`, ties -> lowest index).  workspace: P floats.  out_idx: K int32. */
GVF_API int gvf_fps(const float* pts, int ld, int P, int K, int start, float* workspace,
                    int32_t* out_idx, void* stream);
/* The same indices, bit for bit, through the exactly pruned kernel: for clouds whose rows are spatially ordered (the
 * voxel-major Gaussians of lexicographically ordered voxels that sample_gs receives from to_representation).  Threads skip
 * runs of consecutive points whose bounding box is farther from the new sample than their largest running minimum; the
 * bound uses the distance's own rounded operations, so skipping is exact in floating point.  Any row order is correct;
 * unordered clouds are faster through gvf_fps. */
GVF_API int gvf_fps_ordered(const float* pts, int ld, int P, int K, int start, float* workspace,
                            int32_t* out_idx, void* stream);

/* ------------------------------------------------------------------------------------
 * 2. Dense attention -- replaces flash_attn.flash_attn_{func,kvpacked_func,qkvpacked_func}
 *    as dispatched by reference model/attention/full_attn.py:74-140
 *    (scaled_dot_product_attention, layout [N, L, H, d]) and called directly at
 *    model/autoencoder.py:132-144.  fp16 in/out, fp32 statistics, no mask, no dropout.
 *    q [Nb, Lq, H, D], k / v [Nb, Lk, H, D], o [Nb, Lq, H, D]; *_strides = element strides
 *    {batch, sequence, head} (innermost contiguous, all multiples of 8), so packed qkv / kv
 *    tensors and transposed (temporal) views are consumed in place.  q_shared / kv_shared:
 *    the tensor has no batch dimension and is reused by every batch entry.
 *    D in {32, 64}.  Lq <= 32 self-attention (the temporal attention of model/dit.py:254-260)
 *    runs on CUDA cores; everything else on tcgen05 tensor cores fed by TMA.
 * ---------------------------------------------------------------------------------- */
GVF_API int gvf_attn_fwd_f16(const void* q, const void* k, const void* v, void* o, int Nb, int Lq,
                             int Lk, int H, int D, const long long* q_strides,
                             const long long* k_strides, const long long* v_strides,
                             const long long* o_strides, int q_shared, int kv_shared, float scale,
                             void* stream);

/* Timing-experiment hook (tools/attn_experiments.py): disables parts of the softmax loop; results are
 * numerically meaningless when non-zero.  0 = normal operation. */
GVF_API void gvf_attn_set_debug(int v);
/* Optional device buffer (256 int64) receiving clock64 stamps of CTA (0,0,0)'s first 16 blocks; NULL = off. */
GVF_API void gvf_attn_set_trace(void* device_buffer);

/* ------------------------------------------------------------------------------------
 * 3. Linear layers -- replace nn.Linear (cuBLAS under fp16 autocast) plus the elementwise
 *    kernels around it (reference model/dit.py:128-138,240-277, model/attention/modules.py:
 *    98-146, model/autoencoder.py:90-163).  out = epilogue(A[M,K] * W[N,K]^T + bias).
 *    A, W fp16 row-major (lda/ldw in elements, multiples of 8); tcgen05 + TMA, fp32 accumulate.
 *    epilogue: 0 fp16 store | 1 GELU(tanh) then fp16 store | 2 fp32 residual:
 *    out[m,n] += fp16(gate[m / rows_per_batch, n] * fp16(acc + bias)) (gate optional, fp16)
 *    | 3 fp16 residual: out = fp16(fp16(acc + bias) + out) | 4 fp32 store
 *    | 5 fp32 store of the fp16-rounded result into compact rows of ldo <= N columns (W rows
 *    padded to a multiple of 8, e.g. the VAE's Linear(768 -> 14))
 *    | 8 GELU(tanh) backward: out = fp16(fp16(acc) * gelu'(gate[m, n])), gate = the saved fp16 pre-activation
 *    [M, gate_stride] (the dgrad of an MLP's second Linear with the activation's derivative applied in place).
 *    Epilogue 1 with gate != NULL additionally stores the fp16 pre-activation there (training forward).
 * ---------------------------------------------------------------------------------- */
GVF_API int gvf_gemm_f16(const void* A, int lda, const void* W, int ldw, int M, int N, int K,
                         int epilogue, const float* bias, void* out, int ldo, const void* gate,
                         int gate_stride, int rows_per_batch, void* stream);

/* Caller-owned scratch for the dense attention kernel (optional): with it, the (batch, head) units of the last,
 * partially filled CTA wave are cut into three key ranges whose partial results are merged by a second small
 * kernel (csrc/attn.cu: attn_merge_kernel).  Needs 3 * 512 * 34 * 4 bytes per unit of the last wave (< 148
 * units); without it (or with too small a buffer) every unit runs whole.  The buffer must stay valid while
 * attention launches are in flight; pass NULL to unregister. */
GVF_API void gvf_attn_set_workspace(void* ws, size_t bytes);

/* Benchmark tuning hook: tile scheduling variant of gvf_gemm_f16 (-1 automatic; 0 one 128x128 tile per CTA,
 * 1 / 2 persistent 128x128 / 128x256, 3 three CTAs per SM, 4 / 5 generation 2 (eight epilogue warps, TMA stores)
 * 128x128 / 128x256, 6 / 7 generation 2 on CTA pairs (cta_group::2) 256x128 / 256x256). */
GVF_API void gvf_gemm_set_variant(int v);
/* Split-K of the fp32-store epilogue (4): 0 automatic (few output tiles and a long reduction: the wgrad shapes of the
 * training step; partial tiles are summed by TMA reduce-add stores, summation order not fixed), -1 never, n > 0 forced. */
GVF_API void gvf_gemm_set_ksplit(int k);

/* Programmatic dependent launch for the GEMM / attention / LayerNorm kernels (default OFF: measured slower
 * under CUDA-graph replay, see csrc/launch.h): the next kernel's prologue overlaps the previous grid's drain;
 * data dependences are unchanged (griddepcontrol.wait before the first global access). */
GVF_API void gvf_set_pdl(int on);

/* Self-attention QKV projection with MultiHeadRMSNorm fused into the epilogue (reference
 * model/attention/modules.py:113-125): out fp16 [M,N]; columns [0, norm_cols) are 32-wide heads,
 * the first half normalised with gamma_q [norm_cols/64, 32], the second half with gamma_k;
 * columns >= norm_cols (v) are stored as is. */
GVF_API int gvf_gemm_qkv_rmsnorm_f16(const void* A, int lda, const void* W, int ldw, int M, int N, int K,
                                     const float* bias, void* out, int ldo, const float* gamma_q,
                                     const float* gamma_k, int norm_cols, void* stream);

/* FeedForward's first Linear with GEGLU fused into the epilogue (reference model/autoencoder.py:90-107:
 * `x, gates = net[0](x).chunk(2, -1); x * F.gelu(gates)`): out fp16 [M, N/2] = fp16(fp16(value) * fp16(gelu_erf(fp16(gate)))).
 * W [N, K] / bias [N] are the rows of net.0 interleaved per 256-row tile -- rows [256 t, 256 t + 128) = value rows
 * [128 t, 128 t + 128), rows [256 t + 128, 256 t + 256) = gate rows [N/2 + 128 t, N/2 + 128 t + 128) -- so that value and
 * gate of an output column meet in one accumulator tile.  N % 256 == 0, else GVF_ERR_UNSUPPORTED (caller uses
 * gvf_gemm_f16 + gvf_geglu_f16). */
GVF_API int gvf_gemm_geglu_f16(const void* A, int lda, const void* W, int ldw, int M, int N, int K,
                               const float* bias, void* out, int ldo, void* stream);

/* Residual Linear fused with the LayerNorm (+ adaLN modulate or affine) of the next sub-block (reference
 * model/dit.py:246-277: `x = x + gate * attn_out_proj(...)` then `h = norm(x) * (1 + scale) + shift`):
 *   x[M,512] += gate * fp16(A W^T + b)   (fp32, in place);   y[M,512] = fp16(LN(x) * (1 + scale) + shift)
 * or y = fp16(LN(x) * ln_w + ln_b), or plain LN when all four are NULL.  N must be 512 (one CTA owns whole
 * rows); other widths return GVF_ERR_UNSUPPORTED and the caller uses gvf_gemm_f16 + gvf_ln_mod_f16. */
GVF_API int gvf_gemm_resid_ln_f16(const void* A, int lda, const void* W, int ldw, int M, int N, int K,
                                  const float* bias, float* x, int ldx, const void* gate, int gate_stride,
                                  int rows_per_batch, const float* ln_w, const float* ln_b, const void* shift,
                                  const void* scale, int mod_stride, float eps, void* y, int ldy, void* stream);

/* y = fp16(x[M,K] W[N,K]^T + b) (+ add[m % add_rows, n], fp32) for K <= 32 on CUDA cores:
 * input_layer + APE (model/dit.py:457,470-472), static_cond_proj (:465), VAE proj
 * (model/autoencoder.py:585), gs_embedding (:389).  x fp32, W fp16, out fp32 or fp16. */
GVF_API int gvf_small_linear(const float* x, int ldx, const void* W, const float* bias, int M, int N,
                             int K, const float* add, int add_rows, void* out, int out_is_f16,
                             void* stream);

/* ------------------------------------------------------------------------------------
 * 4. Fused elementwise kernels of the DiT / VAE blocks.
 * ---------------------------------------------------------------------------------- */
/* LayerNorm(eps; optional affine w,b) then optional adaLN modulate y*(1+scale[b])+shift[b]
 * (shift/scale fp16 [batches, mod_stride]) -> fp16.  x fp32 or fp16 [M,C].
 * reference model/dit.py:246-247,254-255,263,268,273-274; model/autoencoder.py:73-88. */
GVF_API int gvf_ln_mod_f16(const void* x, int x_is_f16, void* out, int M, int C, float eps,
                           const float* w, const float* b, const void* shift, const void* scale,
                           int mod_stride, int rows_per_batch, void* stream);
/* Tuning hook for A/B runs: 1 = widths 512 / 768 / 1024 with >= 2048 rows keep two rows per warp in flight (identical
 * bits; measured slower than the one-row-per-warp kernel, so the default is 0). */
GVF_API void gvf_ln_set_two_rows(int on);
/* Same with an activation applied before the single fp16 rounding: act 0 none, 1 SiLU -- the `norm1 -> SiLU -> conv1` and
 * `norm2 * (1 + scale) + shift -> SiLU -> conv2` pairs of SparseResBlock3d (trellis/models/structured_latent_flow.py:57-62).
 * Widths 64 / 128 / 256 / 1024 / 2048 for act 1. */
GVF_API int gvf_ln_mod_act_f16(const void* x, int x_is_f16, void* out, int M, int C, float eps,
                               const float* w, const float* b, const void* shift, const void* scale,
                               int mod_stride, int rows_per_batch, int act, void* stream);
/* MultiHeadRMSNorm on q and k, in place (model/attention/modules.py:8-15,122-125). */
GVF_API int gvf_rmsnorm_heads_f16(void* buf, long long rows, int ld, int H, int D, int k_off,
                                  const float* gamma_q, const float* gamma_k, void* stream);
/* TimestepEmbedder + every adaLN modulation vector of one forward (model/dit.py:59-100,
 * 240-242,299): t[B] -> temb, silu(temb) [B,C] fp16 and mod_out [B,R] fp16 where the R rows of
 * Wmod are the blocks' adaLN_modulation.1 / adaLN_modulation_temporal.1 and the final
 * layer's adaLN_modulation.1 concatenated. */
GVF_API int gvf_dit_modulation(const float* t, int B, int C, int F, const void* W0, const float* b0,
                               const void* W2, const float* b2, const void* Wmod, const float* bmod,
                               int R, void* temb, void* silu_temb, void* mod_out, void* stream);
/* AbsolutePositionEmbedder (model/dit.py:16-56): xyz [R,3] -> [R,C] fp32 */
GVF_API int gvf_ape(const float* xyz, int R, int C, float* out, void* stream);
/* VAE query embedding LN(LN(gs_embedding(q)) + LN(PointEmbed(q.xyz))) -> fp16
 * (model/autoencoder.py:250-301,389-391,560; PreNorm of :562). */
GVF_API int gvf_vae_query_embed(const float* queries, int ldq, const void* gs, int Q, int C, void* out,
                                void* stream);
/* Token embedding of the motion-VAE ENCODER (model/autoencoder.py:529-533): out fp32 [R, C] = LN_1e-5(lin[r]) +
 * LN_1e-5(PointEmbed(xyz[xyz_row[r]])) without the PreNorm that gvf_vae_query_embed applies on top (the sum enters a
 * residual stream first); lin fp16 [R, C] = input_embedding's Linear(3 -> C) of the per-frame displacement; xyz_row int32
 * [R] (or NULL: row r) lets the T frames of a point share its position row. */
GVF_API int gvf_vae_embed_sum(const float* xyz, int ldq, const int* xyz_row, const void* lin, int R, int C, void* out,
                              void* stream);
/* DiagonalGaussianDistribution (model/autoencoder.py:304-326): sample = mean + exp(clamp(logvar, -30, 20) / 2) * noise,
 * kl[b] = 0.5 * mean_b(mean^2 + var - 1 - logvar); fp32 [B, per_batch]; sample / kl / noise may be NULL. */
GVF_API int gvf_diag_gaussian(const float* mean, const float* logvar, const float* noise, int B, long long per_batch,
                              float* sample, float* kl, void* stream);
GVF_API int gvf_diag_gaussian_bwd(const float* mean, const float* logvar, const float* noise, const float* dsample,
                                  const float* dkl, int B, long long per_batch, float* dmean, float* dlogvar, void* stream);
/* GEGLU (model/autoencoder.py:90-93): h [M,2F] fp16 -> [M,F] */
GVF_API int gvf_geglu_f16(const void* h, long long M, int F, void* out, void* stream);
GVF_API int gvf_cast_f32_f16(const float* x, long long n, void* out, void* stream);
/* FinalLayer (model/dit.py:287-303): LN -> modulate -> Linear(C -> O) ; out fp32 [M,O] */
GVF_API int gvf_dit_final_layer(const float* x, int M, int C, int O, const void* shift, const void* scale,
                                int mod_stride, int rows_per_batch, const void* W, const float* bias,
                                float* out, void* stream);

/* ------------------------------------------------------------------------------------
 * 5. DPM-Solver++ state kernels -- replace the ~15 elementwise torch launches per step of
 *    reference model/dpmsolver.py:284-300 (v -> eps), :328-347 (3-way CFG), :450-459 (x0),
 *    :589-597 / :843-848 (first / second order multistep update).
 * ---------------------------------------------------------------------------------- */
/* model_type 1: the model predicts v (eps = alpha v + sigma x); 0: it predicts eps. */
GVF_API int gvf_dpm_x0(const float* x, const float* v, long long n, int branches, int model_type,
                       float alpha, float sigma, float s1, float s2, float* x0, void* stream);
/* adaptive solver error (model/dpmsolver.py:1016-1018): per batch b,
 * E2[b] += sum(((xh - xl) / max(atol, rtol max(|xl|, |xprev|)))^2); caller zeroes E2, then
 * E = max_b sqrt(E2[b] / n_per_batch). */
GVF_API int gvf_dpm_error_sq(const float* x_higher, const float* x_lower, const float* x_prev, int B,
                             long long n_per_batch, float atol, float rtol, float* E2, void* stream);
GVF_API int gvf_dpm_update(const float* x, const float* m0, const float* m1, long long n, float cx, float cm,
                           float inv_r0, int order, float* out, void* stream);
/* Euler step of the TRELLIS flow-matching samplers (trellis/pipelines/samplers/flow_euler.py:36-77 with the guidance
 * mixins): v = (1 + cfg) v - cfg v_neg when v_neg != NULL; x_prev = x - (t - t_prev) v; x0 (optional) =
 * (1 - sigma_min) x - (sigma_min + (1 - sigma_min) t) v.  Bit-identical to the reference's torch expressions. */
GVF_API int gvf_flow_euler_step(const float* x, const float* v, const float* v_neg, long long n, double cfg_strength, double t,
                                double t_prev, double sigma_min, float* x_prev, float* x0, void* stream);
/* out = x * a[c] + b[c] over the last dim (or scalars as/bs when a/b are NULL):
 * latent de-normalisation, inference_dpm_latent.py:250 */
GVF_API int gvf_affine_lastdim(const float* x, long long n, int C, const float* a, const float* b, float as,
                               float bs, float* out, void* stream);

/* ------------------------------------------------------------------------------------
 * 6. vox2seq -- replaces the reference's first-party extension
 *    model/sparse_voxel_diffusion/vox2seq/src/{api,z_order,hilbert}.cu (z_order_encode/decode,
 *    hilbert_encode/decode; called from sparse/attention/serialized_attn.py:67-74).
 *    coords int32 [N,3], codes int32 [N] (10 bits per axis); permute = axis order fed to the curve
 *    (vox2seq/__init__.py:17-19); hilbert 0 = Morton / z-order, 1 = Hilbert.
 * ---------------------------------------------------------------------------------- */
GVF_API int gvf_vox2seq_encode(const int32_t* coords, long long N, const int* permute, int hilbert,
                               int32_t* codes, void* stream);
GVF_API int gvf_vox2seq_decode(const int32_t* codes, long long N, const int* permute, int hilbert,
                               int32_t* coords, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Windowed self-attention over sparse voxels (static-VAE swin blocks).
 * Replaces sparse_windowed_scaled_dot_product_self_attention (sparse/attention/windowed_attn.py:61-135):
 * `qkv.feats[fwd_indices]` -> flash_attn_varlen_qkvpacked_func -> `out[bwd_indices]`.  The gather through
 * fwd_idx is done by the kernel's staging loads and every result row is stored at out[fwd_idx[i]], so
 * neither permutation copy exists.  qkv [T,3,H,64] fp16, out [T,H,64] fp16, fwd_idx [M] int32 (rows ordered
 * by window), cu_seqlens [W+1] int32, max_seqlen >= the longest window.  Head dim 64 only. */
GVF_API int gvf_sparse_window_attn_f16(const void* qkv, void* out, const int* fwd_idx, const int* cu_seqlens,
                                       int num_windows, int max_seqlen, int H, int D, float scale, void* stream);

/* General form of the kernel above: variable-length self-attention over rows of a packed qkv [T, 3, H, 64] tensor.
 * gather_idx [M] int32 or NULL (position i reads qkv row i: sparse full attention, sparse/attention/full_attn.py:90-215,
 * cu_seqlens = the batch layout); scatter_idx [M] int32 or NULL: destination row of position i's result, negative =
 * dropped (serialized attention, sparse/attention/serialized_attn.py:38-192, pads every window to window_size with
 * wrapped-around neighbours and keeps only the valid part via bwd_indices); NULL = the row the query came from. */
GVF_API int gvf_sparse_varlen_attn_f16(const void* qkv, void* out, const int* gather_idx, const int* scatter_idx,
                                       const int* cu_seqlens, int num_seqs, int max_seqlen, int H, int D, float scale,
                                       void* stream);

/* Training forward of the kernel above (also leaves LSE2 [T, H] fp32 per voxel row) and its backward for bijective lists
 * (windowed / full attention; the padded windows of serialized attention are forward-only): o [T, H, 64] the forward
 * output, dout its gradient, dsum [T, H] scratch, dqkv [T, 3, H, 64] receives dq | dk | dv in voxel order.  Replaces the
 * backward of flash_attn_varlen_qkvpacked_func (sparse/attention/windowed_attn.py:125-127 under autograd, cfg 5). */
GVF_API int gvf_sparse_varlen_attn_lse_f16(const void* qkv, void* out, float* lse2, const int* gather_idx,
                                           const int* scatter_idx, const int* cu_seqlens, int num_seqs, int max_seqlen, int H,
                                           int D, float scale, void* stream);
GVF_API int gvf_sparse_varlen_attn_bwd_f16(const void* qkv, const void* o, const void* dout, const float* lse2, float* dsum,
                                           void* dqkv, const int* gather_idx, const int* cu_seqlens, int num_seqs,
                                           int max_seqlen, long long T, int H, int D, float scale, void* stream);
/* PACKED tiling of the forward (lse2 may be NULL) and backward above, for lists whose sequences are contiguous ranges of one
 * position list of M entries (window partitions): one CTA per 64 CONSECUTIVE positions and head -- several short windows
 * share a tile, rows are masked to their own window's key range -- instead of one CTA per (window, head, 64 rows), most of
 * whose rows are padding when windows hold ~16 voxels.  seq_of_pos [M] int32 = the sequence of position p. */
GVF_API int gvf_sparse_packed_attn_f16(const void* qkv, void* out, float* lse2, const int* gather_idx, const int* scatter_idx,
                                       const int* cu_seqlens, const int* seq_of_pos, int M, int H, int D, float scale,
                                       void* stream);
/* A/B switch of the packed forward's staging: 0 (default, 3-6 % faster) per-thread cp.async, 1 TMA tile::gather4 rows on an
 * mbarrier. */
GVF_API void gvf_sparse_attn_set_tma(int on);
GVF_API int gvf_sparse_packed_attn_bwd_f16(const void* qkv, const void* o, const void* dout, const float* lse2, float* dsum,
                                           void* dqkv, const int* gather_idx, const int* cu_seqlens, const int* seq_of_pos,
                                           int M, long long T, int H, int D, float scale, void* stream);

/* Native driver of a stack of un-modulated SparseTransformerBlocks (reference sparse_transformer.py:126-192 stacked at
 * sparse_transformer_vae.py:55-91; block i uses partition parts[i % 2] = un-shifted / shifted windows): the whole launch
 * sequence of the static VAE's encoder or decoder trunk behind one call (csrc/sparse_trunk.cu), so that a 4096-token training
 * step is not bound by per-launch host overhead.  T tokens, C = 64 H channels, F = MLP width; residual stream fp16
 * (use_fp16) or fp32.  All pointers are device pointers except `blocks` and `parts` (host arrays). */
typedef struct gvf_sparse_block {
  const void *w_qkv, *w_out, *w1, *w2;             /* fp16 [3C,C] ([3][H][d] rows), [C,C], [F,C], [C,F] */
  const float *b_qkv, *b_out, *b1, *b2;            /* fp32 */
  const void *w_qkv_t, *w_out_t, *w1_t, *w2_t;     /* unused since the dgrad GEMMs read the weights directly (gvf_gemm_nn_f16) */
  float *g_w_qkv, *g_b_qkv, *g_w_out, *g_b_out, *g_w1, *g_b1, *g_w2, *g_b2;   /* fp32 gradient outputs (backward only) */
} gvf_sparse_block;
typedef struct gvf_window_partition {
  const int* fwd_idx;                              /* [T] token rows ordered by window */
  const int* cu_seqlens;                           /* [num_windows + 1] */
  int num_windows, max_seqlen;
  const int* seq_of_pos;                           /* [T] window of sorted position p, or NULL: one CTA per window tile */
} gvf_window_partition;
GVF_API size_t gvf_sparse_trunk_arena_bytes(int T, int C, int H, int F, int num_blocks, int fp16_residual);
GVF_API size_t gvf_sparse_trunk_scratch_bytes(int T, int C, int H, int F);
/* x_out = blocks(x_in) ([T, C] in the residual dtype).  arena != NULL: training forward, every activation the backward
 * needs is kept there (gvf_sparse_trunk_arena_bytes); arena == NULL: inference, `scratch` holds the temporaries. */
GVF_API int gvf_sparse_trunk_forward(const gvf_sparse_block* blocks, int num_blocks, int T, int C, int H, int F,
                                     int fp16_residual, const gvf_window_partition* parts, const void* x_in, void* arena,
                                     size_t arena_bytes, void* scratch, size_t scratch_bytes, void* x_out, void* stream);
/* d_out fp16 [T, C] (gradient of x_out) -> d_in fp16 [T, C] and every g_* of `blocks`; reduce_ws: gvf_colsum scratch
 * for [T, 3C]. */
GVF_API int gvf_sparse_trunk_backward(const gvf_sparse_block* blocks, int num_blocks, int T, int C, int H, int F,
                                      int fp16_residual, const gvf_window_partition* parts, const void* arena,
                                      size_t arena_bytes, const void* d_out, void* scratch, size_t scratch_bytes,
                                      float* reduce_ws, size_t reduce_ws_bytes, void* d_in, void* stream);

/* ------------------------------------------------------------------------------------
 * 7. Training-step losses (BASELINE configs[4]).
 *    gvf_ssim_l1_*: nn.L1Loss + utils/loss_util.py:33-63 `ssim` as used at train_vae.py:328-330
 *    (11x11 Gaussian window sigma 1.5, zero padding, C1 = 0.01^2, C2 = 0.03^2) in one kernel per
 *    direction instead of five depthwise convolutions + ~20 elementwise launches and their autograd.
 *    img1 (prediction) / img2 (target): fp32 [planes, H, W] (planes = B * C).  Forward writes
 *    sums [planes][2] = (sum of the SSIM map, sum of |img1 - img2|) -- the caller divides by the
 *    element counts (size_average True / False are both host-side reductions of these) -- and, when
 *    dmaps != NULL, the three per-pixel partial derivatives [3][planes][H][W] backward needs.
 *    Backward: grad_img1 = coef_l1[plane] sign(img1 - img2) + coef_ssim[plane] dSSIMsum/dimg1,
 *    coef_* fp32 [planes] on the device (upstream gradient / element count).
 * ---------------------------------------------------------------------------------- */
GVF_API size_t gvf_ssim_l1_workspace_bytes(int planes, int H, int W);
GVF_API int gvf_ssim_l1_fwd(const float* img1, const float* img2, int planes, int H, int W, float* workspace,
                            size_t workspace_bytes, float* sums, float* dmaps, void* stream);
GVF_API int gvf_ssim_l1_bwd(const float* img1, const float* img2, const float* dmaps, int planes, int H, int W,
                            const float* coef_ssim, const float* coef_l1, float* grad_img1, void* stream);
/* Exact K nearest neighbours (K <= 16): replaces pytorch3d.ops.knn_points(p1, p2, lengths1, lengths2, K)
 * as called at train_vae.py:525-530.  queries [B,P1,3], refs [B,P2,3] fp32; lengths int64 [B] or NULL;
 * dists [B,P1,K] squared distances ascending ((dx*dx + dy*dy) + dz*dz, every operation rounded: no FMA),
 * idx [B,P1,K] int64, ties -> lowest index; rows >= lengths1 and columns >= lengths2 are zero. */
/* LPIPS tail of one VGG16 feature tap (utils/lpips/lpips.py:29-34): d[n] = mean_pixels sum_c w_c (fx_c / (||fx|| + 1e-10) -
 * fy_c / (||fy|| + 1e-10))^2 over channels-last fp16 activations fx, fy [N, HW, C], C in {64, 128, 256, 512}; one pass
 * forward (block partial sums [N, gvf_lpips_tap_blocks(HW)], d[n] = their sum / HW), one pass backward (gradient of fx). */
GVF_API int gvf_lpips_tap_blocks(int HW);
/* Glue between the criterion's cuDNN convolutions on channels-last fp16 activations: x[p, c] = max(x[p, c] + bias[c], 0) in
 * place (C % 8 == 0), and the 2 x 2 / stride 2 max pool with its backward (first maximum of a window gets the gradient). */
GVF_API int gvf_bias_relu_nhwc_f16(void* x, const float* bias, long long pixels, int C, void* stream);
GVF_API int gvf_maxpool2_nhwc_f16(const void* x, void* y, int N, int H, int W, int C, void* stream);
GVF_API int gvf_maxpool2_nhwc_bwd_f16(const void* x, const void* y, const void* gy, void* gx, int N, int H, int W, int C,
                                      void* stream);
GVF_API int gvf_lpips_tap_fwd(const void* fx, const void* fy, const float* w, int N, int HW, int C, float* partial, void* stream);
GVF_API int gvf_lpips_tap_bwd(const void* fx, const void* fy, const float* w, const float* gout, int N, int HW, int C, void* gfx,
                              void* stream);
GVF_API int gvf_knn(const float* queries, const float* refs, int B, int P1, int P2, const long long* lengths1,
                    const long long* lengths2, int K, float* dists, long long* idx, void* stream);
/* compute_interpolation_loss_delta_interp's neighbour-motion estimate (train_vae.py:532-563):
 * radius = sqrt(mean_k dists) + 1e-6; w_k = exp(-beta d_k / radius^2) [d_k <= radius^2] (or exp(-beta d_k)),
 * zero for padded rows, normalised by (sum + 1e-8); est [B,T,P1,3] = sum_k w_k (moving[b,t,idx_k] - static[b,idx_k]). */
GVF_API int gvf_knn_interp_deltas(const float* dists, const long long* idx, const float* static_pc,
                                  const float* moving_pc, const long long* lengths1, int B, int P1, int P2, int T,
                                  int K, int adaptive_radius, float beta, float* est, void* stream);

/* ------------------------------------------------------------------------------------
 * 7. Voxel-side operators of the static (canonical Gaussian) VAE and of the TRELLIS stage in front of it.
 *
 *    gvf_to_representation replaces SparseVAE.to_representation (reference
 *    model/sparse_voxel_diffusion/sparse_vae.py:114-180; layout of a feature row :211-227):
 *    feats fp32 [nvox, ldf >= 14 G] = (_xyz (G,3) | _features_dc (G,1,3) | _scaling (G,3) | _rotation (G,4) |
 *    _opacity (G,1)), coords int32 [nvox, 4] = (batch, x, y, z), perturbation fp32 [G, 3] or NULL
 *    (perturb_offset false), lr HOST float[5] in the order (_xyz, _features_dc, _scaling, _rotation, _opacity);
 *    reg_mode 0 none | 1 invoxel tanh(o) / res | 2 soft_invoxel tanh(o) / res * 0.5 * voxel_size.
 *    Outputs are the raw GaussianModel tensors of P = nvox * G Gaussians (voxel-major, Gaussian-minor, exactly
 *    the reference's flatten(0, 1)): xyz [P,3], features_dc [P,3], scaling [P,3], rotation [P,4], opacity [P].
 *
 *    Submanifold sparse convolution (sparse/conv/conv_spconv.py:6-15 -> spconv.SubMConv3d, callers
 *    trellis/models/structured_latent_flow.py:34-35, structured_latent_vae/decoder_mesh.py:43-52):
 *    gvf_sparse_neighbor_map builds nbr int32 [N, ksize^3] (row index of the voxel at
 *    coords[i] + dilation * (k - ksize/2) per axis, k = (kx * ksize + ky) * ksize + kz, -1 where empty) through a
 *    dense [B, D, D, D] int32 grid in the caller's workspace (gvf_sparse_conv_workspace_bytes); *status
 *    (device int, optional) gets bit 0 for out-of-range coordinates, bit 1 for duplicates.
 *    gvf_sparse_im2col_f16 gathers x (fp16 or fp32 [N, ldx]) into the fp16 [N, K3 * Cin] GEMM operand
 *    (zeros where nbr < 0); the contraction itself is gvf_gemm_f16 with W [Cout, K3 * Cin] (spconv's
 *    [Cout, kx, ky, kz, Cin] weight, flattened) and any of its epilogues.
 * ---------------------------------------------------------------------------------- */
GVF_API int gvf_to_representation(const float* feats, int ldf, const int* coords, int nvox, int G,
                                  const float* perturbation, const float* lr /* host */, float resolution,
                                  int reg_mode, float voxel_size, float* xyz, float* features_dc, float* scaling,
                                  float* rotation, float* opacity, void* stream);
/* Backward of gvf_to_representation: gradients of the raw GaussianModel tensors (any may be NULL) -> g_feats [nvox, ldf]. */
GVF_API int gvf_to_representation_bwd(const float* feats, int ldf, int nvox, int G, const float* perturbation,
                                      const float* lr /* host */, float resolution, int reg_mode, float voxel_size,
                                      const float* g_xyz, const float* g_dc, const float* g_scaling, const float* g_rotation,
                                      const float* g_opacity, float* g_feats, void* stream);
GVF_API size_t gvf_sparse_conv_workspace_bytes(int B, int D);
GVF_API int gvf_sparse_neighbor_map(const int* coords, int N, int B, int D, int ksize, int dilation, void* workspace,
                                    size_t workspace_bytes, int* nbr, int* status, void* stream);
GVF_API int gvf_sparse_im2col_f16(const void* x, int x_is_f16, int ldx, const int* nbr, int N, int K3, int Cin,
                                  void* out, void* stream);
/* SparseDownsample (trellis/modules/sparse/spatial.py:13-52): out[p] = (sum of the rows of coarse cell p) / (count + 1)
 * -- the reference's scatter_reduce(zeros, 'mean') counts its zero initial value (include_self).  order int32 [N] = the
 * fine rows grouped by cell, offsets int32 [cells + 1]; x, out fp16; C % 8 == 0. */
GVF_API int gvf_sparse_pool_mean_f16(const void* x, int ldx, const int* order, const int* offsets, int cells, int C,
                                     void* out, int ldo, void* stream);
/* SparseUpsample (spatial.py:55-80: `input.feats[idx]`) and / or the skip concatenation of the flow model's output blocks
 * (structured_latent_flow.py:253-256): out[i] = [ a[idx ? idx[i] : i, 0:Ca] | b[i, 0:Cb] ], fp16, either part may be empty. */
GVF_API int gvf_gather_concat_f16(const void* a, int lda, int Ca, const int* idx, const void* b, int ldb, int Cb, int rows,
                                  void* out, int ldo, void* stream);
/* Submanifold convolution of a nearest-neighbour upsampled tensor without materialising it (the first output block of the
 * flow model, structured_latent_flow.py:166-172 -> SparseUpsample then SparseConv3d): the per-tap products
 * P[c, k * Cout + o] = sum_i W[o, k, i] a[c, i] are computed once per COARSE row by gvf_gemm_f16 (fp32 store), then
 *   out[n, o] = fp16(bias[o] + sum_k P[idx[nbr[n, k]], k * Cout + o])
 * with nbr int32 [N, K3] the fine level's neighbour map (-1 = absent) and idx int32 [N] the coarse cell of every fine row
 * (NULL: P is indexed by the neighbour row itself).  Cout % 4 == 0. */
GVF_API int gvf_sparse_tap_gather_sum_f16(const float* P, long long ldp, const int* nbr, const int* idx, int N, int K3, int Cout,
                                          const float* bias, void* out, int ldo, void* stream);

/* ------------------------------------------------------------------------------------
 * 8. Training step of the motion VAE (SURVEY.md rows g / a9 backward; BASELINE configs[2] and [4]).
 *    What torch autograd derives for reference model/autoencoder.py:552-609 (`decode`) under train_vae.py:293-353:
 *    attention backward on tcgen05, the element-wise / skinny backward passes, and the operand transposes that let
 *    gvf_gemm_f16 compute dgrad (dX = dY W: A = dY, "W" = W^T) and wgrad (dW = dY^T X: A = dY^T, "W" = X^T, fp32
 *    output, epilogue 4).
 * ---------------------------------------------------------------------------------- */
/* Forward attention that also leaves LSE2[Nb, H, lse_ld] = log2(sum_k exp(scale s_qk)) (fp32; lse_ld = Lq rounded up
 * to a multiple of 128, rows >= Lq set to +inf) for gvf_attn_bwd_f16.  Otherwise identical to gvf_attn_fwd_f16. */
GVF_API int gvf_attn_fwd_lse_f16(const void* q, const void* k, const void* v, void* o, float* lse2, int lse_ld, int Nb,
                                 int Lq, int Lk, int H, int D, const long long* q_strides, const long long* k_strides,
                                 const long long* v_strides, const long long* o_strides, int q_shared, int kv_shared,
                                 float scale, void* stream);
/* Backward of flash_attn_func (model/autoencoder.py:132-144): dq / dk / dv fp16 from q, k, v, o, dout and LSE2.
 * q_shared: q [Lq, H, D] is shared by the Nb batch entries (the decoder's frame-independent queries) and dq is the
 * sum over them; k / v / o / dout always carry the batch.  dsum: scratch [Nb, H, lse_ld] fp32 (row sums of dout * o).
 * Every *_strides = element strides {batch, sequence, head}.  D in {32, 64}. */
GVF_API int gvf_attn_bwd_f16(const void* q, const void* k, const void* v, const void* o, const void* dout,
                             const float* lse2, float* dsum, void* dq, void* dk, void* dv, int Nb, int Lq, int Lk, int H,
                             int D, int lse_ld, const long long* q_strides, const long long* k_strides,
                             const long long* v_strides, const long long* o_strides, const long long* do_strides,
                             const long long* dq_strides, const long long* dk_strides, const long long* dv_strides,
                             int q_shared, float scale, void* stream);
/* A/B hook of the attention backward kernels: 1 = the MMA issuer waits for the P / dS consumers of a block before it
 * overwrites their TMEM columns with the next block's scores (default 0: ordered by the tensor pipe itself). */
GVF_API void gvf_attn_bwd_set_serial(int v);
/* Weight gradient of a Linear without transposed copies: out[M, N] fp32 = A[R, M]^T W[R, N] (A = dY, W = X, fp16
 * row-major with row strides lda / ldw; dW[out, in] = dY^T X, reference autograd of nn.Linear under train_vae.py:352).
 * Both operands enter tcgen05.mma MN-major straight from 64 x 64 TMA boxes; split-K over R, partial tiles summed by
 * TMA reduce-add.  M, N multiples of 8. */
GVF_API int gvf_gemm_tn_f16(const void* A, int lda, const void* W, int ldw, int M, int N, int R, float* out, int ldo,
                            void* stream);
/* Input gradient of a Linear without a transposed weight copy: out[M, N] fp16 = epilogue(A[M, R] W[R, N]) with W row-major
 * [R, N] = the Linear's own [out_features, in_features] weight (MN-major B operand).  epilogue 0 or 8 (GELU', see section 3). */
GVF_API int gvf_gemm_nn_f16(const void* A, int lda, const void* W, int ldw, int M, int N, int R, int epilogue, void* out, int ldo,
                            const void* gate, int gate_stride, void* stream);
/* out[C, ld_out] = in[R, C]^T (fp16); columns [R, ld_out) of every output row are zero (ld_out = R rounded up to 8
 * so that the transposed tensor is a legal GEMM operand). */
GVF_API int gvf_transpose_f16(const void* in, int R, int C, long long ld_in, void* out, long long ld_out, void* stream);
/* Bytes of scratch for gvf_colsum (K = 0) / gvf_skinny_outer (K rows). */
GVF_API size_t gvf_colsum_workspace_bytes(long long M, int N, int K);
/* out[n] (+)= sum_m x[m, n]: bias gradients.  x fp16 or fp32 [M, N] with row stride ld. */
GVF_API int gvf_colsum(const void* x, int x_is_f16, long long M, int N, long long ld, float* workspace,
                       size_t workspace_bytes, float* out, int accumulate, void* stream);
/* LayerNorm (no affine, PreNorm model/autoencoder.py:73-88) backward: dx = dLN(dy; x) + dres (dres optional), fp16. */
GVF_API int gvf_ln_bwd_f16(const void* x, int x_is_f16, const void* dy, const void* dres, void* dx, int M, int C, float eps,
                           void* stream);
/* GEGLU backward (model/autoencoder.py:90-93): h [M, 2F], dG [M, F] -> dh [M, 2F] (all fp16). */
GVF_API int gvf_geglu_bwd_f16(const void* h, const void* dG, long long M, int F, void* dh, void* stream);
/* GELU(tanh) of the sparse trunk's MLP as its own pass (training keeps the pre-activation) and its backward; n % 8 == 0. */
GVF_API int gvf_gelu_tanh_f16(const void* h, long long n, void* out, void* stream);
GVF_API int gvf_gelu_tanh_bwd_f16(const void* h, const void* dy, long long n, void* dh, void* stream);
/* Backward of gvf_small_linear with respect to its input: dx[M, K] = dy[M, N] W[N, K] (fp32 out, K <= 32). */
GVF_API int gvf_small_linear_bwd_input(const void* dy, int dy_is_f16, long long ld, const void* W, long long M, int N, int K,
                                       float* dx, int ldx, void* stream);
/* out[K, N] (+)= x[M, K]^T y[M, N] for K <= 16 (fp32 x; y fp16 or fp32): weight gradients of the K <= 16 Linears
 * (transposed: proj / gs_embedding, x = layer input, y = dy) and of to_outputs (x = d out, y = layer input). */
GVF_API int gvf_skinny_outer(const float* x, int ldx, int K, const void* y, int y_is_f16, long long ldy, long long M, int N,
                             float* workspace, size_t workspace_bytes, float* out, int accumulate, void* stream);
/* out[M, N] (fp16 or fp32, row stride ldo) = x[M, K] Wt[K, N] for K <= 16, fp32 operands: rank-K expansions of the
 * decoder's output side (d attention-out = d out [M, 14] @ (W_to_outputs W_to_out) [14, 768]: to_outputs follows
 * decoder_cross_attn.to_out without a non-linearity, model/autoencoder.py:562-574, so their backward composes). */
GVF_API int gvf_skinny_expand(const float* x, int ldx, int K, const float* Wt, long long M, int N, void* out,
                              int out_is_f16, long long ldo, void* stream);
/* Backward of gvf_vae_query_embed: d out fp16 [Q, C] -> d gs fp16 [Q, C], d xyz fp32 (rows of ld_dxyz >= 3 floats,
 * e.g. the first three columns of d queries [Q, 14]; accumulate != 0 adds to what is there). */
GVF_API int gvf_vae_query_embed_bwd(const float* queries, int ldq, const void* gs, const void* dout, int Q, int C,
                                    void* dgs, float* dxyz, int ld_dxyz, int accumulate, void* stream);

/* The same convolution as ONE kernel: the GEMM's TMA producer gathers the neighbour rows itself
 * (cp.async.bulk.tensor tile::gather4 through nbr, absent neighbours zero-filled), so the [N, K3 * Cin] im2col operand is
 * never written.  x fp16 [N, Cin] (row stride ldx), W fp16 [Cout, K3 * Cin], out fp16 (epilogue 0; epilogue 3 = fp16
 * residual, out += fp16(conv + bias) in place: SparseResBlock3d's `h + skip_connection(x)`) or fp32 (4) [N, Cout].
 * Cin % 64 == 0, else GVF_ERR_UNSUPPORTED (callers fall back to im2col + gvf_gemm_f16).  Bit-identical to that path. */
GVF_API int gvf_sparse_conv_gemm_f16(const void* x, int ldx, const int* nbr, int N, int K3, int Cin, const void* W, int ldw,
                                     int Cout, const float* bias, void* out, int ldo, int epilogue, void* stream);

/* Tuning hook of the rasteriser's per-tile depth sort: -1 environment (GVF_RASTER_SORT=bucket|bitonic,
 * default bucket), 0 bitonic network, 1 one-pass bucket sort (identical point lists). */
GVF_API void gvf_raster_set_sort(int mode);

#ifdef __cplusplus
}
#endif
#endif /* GVF_B200_H_ */
