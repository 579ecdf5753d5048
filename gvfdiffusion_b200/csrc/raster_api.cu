// raster_api.cu -- C ABI of the rasteriser (include/gvf_b200.h section 1).
#include "raster_common.h"

using namespace gvf;

extern "C" {

const char* gvf_status_string(int s) {
  switch (s) {
    case GVF_OK: return "ok";
    case GVF_ERR_INVALID: return "invalid argument";
    case GVF_ERR_WORKSPACE: return "workspace too small";
    case GVF_ERR_CUDA: return "CUDA error";
    case GVF_ERR_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown status";
  }
}

int gvf_abi_version(void) { return 1; }

size_t gvf_raster_workspace_bytes(int F, int P, int H, int W, int64_t cap) {
  if (F <= 0 || P <= 0 || H <= 0 || W <= 0 || cap < 0) return 0;
  return raster_layout(F, P, H, W, cap).total;
}

size_t gvf_raster_workspace_offset(int which, int F, int P, int H, int W, int64_t cap) {
  if (which < 0 || which >= GVF_RB_COUNT_) return (size_t)-1;
  return raster_layout(F, P, H, W, cap).off[which];
}

static int raster_forward_impl(const gvf_raster_params* prm, int F, int P, int activated, int views,
                               const float* xyz, const float* dc, const float* scaling,
                               const float* rotation, const float* opacity, const float* delta,
                               const float* cams, const float* subpixel_offset, float* out_rgba,
                               int32_t* radii, void* workspace, size_t workspace_bytes, int64_t cap,
                               void* stream) {
  if (!prm || !xyz || !dc || !scaling || !rotation || !opacity || !cams || !out_rgba || !workspace)
    return GVF_ERR_INVALID;
  if (F <= 0 || P <= 0 || prm->H <= 0 || prm->W <= 0 || cap <= 0) return GVF_ERR_INVALID;
  if (activated && delta) return GVF_ERR_INVALID;
  if (views < 1 || F % views != 0 || (views > 1 && activated)) return GVF_ERR_INVALID;
  const int gx = (prm->W + GVF_TILE - 1) / GVF_TILE, gy = (prm->H + GVF_TILE - 1) / GVF_TILE;
  if (gx > 65535 || gy > 65535) return GVF_ERR_UNSUPPORTED;
  if ((long long)F * gx * gy >= (1ll << 31) || (long long)F * P >= (1ll << 31) || cap >= (1ll << 32))
    return GVF_ERR_UNSUPPORTED;
  const RasterLayout L = raster_layout(F, P, prm->H, prm->W, cap);
  if (workspace_bytes < L.total) return GVF_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  char* w = (char*)workspace;
  RasterWs ws;
  ws.splat = (float4*)(w + L.off[GVF_RB_SPLAT]);
  ws.rect = (ushort4*)(w + L.off[GVF_RB_RECT]);
  ws.tile_count = (uint32_t*)(w + L.off[GVF_RB_TILE_COUNT]);
  ws.tile_start = (uint32_t*)(w + L.off[GVF_RB_TILE_START]);
  ws.keys = (unsigned long long*)(w + L.off[GVF_RB_KEYS]);
  ws.point_list = (uint32_t*)(w + L.off[GVF_RB_POINT_LIST]);
  ws.final_T = (float*)(w + L.off[GVF_RB_FINAL_T]);
  ws.n_contrib = (uint32_t*)(w + L.off[GVF_RB_N_CONTRIB]);
  ws.status = (uint32_t*)(w + L.off[GVF_RB_STATUS]);
  ws.scan_tmp = (uint32_t*)(w + L.off[GVF_RB_SCAN_TMP]);

  const int FT = F * gx * gy;
  // tile_count, status and the scan ticket are contiguous-ish small buffers: clear them.
  if (cudaMemsetAsync(ws.tile_count, 0, (size_t)FT * sizeof(uint32_t), st) != cudaSuccess) return GVF_ERR_CUDA;
  if (cudaMemsetAsync(ws.status, 0, 4 * sizeof(uint32_t), st) != cudaSuccess) return GVF_ERR_CUDA;
  if (cudaMemsetAsync(ws.scan_tmp, 0, sizeof(uint32_t), st) != cudaSuccess) return GVF_ERR_CUDA;
  if (launch_preprocess(*prm, F, P, activated, xyz, dc, scaling, rotation, opacity, delta, cams, ws,
                        radii, st, views) != cudaSuccess) return GVF_ERR_CUDA;
  if (launch_scan(FT, ws, st) != cudaSuccess) return GVF_ERR_CUDA;
  if (launch_scatter(*prm, F, P, ws, cap, st) != cudaSuccess) return GVF_ERR_CUDA;
  if (launch_sort_blend(*prm, F, P, ws, cap, subpixel_offset, out_rgba, st) != cudaSuccess)
    return GVF_ERR_CUDA;
  return GVF_OK;
}

int gvf_raster_forward(const gvf_raster_params* prm, int F, int P, int activated,
                       const float* xyz, const float* dc, const float* scaling,
                       const float* rotation, const float* opacity, const float* delta,
                       const float* cams, const float* subpixel_offset, float* out_rgba,
                       int32_t* radii, void* workspace, size_t workspace_bytes, int64_t cap,
                       void* stream) {
  return raster_forward_impl(prm, F, P, activated, 1, xyz, dc, scaling, rotation, opacity, delta, cams,
                             subpixel_offset, out_rgba, radii, workspace, workspace_bytes, cap, stream);
}

int gvf_raster_forward_views(const gvf_raster_params* prm, int F, int P, int views_per_delta,
                             const float* xyz, const float* dc, const float* scaling,
                             const float* rotation, const float* opacity, const float* delta,
                             const float* cams, const float* subpixel_offset, float* out_rgba,
                             int32_t* radii, void* workspace, size_t workspace_bytes, int64_t cap,
                             void* stream) {
  return raster_forward_impl(prm, F, P, 0, views_per_delta, xyz, dc, scaling, rotation, opacity, delta, cams,
                             subpixel_offset, out_rgba, radii, workspace, workspace_bytes, cap, stream);
}

// (clamp(rgb, 0, 1) * 255).astype(uint8), planar fp32 -> interleaved HWC: utils/inference_utils.py:278-283
__global__ void __launch_bounds__(256) rgba_to_u8_kernel(const float* __restrict__ rgba, long long n_pix, long long HW,
                                                         uint8_t* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pix) return;
  const long long f = i / HW, p = i - f * HW;
  const float* src = rgba + f * 4 * HW + p;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = fminf(fmaxf(__ldg(src + c * HW), 0.0f), 1.0f);
    out[i * 3 + c] = (uint8_t)(__fmul_rn(v, 255.0f));      // truncation, like numpy's astype
  }
}

int gvf_rgba_to_u8(const float* rgba, int F, int H, int W, uint8_t* out, void* stream) {
  if (!rgba || !out || F <= 0 || H <= 0 || W <= 0) return GVF_ERR_INVALID;
  const long long HW = (long long)H * W, n = HW * F;
  rgba_to_u8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rgba, n, HW, out);
  return cudaGetLastError() == cudaSuccess ? GVF_OK : GVF_ERR_CUDA;
}

}  // extern "C"
