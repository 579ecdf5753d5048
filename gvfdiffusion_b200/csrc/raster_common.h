// raster_common.h -- shared declarations of the sm_100a tile rasteriser.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gvf_b200.h"

#define GVF_TILE 16
#define GVF_TILE_PIX 256

namespace gvf {

struct RasterWs {
  float4* splat;         // [F*P*3]
  ushort4* rect;         // [F*P]
  uint32_t* tile_count;  // [F*T]
  uint32_t* tile_start;  // [F*T+1]
  unsigned long long* keys;  // [cap]
  uint32_t* point_list;  // [cap]
  float* final_T;        // [F*H*W]
  uint32_t* n_contrib;   // [F*H*W]
  uint32_t* status;      // [4]
  uint32_t* scan_tmp;    // [chunks + 2]
};

constexpr int kScanChunk = 2048;  // tiles per scan CTA

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct RasterLayout {
  size_t off[GVF_RB_COUNT_];
  size_t total;
};

inline RasterLayout raster_layout(int F, int P, int H, int W, int64_t cap) {
  const size_t T = (size_t)((W + GVF_TILE - 1) / GVF_TILE) * ((H + GVF_TILE - 1) / GVF_TILE);
  const size_t FP = (size_t)F * P, FT = (size_t)F * T, HW = (size_t)H * W;
  const size_t chunks = (FT + kScanChunk - 1) / kScanChunk;
  size_t sz[GVF_RB_COUNT_];
  sz[GVF_RB_SPLAT] = FP * 3 * sizeof(float4);
  sz[GVF_RB_RECT] = FP * sizeof(ushort4);
  sz[GVF_RB_TILE_COUNT] = FT * sizeof(uint32_t);
  sz[GVF_RB_TILE_START] = (FT + 1) * sizeof(uint32_t);
  sz[GVF_RB_KEYS] = (size_t)cap * sizeof(unsigned long long);
  sz[GVF_RB_POINT_LIST] = (size_t)cap * sizeof(uint32_t);
  sz[GVF_RB_FINAL_T] = (size_t)F * HW * sizeof(float);
  sz[GVF_RB_N_CONTRIB] = (size_t)F * HW * sizeof(uint32_t);
  sz[GVF_RB_STATUS] = 4 * sizeof(uint32_t);
  sz[GVF_RB_SCAN_TMP] = (chunks + 2) * sizeof(uint32_t);
  sz[GVF_RB_DSPLAT] = FP * 12 * sizeof(float);
  RasterLayout L;
  size_t o = 0;
  for (int i = 0; i < GVF_RB_COUNT_; ++i) {
    L.off[i] = o;
    o += align_up(sz[i], 256);
  }
  L.total = o;
  return L;
}

// kernels' host launchers (each returns cudaGetLastError())
cudaError_t launch_preprocess(const gvf_raster_params& prm, int F, int P, int activated,
                              const float* xyz, const float* dc, const float* scaling,
                              const float* rotation, const float* opacity, const float* delta,
                              const float* cams, const RasterWs& ws, int32_t* radii,
                              cudaStream_t st, int views = 1);
cudaError_t launch_scan(int n_tiles_total, const RasterWs& ws, cudaStream_t st);
cudaError_t launch_scatter(const gvf_raster_params& prm, int F, int P, const RasterWs& ws,
                           int64_t cap, cudaStream_t st);
cudaError_t launch_sort_blend(const gvf_raster_params& prm, int F, int P, const RasterWs& ws,
                              int64_t cap, const float* subpixel_offset, float* out_rgba,
                              cudaStream_t st);

}  // namespace gvf
