// raster_blend.cu -- per-tile depth sort fused with the front-to-back alpha blend.
//
// Replaces `cub::DeviceRadixSort::SortPairs` + `identifyTileRanges` + `renderCUDA` of
// diff_gaussian_rasterization (mip-splatting fork; call site reference
// renderers/gaussian_render.py:198-206), batched over frames: one CTA per (frame, tile).
//
//   1. the tile's unordered key segment (depth bits << 32 | gaussian id) is pulled into
//      shared memory and sorted with an all-ascending bitonic network (works for any n:
//      missing partners act as +inf and never move); keys are unique, so the order is the
//      one a stable (tile, depth) radix sort produces.  Segments longer than kSortCap are
//      sorted in place in global memory by the same CTA (rare; correctness path).
//   2. sorted ids are written to point_list (for the backward pass and for parity tests).
//   3. 256 threads = 16x16 pixels blend front to back; splat records are gathered by id
//      into shared memory 256 at a time; the CTA stops when every pixel is saturated.
//
// No tensor cores on this path.  Blend math: alpha = min(.99, op * exp(power)), skip
// alpha < 1/255, stop when T(1-alpha) < 1e-4, out = C + T * bg, A = 1 - T.
#include "raster_common.h"

namespace gvf {

constexpr int kSortCap = 2048;

template <typename KeyPtr>
__device__ __forceinline__ void bitonic_sort_any_n(KeyPtr keys, int n) {
  // all-ascending bitonic network ("flip" then "disperse"), valid for arbitrary n
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  const int ncmp = np2 >> 1;
  for (int k = 2; k <= np2; k <<= 1) {
    const int h = k >> 1;
    for (int c = threadIdx.x; c < ncmp; c += blockDim.x) {
      const int blk = c / h, r = c - blk * h;
      const int i = blk * k + r, p = blk * k + (k - 1 - r);
      if (p < n) {
        const unsigned long long a = keys[i], b = keys[p];
        if (a > b) { keys[i] = b; keys[p] = a; }
      }
    }
    __syncthreads();
    for (int j = h >> 1; j >= 1; j >>= 1) {
      for (int c = threadIdx.x; c < ncmp; c += blockDim.x) {
        const int blk = c / j, r = c - blk * j;
        const int i = blk * 2 * j + r, p = i + j;
        if (p < n) {
          const unsigned long long a = keys[i], b = keys[p];
          if (a > b) { keys[i] = b; keys[p] = a; }
        }
      }
      __syncthreads();
    }
  }
}

struct BlendArgs {
  int F, P, H, W, gx, gy;
  float bg0, bg1, bg2;
  const float4* splat;
  const uint32_t* tile_start;
  unsigned long long* keys;
  uint32_t* point_list;
  long long cap;
  const float2* subpixel_offset;
  float* out_rgba;
  float* final_T;
  uint32_t* n_contrib;
  uint32_t* status;
};

__global__ void __launch_bounds__(GVF_TILE_PIX) sort_blend_kernel(const BlendArgs a) {
  __shared__ unsigned long long skeys[kSortCap];
  __shared__ float4 sA[GVF_TILE_PIX];   // px, py, conic a, conic b
  __shared__ float2 sT[GVF_TILE_PIX];   // conic c, rejection threshold on `power` (see below)
  __shared__ float4 sC[GVF_TILE_PIX];   // opacity', r, g, b  (read only by pixels the splat reaches)

  const int T = a.gx * a.gy;
  const int tile = blockIdx.x;
  const int f = tile / T, t = tile - f * T;
  const int tyi = t / a.gx, txi = t - tyi * a.gx;
  const int tid = threadIdx.x;

  long long s = a.tile_start[tile], e = a.tile_start[tile + 1];
  if (s > a.cap) s = a.cap;
  if (e > a.cap) e = a.cap;
  const int n = (int)(e - s);
  const bool in_smem = n <= kSortCap;
  unsigned long long* gk = a.keys + s;

  if (n > 0) {
    if (in_smem) {
      for (int j = tid; j < n; j += GVF_TILE_PIX) skeys[j] = gk[j];
      __syncthreads();
      bitonic_sort_any_n(skeys, n);
      for (int j = tid; j < n; j += GVF_TILE_PIX) {
        const unsigned long long k = skeys[j];
        gk[j] = k;
        a.point_list[s + j] = (uint32_t)k;
      }
    } else {
      __syncthreads();
      bitonic_sort_any_n(gk, n);
      for (int j = tid; j < n; j += GVF_TILE_PIX) a.point_list[s + j] = (uint32_t)gk[j];
      if (tid == 0) atomicMax(a.status + 2, (uint32_t)n);
      __syncthreads();
    }
  }

  const int px = txi * GVF_TILE + (tid & (GVF_TILE - 1));
  const int py = tyi * GVF_TILE + (tid >> 4);
  const bool inside = px < a.W && py < a.H;
  const size_t HW = (size_t)a.H * a.W;
  const size_t pid = (size_t)py * a.W + px;
  float pfx = (float)px, pfy = (float)py;
  if (inside && a.subpixel_offset) {
    const float2 o = a.subpixel_offset[pid];
    pfx += o.x;
    pfy += o.y;
  }
  float Tr = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
  uint32_t contributor = 0, last = 0;
  bool done = !inside;
  const float4* sp = a.splat + (size_t)f * a.P * 3;

  for (int base = 0; base < n; base += GVF_TILE_PIX) {
    if (__syncthreads_count(done) == GVF_TILE_PIX) break;
    const int j = base + tid;
    if (j < n) {
      const uint32_t id = in_smem ? (uint32_t)skeys[j] : (uint32_t)gk[j];
      const float4* r = sp + (size_t)id * 3;
      const float4 r1 = __ldg(r + 1);
      sA[tid] = __ldg(r);
      // 9 of 10 (pixel, splat) pairs of a tile end at "alpha < 1/255".  op * exp(power) < 1/255 is decided on
      // `power` alone against log(1 / (255 op)); the 0.01 margin (1 % in alpha, __expf is good to 1e-6) keeps
      // the pre-test strictly conservative, so the exact test below takes the same decisions as before
      sT[tid] = make_float2(r1.x, -__logf(255.0f * r1.y) - 0.01f);
      sC[tid] = make_float4(r1.y, r1.z, r1.w, __ldg(reinterpret_cast<const float*>(r + 2)));
    }
    __syncthreads();
    const int m = min(GVF_TILE_PIX, n - base);
    for (int k = 0; !done && k < m; ++k) {
      ++contributor;
      const float4 A = sA[k];
      const float2 ct = sT[k];
      const float dx = A.x - pfx, dy = A.y - pfy;
      const float power = -0.5f * (A.z * dx * dx + ct.x * dy * dy) - A.w * dx * dy;
      if (power > 0.0f || power < ct.y) continue;
      const float4 B = sC[k];
      const float alpha = fminf(0.99f, B.x * __expf(power));
      if (alpha < 1.0f / 255.0f) continue;
      const float test_T = Tr * (1.0f - alpha);
      if (test_T < 0.0001f) { done = true; continue; }
      const float w = alpha * Tr;
      C0 += B.y * w;
      C1 += B.z * w;
      C2 += B.w * w;
      Tr = test_T;
      last = contributor;
    }
  }
  if (inside) {
    float* o = a.out_rgba + (size_t)f * 4 * HW + pid;
    o[0] = C0 + Tr * a.bg0;
    o[HW] = C1 + Tr * a.bg1;
    o[2 * HW] = C2 + Tr * a.bg2;
    o[3 * HW] = 1.0f - Tr;
    a.final_T[(size_t)f * HW + pid] = Tr;
    a.n_contrib[(size_t)f * HW + pid] = last;
  }
}

cudaError_t launch_sort_blend(const gvf_raster_params& prm, int F, int P, const RasterWs& ws,
                              int64_t cap, const float* subpixel_offset, float* out_rgba,
                              cudaStream_t st) {
  BlendArgs a;
  a.F = F; a.P = P; a.H = prm.H; a.W = prm.W;
  a.gx = (prm.W + GVF_TILE - 1) / GVF_TILE;
  a.gy = (prm.H + GVF_TILE - 1) / GVF_TILE;
  a.bg0 = prm.bg[0]; a.bg1 = prm.bg[1]; a.bg2 = prm.bg[2];
  a.splat = ws.splat; a.tile_start = ws.tile_start; a.keys = ws.keys;
  a.point_list = ws.point_list; a.cap = cap;
  a.subpixel_offset = reinterpret_cast<const float2*>(subpixel_offset);
  a.out_rgba = out_rgba; a.final_T = ws.final_T; a.n_contrib = ws.n_contrib;
  a.status = ws.status;
  const unsigned grid = (unsigned)((size_t)F * a.gx * a.gy);
  sort_blend_kernel<<<grid, GVF_TILE_PIX, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace gvf
