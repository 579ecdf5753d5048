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
#include <stdlib.h>

#include "raster_common.h"

namespace gvf {

constexpr int kSortCap = 2048;
#ifndef GVF_BLEND_MIN_CTAS
#define GVF_BLEND_MIN_CTAS 6
#endif

template <bool PADDED, typename KeyPtr>
__device__ __forceinline__ void bitonic_sort_any_n(KeyPtr keys, int n) {
  // all-ascending bitonic network ("flip" then "disperse"), valid for arbitrary n.  Block sizes are powers of
  // two: comparator -> element indices by shifts and masks (lk = log2 k), no integer division.  PADDED: the
  // caller filled keys[n .. 2^ceil(log2 n)) with +inf, so no comparator needs a bounds test.
  // MEASURED and rejected: the same network with four / eight keys per thread in registers (in-thread
  // comparators, warp shuffles up to 16 threads, shared memory only for the six longest distances of 1024
  // keys): identical order, 24 frames 0.558 ms against 0.520 ms with this version -- 64-bit shuffles cost what
  // the shared-memory accesses did, and lists of 65..128 keys ran on a single warp.
  int lnp2 = 0;
  while ((1 << lnp2) < n) ++lnp2;
  const int ncmp = (1 << lnp2) >> 1;
  for (int lk = 1; lk <= lnp2; ++lk) {
    const int lh = lk - 1, hmask = (1 << lh) - 1, kfull = (1 << lk) - 1;
    for (int c = threadIdx.x; c < ncmp; c += blockDim.x) {
      const int blk = c >> lh, r = c & hmask;
      const int i = (blk << lk) + r, p = (blk << lk) + (kfull - r);
      if (PADDED || p < n) {
        const unsigned long long a = keys[i], b = keys[p];
        if (a > b) { keys[i] = b; keys[p] = a; }
      }
    }
    __syncthreads();
    for (int lj = lh - 1; lj >= 0; --lj) {
      const int jmask = (1 << lj) - 1;
      for (int c = threadIdx.x; c < ncmp; c += blockDim.x) {
        const int i = ((c >> lj) << (lj + 1)) + (c & jmask), p = i + (1 << lj);
        if (PADDED || p < n) {
          const unsigned long long a = keys[i], b = keys[p];
          if (a > b) { keys[i] = b; keys[p] = a; }
        }
      }
      __syncthreads();
    }
  }
}

__device__ __forceinline__ float ex2_approx(float x) {   // MUFU.EX2; flushes denormal results to zero
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Generation 2 of the per-tile depth sort (default; GVF_RASTER_SORT=bitonic selects the network): one-pass
// distribution sort.
// A tile list holds a few hundred keys whose depth bits are spread over a narrow range, so a monotone map
// of the depth bits onto kBuckets buckets leaves 0..3 keys per bucket.  thread t keeps keys t, t + 256, ..
// in registers: (A) min / max of the depth bits, (B) bucket + arrival rank by a shared-memory atomic,
// (C) exclusive scan of the bucket counts, (D) keys stored bucket-major, (E) every key counts the smaller
// keys of its own bucket = its final place.  bucket() is monotone in the key and keys are unique, so the
// result is the ascending order the bitonic network produces (bit-identical point lists).  ~8 barriers
// instead of 55 for 1024 keys.  Returns false (list left permuted, caller runs the network) when a bucket
// holds more than kBucketMax keys (near-constant depth): correctness never depends on the distribution.
constexpr int kBuckets = 1024;
constexpr int kBucketMax = 24;
constexpr int kKeysPerThread = kSortCap / GVF_TILE_PIX;
constexpr int kBucketMin = 128;          // shorter lists: the network's few steps are cheaper than the fixed cost

__device__ __forceinline__ bool bucket_sort_tile(const unsigned long long* __restrict__ gk, int n,
                                                 unsigned long long* skeys, uint32_t* cnt, uint32_t* red) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned long long key[kKeysPerThread];
  uint32_t lo = 0xffffffffu, hi = 0u;
#pragma unroll
  for (int i = 0; i < kKeysPerThread; ++i) {
    const int j = tid + i * GVF_TILE_PIX;
    key[i] = j < n ? gk[j] : ~0ull;
    if (j < n) {
      const uint32_t d = (uint32_t)(key[i] >> 32);
      lo = min(lo, d);
      hi = max(hi, d);
    }
  }
#pragma unroll
  for (int c = tid; c < kBuckets; c += GVF_TILE_PIX) cnt[c] = 0;
  lo = __reduce_min_sync(0xffffffffu, lo);
  hi = __reduce_max_sync(0xffffffffu, hi);
  if (lane == 0) { red[warp] = lo; red[8 + warp] = hi; }
  __syncthreads();
  lo = red[lane & 7];
  hi = red[8 + (lane & 7)];
  lo = __reduce_min_sync(0xffffffffu, lo);
  hi = __reduce_max_sync(0xffffffffu, hi);
  const float scale = (float)kBuckets / (__uint2float_rz(hi - lo) + 1.0f);
  uint32_t bk[kKeysPerThread];          // bucket << 16 | arrival rank
#pragma unroll
  for (int i = 0; i < kKeysPerThread; ++i) {
    const int j = tid + i * GVF_TILE_PIX;
    if (j < n) {
      const uint32_t d = (uint32_t)(key[i] >> 32) - lo;
      const int b = min(kBuckets - 1, (int)(__uint2float_rz(d) * scale));
      bk[i] = ((uint32_t)b << 16) | atomicAdd(&cnt[b], 1u);
    }
  }
  __syncthreads();
  // exclusive scan of cnt: thread t owns buckets 4t .. 4t + 3
  uint32_t c4[4], tot = 0, big = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    c4[q] = cnt[4 * tid + q];
    big |= c4[q] > (uint32_t)kBucketMax;
    tot += c4[q];
  }
  uint32_t inc = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) red[16 + warp] = inc;
  const int any_big = __syncthreads_or((int)big);
  uint32_t base = inc - tot;
#pragma unroll
  for (int w = 0; w < 7; ++w)
    if (w < warp) base += red[16 + w];
  // cnt becomes (start << 8 | count): start < 2048 and count <= kBucketMax fit comfortably
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    cnt[4 * tid + q] = (base << 8) | min(c4[q], 255u);
    base += c4[q];
  }
  __syncthreads();
  if (any_big) {                       // degenerate depth distribution: hand the keys back unsorted
#pragma unroll
    for (int i = 0; i < kKeysPerThread; ++i) {
      const int j = tid + i * GVF_TILE_PIX;
      if (j < n) skeys[j] = key[i];
    }
    __syncthreads();
    return false;
  }
#pragma unroll
  for (int i = 0; i < kKeysPerThread; ++i) {
    const int j = tid + i * GVF_TILE_PIX;
    if (j < n) skeys[(cnt[bk[i] >> 16] >> 8) + (bk[i] & 0xffffu)] = key[i];
  }
  __syncthreads();
  uint32_t fin[kKeysPerThread];
#pragma unroll
  for (int i = 0; i < kKeysPerThread; ++i) {
    const int j = tid + i * GVF_TILE_PIX;
    if (j < n) {
      const uint32_t sc = cnt[bk[i] >> 16];
      const uint32_t st = sc >> 8, c = sc & 0xffu;
      uint32_t r = 0;
      for (uint32_t q = 0; q < c; ++q) r += skeys[st + q] < key[i] ? 1u : 0u;
      fin[i] = st + r;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kKeysPerThread; ++i) {
    const int j = tid + i * GVF_TILE_PIX;
    if (j < n) skeys[fin[i]] = key[i];
  }
  __syncthreads();
  return true;
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
  int sort_mode;              // 0 bitonic network, 1 bucket sort for lists of more than kBucketMin keys
};

__global__ void __launch_bounds__(GVF_TILE_PIX, GVF_BLEND_MIN_CTAS) sort_blend_kernel(const BlendArgs a) {
  __shared__ unsigned long long skeys[kSortCap];
  __shared__ float4 sA[GVF_TILE_PIX];   // px, py, conic a, conic b
  __shared__ float2 sT[GVF_TILE_PIX];   // conic c, rejection threshold on `power` (see below)
  __shared__ float4 sC[GVF_TILE_PIX];   // opacity', r, g, b  (read only by pixels the splat reaches)
  __shared__ uint32_t sRed[24];
  uint32_t* sCnt = reinterpret_cast<uint32_t*>(sA);   // bucket sort scratch (kBuckets words): dead before sA is staged
  static_assert(sizeof(float4) * GVF_TILE_PIX >= sizeof(uint32_t) * kBuckets, "sCnt aliases sA");
  __shared__ uint32_t sM[GVF_TILE_PIX]; // bit w: the splat's alpha >= 1/255 ellipse may reach warp w's 8 x 4 block

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
      int np2 = 1;
      while (np2 < n) np2 <<= 1;
      bool sorted = false;
      if (a.sort_mode == 1 && n > kBucketMin) {
        sorted = bucket_sort_tile(gk, n, skeys, sCnt, sRed);
        if (!sorted) {
          for (int j = n + tid; j < np2; j += GVF_TILE_PIX) skeys[j] = ~0ull;
          __syncthreads();
        }
      } else {
        for (int j = tid; j < np2; j += GVF_TILE_PIX) skeys[j] = j < n ? gk[j] : ~0ull;
        __syncthreads();
      }
      if (!sorted) bitonic_sort_any_n<true>(skeys, n);
      for (int j = tid; j < n; j += GVF_TILE_PIX) {
        const unsigned long long k = skeys[j];
        gk[j] = k;
        a.point_list[s + j] = (uint32_t)k;
      }
    } else {
      __syncthreads();
      bitonic_sort_any_n<false>(gk, n);
      for (int j = tid; j < n; j += GVF_TILE_PIX) a.point_list[s + j] = (uint32_t)gk[j];
      if (tid == 0) atomicMax(a.status + 2, (uint32_t)n);
      __syncthreads();
    }
  }

  // a warp covers an 8 x 4 pixel block (not a 16 x 2 strip): the more compact footprint lets whole warps
  // leave the loop body at the rejection test more often
  const int wq = tid >> 5, ln = tid & 31;
  const int px = txi * GVF_TILE + (ln & 7) + 8 * (wq & 1);
  const int py = tyi * GVF_TILE + (ln >> 3) + 4 * (wq >> 1);
  const bool inside = px < a.W && py < a.H;
  const size_t HW = (size_t)a.H * a.W;
  const size_t pid = (size_t)py * a.W + px;
  float pfx = (float)px, pfy = (float)py;
  if (inside && a.subpixel_offset) {
    const float2 o = a.subpixel_offset[pid];
    pfx += o.x;
    pfy += o.y;
  }
  float Tr = 1.0f, Tfin = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
  uint32_t last = 0;
  bool done = !inside;
  const float4* sp = a.splat + (size_t)f * a.P * 3;
  const float tile_x0 = (float)(txi * GVF_TILE), tile_y0 = (float)(tyi * GVF_TILE);
  const float slack = a.subpixel_offset ? 1.01f : 0.01f;      // pixel centres move by the sub-pixel offset

  // MEASURED and rejected: requesting the next batch's 48 B splat records before the blend loop of the current
  // one (registers as the second stage buffer): 0.454 vs 0.456 ms per 24 frames -- with six CTAs per SM the
  // gather latency is already covered by other CTAs.
  for (int base = 0; base < n; base += GVF_TILE_PIX) {
    if (__syncthreads_count(done) == GVF_TILE_PIX) break;
    const int j = base + tid;
    if (j < n) {
      const uint32_t id = in_smem ? (uint32_t)skeys[j] : (uint32_t)gk[j];
      const float4* r = sp + (size_t)id * 3;
      const float4 r0 = __ldg(r), r1 = __ldg(r + 1);
      // The conic is staged pre-multiplied by -0.5 log2(e) (a, c) and -log2(e) (b): the loop evaluates
      // log2 of the Gaussian with five FMUL / FFMA and feeds MUFU.EX2 directly.
      // 9 of 10 (pixel, splat) pairs of a tile end at "alpha < 1/255".  op * 2^p < 1/255 is decided on p alone
      // against -log2(255 op); the 0.0145 margin (1 % in alpha, ex2.approx is good to 2^-22) keeps the pre-test
      // strictly conservative, so the exact test below takes the same decisions as without it
      constexpr float kLog2e = 1.4426950408889634f;
      const float qa = 0.5f * kLog2e * r0.z, qb = kLog2e * r0.w, qc = 0.5f * kLog2e * r1.x;
      const float tau = __log2f(255.0f * r1.y) + 0.0145f;      // the pair survives iff qa dx^2 + qb dx dy + qc dy^2 <= tau
      sA[tid] = make_float4(r0.x, r0.y, -qa, -qb);
      sT[tid] = make_float2(-qc, -tau);
      sC[tid] = make_float4(r1.y, r1.z, r1.w, __ldg(reinterpret_cast<const float*>(r + 2)));
      // Sub-tile culling.  ncu: 86 % of the (warp, splat) pairs of the list were rejected by all 32 lanes, each
      // at the price of the full rejection test.  The thread that stages a splat intersects the axis-aligned
      // bounding box of its {alpha >= 1/255} ellipse (half extents sqrt(tau qc / det), sqrt(tau qa / det),
      // inflated against rounding) with the eight 8 x 4 pixel blocks of the tile, one per warp; a warp then
      // walks only the splats whose bit is set.  Skipped pairs are pairs the exact test rejects: same image.
      uint32_t mask = 0;
      if (tau > 0.0f) {
        const float det = qa * qc - 0.25f * qb * qb;
        mask = 0xffu;
        if (det > 0.0f) {
          const float k = tau / det;
          const float hx = sqrtf(k * qc) * 1.001f + slack, hy = sqrtf(k * qa) * 1.001f + slack;
          const float cx = r0.x - tile_x0, cy = r0.y - tile_y0;
          const float xl = cx - hx, xh = cx + hx, yl = cy - hy, yh = cy + hy;
          const uint32_t cols = (xl <= 7.0f && xh >= 0.0f ? 1u : 0u) | (xl <= 15.0f && xh >= 8.0f ? 2u : 0u);
          mask = 0;
#pragma unroll
          for (int rr = 0; rr < 4; ++rr)
            if (yl <= (float)(4 * rr + 3) && yh >= (float)(4 * rr)) mask |= cols << (2 * rr);
        }
      }
      sM[tid] = mask;
    }
    __syncthreads();
    const int m = min(GVF_TILE_PIX, n - base);
    for (int g = 0; g < m; g += 32) {
      if (__all_sync(0xffffffffu, done)) break;
      const uint32_t mk = (g + ln < m) ? sM[g + ln] : 0u;
      uint32_t bits = __ballot_sync(0xffffffffu, (mk >> wq) & 1u);
      while (bits) {
        const int k = g + __ffs(bits) - 1;
        bits &= bits - 1;
        const float4 A = sA[k];
        const float2 ct = sT[k];
        const float dx = A.x - pfx, dy = A.y - pfy;
        const float t = fmaf(A.z, dx, A.w * dy);
        const float p2 = fmaf(dx, t, (ct.x * dy) * dy);       // log2 of exp(power)
        if (p2 > 0.0f || p2 < ct.y) continue;
        const float4 B = sC[k];
        const float alpha = fminf(0.99f, B.x * ex2_approx(p2));
        if (alpha < 1.0f / 255.0f) continue;
        const float test_T = Tr * (1.0f - alpha);
        if (test_T < 0.0001f) {
          // saturated: this pixel is finished.  Its lane stays in the warp's loop (the ballots need it) with a
          // live transmittance of zero, which sends every later candidate to this branch again.
          if (!done) { Tfin = Tr; done = true; Tr = 0.0f; }
          continue;
        }
        const float w = alpha * Tr;
        C0 += B.y * w;
        C1 += B.z * w;
        C2 += B.w * w;
        Tr = test_T;
        last = (uint32_t)(base + k + 1);                      // entries of the tile list seen up to this one
      }
    }
  }
  if (done) Tr = Tfin;
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

int g_raster_sort_mode = -1;

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
  // GVF_RASTER_SORT=bitonic selects generation 1 of the per-tile sort (identical point lists);
  // gvf_raster_set_sort overrides the environment
  static int env_mode = -1;
  if (env_mode < 0) {
    const char* e = getenv("GVF_RASTER_SORT");
    env_mode = (e && e[0] == 'b' && e[1] == 'i') ? 0 : 1;      // default: bucket sort
  }
  a.sort_mode = g_raster_sort_mode >= 0 ? g_raster_sort_mode : env_mode;
  // 27.7 KB of shared memory and 40 registers per thread: six CTAs per SM once the carve-out leaves room
  // (ncu: four CTAs, limited by both, left the barrier stalls of the sort and staging phases exposed)
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(sort_blend_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    configured = true;
  }
  const unsigned grid = (unsigned)((size_t)F * a.gx * a.gy);
  sort_blend_kernel<<<grid, GVF_TILE_PIX, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace gvf

// tuning hook: -1 environment (GVF_RASTER_SORT), 0 bitonic network, 1 bucket sort
extern "C" GVF_API void gvf_raster_set_sort(int mode) { gvf::g_raster_sort_mode = mode; }
