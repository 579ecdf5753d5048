// raster_bin.cu -- tile binning: exclusive scan of per-(frame,tile) counts and the scatter
// of (depth | id) keys into per-tile segments.
//
// Replaces the `cub::DeviceScan::InclusiveSum` -> `duplicateWithKeys` ->
// `cub::DeviceRadixSort::SortPairs` -> `identifyTileRanges` sequence of
// diff_gaussian_rasterization (call site renderers/gaussian_render.py:198-206).  Instead
// of one global 64-bit radix sort over all tile instances (8 passes over R pairs), keys are
// scattered straight into their tile's segment (unordered) and each tile's segment is
// sorted on chip by the blend kernel (raster_blend.cu).  The resulting order is identical
// to upstream's stable sort on (tile<<32 | depth) with ids ascending on ties, because the
// per-tile key is the unique pair (depth bits, gaussian id).
#include "raster_common.h"

namespace gvf {

// One launch: every CTA scans a chunk of kScanChunk counts (exclusive, chunk-local) and
// publishes its total; the last CTA to finish scans the chunk totals.  Consumers add
// chunk_base[tile / kScanChunk].  tile_start[n] (grand total) is written by the last CTA.
__global__ void __launch_bounds__(256) scan_kernel(const uint32_t* __restrict__ cnt,
                                                   uint32_t* __restrict__ start, int n,
                                                   uint32_t* __restrict__ tmp,
                                                   uint32_t* __restrict__ status) {
  constexpr int PER = kScanChunk / 256;  // 8 items per thread
  __shared__ uint32_t warp_sums[8];
  __shared__ bool is_last;
  const int chunk = blockIdx.x, nchunks = gridDim.x;
  const int base = chunk * kScanChunk + threadIdx.x * PER;
  uint32_t v[PER], s = 0;
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    v[k] = (base + k < n) ? cnt[base + k] : 0u;
    s += v[k];
  }
  // warp inclusive scan of per-thread sums
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t inc = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[wid] = inc;
  __syncthreads();
  uint32_t wbase = 0, total = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    if (w < wid) wbase += warp_sums[w];
    total += warp_sums[w];
  }
  uint32_t run = wbase + inc - s;
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    if (base + k < n) start[base + k] = run;  // chunk-local exclusive prefix
    run += v[k];
  }
  // tmp: [0] ticket counter, [1 .. nchunks] chunk totals -> chunk bases
  if (threadIdx.x == 0) {
    tmp[1 + chunk] = total;
    __threadfence();
    const uint32_t ticket = atomicAdd(&tmp[0], 1u);
    is_last = (ticket == (uint32_t)nchunks - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // last CTA: exclusive scan of chunk totals (serial over 256-wide strips)
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int b = 0; b < nchunks; b += 256) {
    const int idx = b + threadIdx.x;
    const uint32_t x = (idx < nchunks) ? ((volatile uint32_t*)tmp)[1 + idx] : 0u;
    uint32_t in2 = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, in2, o);
      if (lane >= o) in2 += t;
    }
    if (lane == 31) warp_sums[wid] = in2;
    __syncthreads();
    uint32_t wb = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      if (w < wid) wb += warp_sums[w];
      tot += warp_sums[w];
    }
    const uint32_t c = carry;
    if (idx < nchunks) tmp[1 + idx] = c + wb + in2 - x;
    __syncthreads();
    if (threadIdx.x == 0) carry = c + tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    start[n] = carry;   // grand total (chunk base of a virtual chunk; consumers never add)
    status[0] = carry;  // num_rendered
    tmp[0] = 0;         // re-arm the ticket for the next call
  }
}

// finalize: start[i] += chunk_base[i / kScanChunk]  (so consumers read one array)
__global__ void __launch_bounds__(256) scan_fix_kernel(uint32_t* __restrict__ start, int n,
                                                       const uint32_t* __restrict__ tmp) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < n) start[i] += tmp[1 + i / kScanChunk];
}

cudaError_t launch_scan(int n, const RasterWs& ws, cudaStream_t st) {
  const int chunks = (n + kScanChunk - 1) / kScanChunk;
  scan_kernel<<<chunks, 256, 0, st>>>(ws.tile_count, ws.tile_start, n, ws.scan_tmp, ws.status);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (chunks > 1) {
    scan_fix_kernel<<<(n + 255) / 256, 256, 0, st>>>(ws.tile_start, n, ws.scan_tmp);
    e = cudaGetLastError();
  }
  return e;
}

// Scatter: one thread per (frame, Gaussian); slots inside a tile's segment are claimed by
// counting tile_count back down to zero (which also leaves tile_count cleared).  Each claim is an
// atomic round trip (~700 clocks), so a splat covering hundreds of tiles used to keep one lane -- and its
// warp -- busy for hundreds of round trips (447 us at the benchmark scene, most of it that tail): rectangles
// of more than kScatterSmall tiles are now spread over the 32 lanes of the warp, one after the other.
constexpr int kScatterSmall = 8;

__device__ __forceinline__ void scatter_one(size_t t, unsigned long long key, uint32_t* __restrict__ tile_count,
                                            const uint32_t* __restrict__ tile_start,
                                            unsigned long long* __restrict__ keys, long long cap,
                                            uint32_t* __restrict__ status) {
  const uint32_t slot = atomicSub(tile_count + t, 1u) - 1u;
  const long long pos = (long long)tile_start[t] + slot;
  if (pos < cap) keys[pos] = key;
  else status[1] = 1u;
}

__global__ void __launch_bounds__(256) scatter_kernel(int F, int P, int gx, int gy,
                                                      const float4* __restrict__ splat,
                                                      const ushort4* __restrict__ rect,
                                                      uint32_t* __restrict__ tile_count,
                                                      const uint32_t* __restrict__ tile_start,
                                                      unsigned long long* __restrict__ keys,
                                                      long long cap, uint32_t* __restrict__ status) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  int rx = 0, ry = 0, w = 0, n = 0;
  unsigned long long key = 0;
  size_t tb = 0;
  if (gid < (long long)F * P) {
    const ushort4 r = rect[gid];
    if (r.x < r.z && r.y < r.w) {
      const int f = (int)(gid / P);
      const uint32_t i = (uint32_t)(gid - (long long)f * P);
      const float depth = splat[gid * 3 + 2].y;
      key = ((unsigned long long)__float_as_uint(depth) << 32) | i;
      tb = (size_t)f * gx * gy;
      rx = r.x; ry = r.y; w = r.z - r.x; n = w * (r.w - r.y);
    }
  }
  if (n > 0 && n <= kScatterSmall) {
#pragma unroll 4
    for (int t = 0; t < n; ++t)
      scatter_one(tb + (size_t)(ry + t / w) * gx + rx + t % w, key, tile_count, tile_start, keys, cap, status);
  }
  unsigned big = __ballot_sync(0xffffffffu, n > kScatterSmall);
  while (big) {
    const int j = __ffs(big) - 1;
    big &= big - 1;
    const int jrx = __shfl_sync(0xffffffffu, rx, j), jry = __shfl_sync(0xffffffffu, ry, j);
    const int jw = __shfl_sync(0xffffffffu, w, j), jn = __shfl_sync(0xffffffffu, n, j);
    const unsigned long long jkey = __shfl_sync(0xffffffffu, key, j);
    const size_t jtb = __shfl_sync(0xffffffffu, (unsigned long long)tb, j);
    for (int t = lane; t < jn; t += 32)
      scatter_one(jtb + (size_t)(jry + t / jw) * gx + jrx + t % jw, jkey, tile_count, tile_start, keys, cap, status);
  }
}

cudaError_t launch_scatter(const gvf_raster_params& prm, int F, int P, const RasterWs& ws,
                           int64_t cap, cudaStream_t st) {
  const int gx = (prm.W + GVF_TILE - 1) / GVF_TILE, gy = (prm.H + GVF_TILE - 1) / GVF_TILE;
  const long long n = (long long)F * P;
  scatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
      F, P, gx, gy, ws.splat, ws.rect, ws.tile_count, ws.tile_start, ws.keys, (long long)cap,
      ws.status);
  return cudaGetLastError();
}

}  // namespace gvf
