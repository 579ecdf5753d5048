// losses.cu -- render-loss and point-cloud kernels of the training step.
//
//   * gvf_ssim_l1_fwd / gvf_ssim_l1_bwd: the pixel losses of reference train_vae.py:328-330
//       L1 = |pred - gt|.mean()            (nn.L1Loss)
//       SSIM(pred, gt)                     (utils/loss_util.py:33-63: 11x11 Gaussian window, sigma 1.5,
//                                           zero padding, C1 = 0.01^2, C2 = 0.03^2)
//     The reference runs five depthwise cuDNN convolutions plus ~20 elementwise launches over
//     (B*cams, 3, 512, 512) and autograd replays them backwards.  Here one kernel per direction:
//     a 32x32 pixel tile with its 5-pixel halo is staged in shared memory once, the window is applied
//     separably (rows, then columns) to the five moments at the same time, the SSIM map is reduced to
//     per-CTA partial sums (fixed-order second pass: deterministic), and the three per-pixel partial
//     derivatives the backward pass needs are written out.  Backward convolves those three maps with the
//     same window and adds the L1 sign term.  HBM class: forward reads 8 B and writes 12 B per pixel-channel,
//     backward reads 20 B and writes 4 B.
//   * gvf_knn: exact K-nearest-neighbour search (K <= 16), replaces pytorch3d.ops.knn_points as called
//     at reference train_vae.py:525-530 (and model/autoencoder.py's encode path).  One thread per query,
//     reference cloud staged through shared memory, sorted insertion in registers; squared distances
//     ((dx*dx + dy*dy) + dz*dz, no FMA contraction: bit-exact against the numpy oracle), ascending, ties ->
//     lowest index; rows beyond lengths1 / columns beyond lengths2 are zero like pytorch3d's.
//   * gvf_knn_interp_deltas: the RBF-weighted neighbour-motion estimate of
//     compute_interpolation_loss_delta_interp (train_vae.py:532-563) in one pass.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gvf_b200.h"

namespace gvf {

constexpr int kWin = 11, kHalo = 5, kTile = 32, kReg = kTile + 2 * kHalo;   // 42
constexpr int kRegLd = 44;       // staged rows: 16 B aligned 8-column groups, conflict-free LDS.128
constexpr int kHzLd = kTile + 1; // row-filtered rows: conflict-free column-group stores
constexpr int kGrp = 8;          // output columns per thread in the row pass (18 staged values -> 8 outputs)
constexpr int kRows = 4;         // output rows per thread in the column pass (14 row-filtered values -> 4 outputs)

struct SsimWindow { float g[kWin]; };

static SsimWindow make_window() {                 // utils/loss_util.py:24-26 (fp32 tensor of python doubles)
  SsimWindow w;
  float s = 0.f;
  for (int i = 0; i < kWin; ++i) {
    w.g[i] = (float)exp(-(double)((i - kWin / 2) * (i - kWin / 2)) / (2.0 * 1.5 * 1.5));
    s += w.g[i];
  }
  for (int i = 0; i < kWin; ++i) w.g[i] /= s;
  return w;
}

__device__ __forceinline__ float block_sum_256(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < 8) t = red[threadIdx.x];
  if (w == 0) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;                                       // valid in thread 0
}

// grid (ceil(W/32), ceil(H/32), planes), 256 threads.  partials [planes][gy*gx][2] = (sum ssim_map, sum |a-b|)
__global__ void __launch_bounds__(256) ssim_l1_fwd_kernel(const float* __restrict__ img1, const float* __restrict__ img2,
                                                          int H, int W, const SsimWindow win,
                                                          float* __restrict__ partials, float* __restrict__ dmaps,
                                                          long long plane_count) {
  __shared__ __align__(16) float sa[kReg][kRegLd], sb[kReg][kRegLd];
  __shared__ float hz[5][kReg][kHzLd];
  __shared__ float red[8];
  const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile;
  const size_t HW = (size_t)H * W;
  const float* a = img1 + (size_t)blockIdx.z * HW;
  const float* b = img2 + (size_t)blockIdx.z * HW;
  for (int i = threadIdx.x; i < kReg * kRegLd; i += 256) {
    const int r = i / kRegLd, c = i - r * kRegLd;
    const int y = y0 + r - kHalo, x = x0 + c - kHalo;
    const bool in = c < kReg && y >= 0 && y < H && x >= 0 && x < W;
    sa[r][c] = in ? __ldg(a + (size_t)y * W + x) : 0.f;
    sb[r][c] = in ? __ldg(b + (size_t)y * W + x) : 0.f;
  }
  __syncthreads();
  // row pass: one thread = one staged row x 8 output columns; the 18 staged values it needs are read once
  // (4 x LDS.128 + LDS.64 per image) and the three products formed once per value, not once per tap
  if (threadIdx.x < kReg * (kTile / kGrp)) {
    const int r = threadIdx.x / (kTile / kGrp), c0 = (threadIdx.x % (kTile / kGrp)) * kGrp;
    float u[kGrp + kWin - 1], v[kGrp + kWin - 1];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 t = *reinterpret_cast<const float4*>(&sa[r][c0 + 4 * q]);
      const float4 w = *reinterpret_cast<const float4*>(&sb[r][c0 + 4 * q]);
      u[4 * q] = t.x; u[4 * q + 1] = t.y; u[4 * q + 2] = t.z; u[4 * q + 3] = t.w;
      v[4 * q] = w.x; v[4 * q + 1] = w.y; v[4 * q + 2] = w.z; v[4 * q + 3] = w.w;
    }
    {
      const float2 t = *reinterpret_cast<const float2*>(&sa[r][c0 + 16]);
      const float2 w = *reinterpret_cast<const float2*>(&sb[r][c0 + 16]);
      u[16] = t.x; u[17] = t.y; v[16] = w.x; v[17] = w.y;
    }
    float m1[kGrp], m2[kGrp], s11[kGrp], s22[kGrp], s12[kGrp];
#pragma unroll
    for (int j = 0; j < kGrp; ++j) m1[j] = m2[j] = s11[j] = s22[j] = s12[j] = 0.f;
#pragma unroll
    for (int i = 0; i < kGrp + kWin - 1; ++i) {
      const float uu = u[i] * u[i], vv = v[i] * v[i], uv = u[i] * v[i];
#pragma unroll
      for (int j = 0; j < kGrp; ++j) {
        const int k = i - j;                       // tap index of value i for output j
        if (k >= 0 && k < kWin) {
          const float g = win.g[k];
          m1[j] += g * u[i]; m2[j] += g * v[i]; s11[j] += g * uu; s22[j] += g * vv; s12[j] += g * uv;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kGrp; ++j) {
      hz[0][r][c0 + j] = m1[j]; hz[1][r][c0 + j] = m2[j]; hz[2][r][c0 + j] = s11[j];
      hz[3][r][c0 + j] = s22[j]; hz[4][r][c0 + j] = s12[j];
    }
  }
  __syncthreads();
  // column pass: one thread = one column x 4 output rows (14 row-filtered values per moment, read once)
  const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
  float acc_s = 0.f, acc_l = 0.f;
  const int tx = threadIdx.x & 31, ty0 = (threadIdx.x >> 5) * kRows;
  float o[5][kRows];
#pragma unroll
  for (int m = 0; m < 5; ++m) {
#pragma unroll
    for (int j = 0; j < kRows; ++j) o[m][j] = 0.f;
#pragma unroll
    for (int i = 0; i < kRows + kWin - 1; ++i) {
      const float h = hz[m][ty0 + i][tx];
#pragma unroll
      for (int j = 0; j < kRows; ++j) {
        const int k = i - j;
        if (k >= 0 && k < kWin) o[m][j] += win.g[k] * h;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kRows; ++j) {
    const int ty = ty0 + j;
    const int y = y0 + ty, x = x0 + tx;
    const float m1 = o[0][j], m2 = o[1][j], s11 = o[2][j], s22 = o[3][j], s12 = o[4][j];
    if (y < H && x < W) {
      const float mu1_sq = m1 * m1, mu2_sq = m2 * m2, mu12 = m1 * m2;
      const float sig1 = s11 - mu1_sq, sig2 = s22 - mu2_sq, sig12 = s12 - mu12;
      const float A1 = 2.f * mu12 + C1, A2 = 2.f * sig12 + C2;
      const float B1 = mu1_sq + mu2_sq + C1, B2 = sig1 + sig2 + C2;
      const float inv = 1.f / (B1 * B2);
      const float s = A1 * A2 * inv;
      acc_s += s;
      acc_l += fabsf(sa[ty + kHalo][tx + kHalo] - sb[ty + kHalo][tx + kHalo]);
      if (dmaps) {
        // d s / d sigma12, d s / d sigma1^2, and the total derivative with respect to mu1 (sigma1^2 and
        // sigma12 contain -mu1^2 and -mu1 mu2)
        const float d12 = 2.f * A1 * inv;
        const float d11 = -s / B2;
        const float dmu = 2.f * m2 * A2 * inv - 2.f * m1 * s / B1 - 2.f * m1 * d11 - m2 * d12;
        const size_t oo = (size_t)blockIdx.z * HW + (size_t)y * W + x;
        dmaps[oo] = dmu;
        dmaps[oo + plane_count * HW] = d11;
        dmaps[oo + 2 * plane_count * HW] = d12;
      }
    }
  }
  const float ts = block_sum_256(acc_s, red);
  const float tl = block_sum_256(acc_l, red);
  if (threadIdx.x == 0) {
    const size_t cta = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    partials[2 * cta] = ts;
    partials[2 * cta + 1] = tl;
  }
}

// one CTA per plane: fixed-order sum of that plane's partials in double -> sums [planes][2]
__global__ void __launch_bounds__(256) ssim_l1_finalize_kernel(const float* __restrict__ partials, int per_plane,
                                                               float* __restrict__ sums) {
  __shared__ double rs[256], rl[256];
  const float* p = partials + (size_t)blockIdx.x * per_plane * 2;
  double s = 0.0, l = 0.0;
  for (int i = threadIdx.x; i < per_plane; i += 256) { s += p[2 * i]; l += p[2 * i + 1]; }
  rs[threadIdx.x] = s; rl[threadIdx.x] = l;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { rs[threadIdx.x] += rs[threadIdx.x + o]; rl[threadIdx.x] += rl[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { sums[2 * blockIdx.x] = (float)rs[0]; sums[2 * blockIdx.x + 1] = (float)rl[0]; }
}

// grad_img1 = coef_l1[plane] * sign(a - b) + coef_ssim[plane] * (w * dmu + 2 a (w * d11) + b (w * d12))
__global__ void __launch_bounds__(256) ssim_l1_bwd_kernel(const float* __restrict__ img1, const float* __restrict__ img2,
                                                          const float* __restrict__ dmaps, int H, int W,
                                                          const SsimWindow win, const float* __restrict__ coef_ssim,
                                                          const float* __restrict__ coef_l1,
                                                          float* __restrict__ grad, long long plane_count) {
  __shared__ __align__(16) float sm[3][kReg][kRegLd];
  __shared__ float hz[3][kReg][kHzLd];
  const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile;
  const size_t HW = (size_t)H * W;
  const size_t pl = (size_t)blockIdx.z * HW;
  for (int i = threadIdx.x; i < kReg * kRegLd; i += 256) {
    const int r = i / kRegLd, c = i - r * kRegLd;
    const int y = y0 + r - kHalo, x = x0 + c - kHalo;
    const bool in = c < kReg && y >= 0 && y < H && x >= 0 && x < W;
    const size_t o = pl + (size_t)y * W + x;
#pragma unroll
    for (int m = 0; m < 3; ++m) sm[m][r][c] = in ? __ldg(dmaps + o + (size_t)m * plane_count * HW) : 0.f;
  }
  __syncthreads();
  if (threadIdx.x < kReg * (kTile / kGrp)) {
    const int r = threadIdx.x / (kTile / kGrp), c0 = (threadIdx.x % (kTile / kGrp)) * kGrp;
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      float u[kGrp + kWin - 1];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 t = *reinterpret_cast<const float4*>(&sm[m][r][c0 + 4 * q]);
        u[4 * q] = t.x; u[4 * q + 1] = t.y; u[4 * q + 2] = t.z; u[4 * q + 3] = t.w;
      }
      const float2 t2 = *reinterpret_cast<const float2*>(&sm[m][r][c0 + 16]);
      u[16] = t2.x; u[17] = t2.y;
      float acc[kGrp];
#pragma unroll
      for (int j = 0; j < kGrp; ++j) acc[j] = 0.f;
#pragma unroll
      for (int i = 0; i < kGrp + kWin - 1; ++i) {
#pragma unroll
        for (int j = 0; j < kGrp; ++j) {
          const int k = i - j;
          if (k >= 0 && k < kWin) acc[j] += win.g[k] * u[i];
        }
      }
#pragma unroll
      for (int j = 0; j < kGrp; ++j) hz[m][r][c0 + j] = acc[j];
    }
  }
  __syncthreads();
  const float cs = coef_ssim[blockIdx.z], cl = coef_l1[blockIdx.z];
  const int tx = threadIdx.x & 31, ty0 = (threadIdx.x >> 5) * kRows;
  float o[3][kRows];
#pragma unroll
  for (int m = 0; m < 3; ++m) {
#pragma unroll
    for (int j = 0; j < kRows; ++j) o[m][j] = 0.f;
#pragma unroll
    for (int i = 0; i < kRows + kWin - 1; ++i) {
      const float h = hz[m][ty0 + i][tx];
#pragma unroll
      for (int j = 0; j < kRows; ++j) {
        const int k = i - j;
        if (k >= 0 && k < kWin) o[m][j] += win.g[k] * h;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kRows; ++j) {
    const int y = y0 + ty0 + j, x = x0 + tx;
    if (y >= H || x >= W) continue;
    const size_t oo = pl + (size_t)y * W + x;
    const float u = __ldg(img1 + oo), v = __ldg(img2 + oo);
    const float d = u - v;
    const float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    grad[oo] = cl * sg + cs * (o[0][j] + 2.f * u * o[1][j] + v * o[2][j]);
  }
}

// ---------------------------------------------------------------------------------------------------------
constexpr int kKnnMax = 16, kKnnChunk = 1024;

template <int K>
__global__ void __launch_bounds__(256) knn_kernel(const float* __restrict__ q, const float* __restrict__ r, int P1, int P2,
                                                  const long long* __restrict__ len1, const long long* __restrict__ len2,
                                                  int k_used, float* __restrict__ out_d, long long* __restrict__ out_i) {
  __shared__ float sx[kKnnChunk], sy[kKnnChunk], sz[kKnnChunk];
  const int b = blockIdx.y;
  const int i = blockIdx.x * 256 + threadIdx.x;
  const int n1 = len1 ? (int)len1[b] : P1, n2 = len2 ? (int)len2[b] : P2;
  const bool live = i < n1 && i < P1;
  float qx = 0.f, qy = 0.f, qz = 0.f;
  if (live) {
    const float* p = q + ((size_t)b * P1 + i) * 3;
    qx = p[0]; qy = p[1]; qz = p[2];
  }
  float bd[K];
  int bi[K];
#pragma unroll
  for (int k = 0; k < K; ++k) { bd[k] = 3.4e38f; bi[k] = -1; }
  const float* rb = r + (size_t)b * P2 * 3;
  for (int base = 0; base < n2; base += kKnnChunk) {
    const int m = min(kKnnChunk, n2 - base);
    __syncthreads();
    for (int j = threadIdx.x; j < m; j += 256) {
      const float* p = rb + (size_t)(base + j) * 3;
      sx[j] = p[0]; sy[j] = p[1]; sz[j] = p[2];
    }
    __syncthreads();
    if (!live) continue;
    for (int j = 0; j < m; ++j) {
      const float dx = __fsub_rn(qx, sx[j]), dy = __fsub_rn(qy, sy[j]), dz = __fsub_rn(qz, sz[j]);
      const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      if (d < bd[K - 1]) {                         // strict: among equal distances the earlier index stays in front
        bd[K - 1] = d; bi[K - 1] = base + j;
#pragma unroll
        for (int k = K - 1; k > 0; --k) {
          if (bd[k] < bd[k - 1]) {
            const float td = bd[k]; bd[k] = bd[k - 1]; bd[k - 1] = td;
            const int ti = bi[k]; bi[k] = bi[k - 1]; bi[k - 1] = ti;
          }
        }
      }
    }
  }
  if (i < P1) {
    float* od = out_d + ((size_t)b * P1 + i) * k_used;
    long long* oi = out_i + ((size_t)b * P1 + i) * k_used;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (k < k_used) {
        const bool ok = live && bi[k] >= 0;
        od[k] = ok ? bd[k] : 0.f;
        oi[k] = ok ? bi[k] : 0;
      }
    }
  }
}

// est[b, t, i, :] = sum_k w_k (moving[b, t, idx_k] - static[b, idx_k]); weights as train_vae.py:532-548
__global__ void __launch_bounds__(256) knn_interp_kernel(const float* __restrict__ dists, const long long* __restrict__ idx,
                                                         const float* __restrict__ stat, const float* __restrict__ mov,
                                                         const long long* __restrict__ len1, int P1, int P2, int T, int K,
                                                         int adaptive, float beta, float* __restrict__ est) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= P1) return;
  const float* d = dists + ((size_t)b * P1 + i) * K;
  const long long* id = idx + ((size_t)b * P1 + i) * K;
  float w[kKnnMax];
  int nb[kKnnMax];
  float mean = 0.f;
  for (int k = 0; k < K; ++k) mean += d[k];
  mean /= (float)K;
  const float rad = sqrtf(mean) + 1e-6f;
  const float r2 = rad * rad;
  const bool valid = len1 ? i < (int)len1[b] : true;
  float ws = 0.f;
  for (int k = 0; k < K; ++k) {
    float wk;
    if (adaptive) wk = d[k] <= r2 ? expf(-beta * d[k] / r2) : 0.f;
    else wk = expf(-beta * d[k]);
    if (!valid) wk = 0.f;
    w[k] = wk;
    ws += wk;
    nb[k] = (int)id[k];
  }
  const float inv = 1.f / (ws + 1e-8f);
  const float* sb = stat + (size_t)b * P2 * 3;
  for (int t = 0; t < T; ++t) {
    const float* mb = mov + ((size_t)b * T + t) * P2 * 3;
    float ex = 0.f, ey = 0.f, ez = 0.f;
    for (int k = 0; k < K; ++k) {
      const float wk = w[k] * inv;
      const size_t o = (size_t)nb[k] * 3;
      ex += wk * (mb[o] - sb[o]);
      ey += wk * (mb[o + 1] - sb[o + 1]);
      ez += wk * (mb[o + 2] - sb[o + 2]);
    }
    float* e = est + (((size_t)b * T + t) * P1 + i) * 3;
    e[0] = ex; e[1] = ey; e[2] = ez;
  }
}

// =======================================================================================
// LPIPS tail (reference utils/lpips/lpips.py:29-34, networks.py:60-62, utils.py:6-8) for one feature tap:
//     d[n] = mean_pixels sum_c w_c (fx_c / (||fx|| + eps) - fy_c / (||fy|| + eps))^2
// fx, fy fp16 [N, HW, C] (the channels-last activations cuDNN leaves), w fp32 [C].  The torch formulation (float copies,
// square, channel sum, sqrt, broadcast divide, difference, square, 1x1 conv, spatial mean -- each a pass over up to 1.7 GB,
// several through TensorIterator's strided slow path) cost 80 of the criterion's 146 ms on the joint train step's 2 x 50
// images; here it is one read of fx and fy forward, one read + one write backward.  One warp per pixel, the pixel's C
// channels in registers (C / 64 half2 per lane), deterministic: per-block partial sums, summed by the caller.
template <int C>
__global__ void __launch_bounds__(256) lpips_tap_fwd_kernel(const __half* __restrict__ fx, const __half* __restrict__ fy,
                                                            const float* __restrict__ w, int HW, float* __restrict__ partial) {
  constexpr int PER = C / 64;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n = blockIdx.y;
  float wr[2 * PER];
#pragma unroll
  for (int k = 0; k < PER; ++k) { wr[2 * k] = w[64 * k + 2 * lane]; wr[2 * k + 1] = w[64 * k + 2 * lane + 1]; }
  float acc = 0.f;
  for (int p = blockIdx.x * 8 + warp; p < HW; p += gridDim.x * 8) {
    const __half2* px = reinterpret_cast<const __half2*>(fx + ((size_t)n * HW + p) * C);
    const __half2* py = reinterpret_cast<const __half2*>(fy + ((size_t)n * HW + p) * C);
    float2 x[PER], y[PER];
    float sx = 0.f, sy = 0.f;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      x[k] = __half22float2(px[32 * k + lane]);
      y[k] = __half22float2(py[32 * k + lane]);
      sx += x[k].x * x[k].x + x[k].y * x[k].y;
      sy += y[k].x * y[k].x + y[k].y * y[k].y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sx += __shfl_xor_sync(0xffffffffu, sx, o); sy += __shfl_xor_sync(0xffffffffu, sy, o); }
    const float ix = 1.0f / (sqrtf(sx) + 1e-10f), iy = 1.0f / (sqrtf(sy) + 1e-10f);
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const float e0 = x[k].x * ix - y[k].x * iy, e1 = x[k].y * ix - y[k].y * iy;
      acc += wr[2 * k] * e0 * e0 + wr[2 * k + 1] * e1 * e1;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float red[8];
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];
    partial[(size_t)n * gridDim.x + blockIdx.x] = t;
  }
}
// d d[n] / d fx:  a = fx / (r + eps), e = a - b, u_c = 2 w_c e_c:
//     grad_fx_k = g[n] / HW * (u_k / (r + eps) - fx_k / (r (r + eps)^2) * sum_c u_c fx_c)
template <int C>
__global__ void __launch_bounds__(256) lpips_tap_bwd_kernel(const __half* __restrict__ fx, const __half* __restrict__ fy,
                                                            const float* __restrict__ w, const float* __restrict__ gout,
                                                            int HW, __half* __restrict__ gfx) {
  constexpr int PER = C / 64;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n = blockIdx.y;
  float wr[2 * PER];
#pragma unroll
  for (int k = 0; k < PER; ++k) { wr[2 * k] = w[64 * k + 2 * lane]; wr[2 * k + 1] = w[64 * k + 2 * lane + 1]; }
  const float g = gout[n] / (float)HW;
  for (int p = blockIdx.x * 8 + warp; p < HW; p += gridDim.x * 8) {
    const size_t base = ((size_t)n * HW + p) * C;
    const __half2* px = reinterpret_cast<const __half2*>(fx + base);
    const __half2* py = reinterpret_cast<const __half2*>(fy + base);
    float2 x[PER], y[PER];
    float sx = 0.f, sy = 0.f;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      x[k] = __half22float2(px[32 * k + lane]);
      y[k] = __half22float2(py[32 * k + lane]);
      sx += x[k].x * x[k].x + x[k].y * x[k].y;
      sy += y[k].x * y[k].x + y[k].y * y[k].y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sx += __shfl_xor_sync(0xffffffffu, sx, o); sy += __shfl_xor_sync(0xffffffffu, sy, o); }
    const float r = sqrtf(sx), ix = 1.0f / (r + 1e-10f), iy = 1.0f / (sqrtf(sy) + 1e-10f);
    float2 u[PER];
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      u[k].x = 2.0f * wr[2 * k] * (x[k].x * ix - y[k].x * iy);
      u[k].y = 2.0f * wr[2 * k + 1] * (x[k].y * ix - y[k].y * iy);
      dot += u[k].x * x[k].x + u[k].y * x[k].y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    const float cr = (r > 0.f) ? dot * ix * ix / r : 0.f;
    __half2* pg = reinterpret_cast<__half2*>(gfx + base);
#pragma unroll
    for (int k = 0; k < PER; ++k)
      pg[32 * k + lane] = __floats2half2_rn(g * (u[k].x * ix - x[k].x * cr), g * (u[k].y * ix - x[k].y * cr));
  }
}

// Glue between the VGG16 convolutions (cuDNN, channels-last fp16): bias + ReLU in place (torch adds a channels-last bias
// through TensorIterator's strided path: 16 ms per criterion call against 2 for this pass) and the 2 x 2 / stride 2 max pool
// with its backward (torch's NHWC pool kernels: 13.5 ms per call).  Thread = 8 channels of one pixel.
__global__ void __launch_bounds__(256) bias_relu_nhwc_kernel(__half* __restrict__ x, const float* __restrict__ bias, long long nvec,
                                                             int C8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nvec) return;
  const int c = (int)(i % C8) * 8;
  uint4 v = reinterpret_cast<uint4*>(x)[i];
  __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float2 f = __half22float2(h[j]);
    f.x = fmaxf(f.x + bias[c + 2 * j], 0.f);
    f.y = fmaxf(f.y + bias[c + 2 * j + 1], 0.f);
    h[j] = __floats2half2_rn(f.x, f.y);
  }
  reinterpret_cast<uint4*>(x)[i] = v;
}
__global__ void __launch_bounds__(256) maxpool2_nhwc_kernel(const __half* __restrict__ x, __half* __restrict__ y, int N, int Ho,
                                                            int Wo, int C8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)N * Ho * Wo * C8;
  if (i >= total) return;
  const int c = (int)(i % C8);
  long long p = i / C8;
  const int wo = (int)(p % Wo); p /= Wo;
  const int ho = (int)(p % Ho);
  const int n = (int)(p / Ho);
  const int W = 2 * Wo;
  const uint4* src = reinterpret_cast<const uint4*>(x) + (((long long)n * 2 * Ho + 2 * ho) * W + 2 * wo) * C8 + c;
  uint4 a = __ldg(src), b = __ldg(src + C8), d = __ldg(src + (long long)W * C8), e = __ldg(src + (long long)W * C8 + C8);
  __half2 *pa = reinterpret_cast<__half2*>(&a), *pb = reinterpret_cast<__half2*>(&b), *pd = reinterpret_cast<__half2*>(&d),
          *pe = reinterpret_cast<__half2*>(&e);
#pragma unroll
  for (int j = 0; j < 4; ++j) pa[j] = __hmax2(__hmax2(pa[j], pb[j]), __hmax2(pd[j], pe[j]));
  reinterpret_cast<uint4*>(y)[i] = a;
}
// gx = gy at the FIRST position of the window (row-major) whose value equals the maximum, 0 elsewhere (torch's choice)
__global__ void __launch_bounds__(256) maxpool2_nhwc_bwd_kernel(const __half* __restrict__ x, const __half* __restrict__ y,
                                                                const __half* __restrict__ gy, __half* __restrict__ gx, int N,
                                                                int Ho, int Wo, int C8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)N * Ho * Wo * C8;
  if (i >= total) return;
  const int c = (int)(i % C8);
  long long p = i / C8;
  const int wo = (int)(p % Wo); p /= Wo;
  const int ho = (int)(p % Ho);
  const int n = (int)(p / Ho);
  const int W = 2 * Wo;
  const long long base = (((long long)n * 2 * Ho + 2 * ho) * W + 2 * wo) * C8 + c;
  const long long off[4] = {0, C8, (long long)W * C8, (long long)W * C8 + C8};
  const uint4 m = __ldg(reinterpret_cast<const uint4*>(y) + i), g = __ldg(reinterpret_cast<const uint4*>(gy) + i);
  const unsigned short* pm = reinterpret_cast<const unsigned short*>(&m);
  const unsigned short* pg = reinterpret_cast<const unsigned short*>(&g);
  unsigned found = 0;                                    // bit j: channel j already assigned
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(x) + base + off[q]);
    const unsigned short* pv = reinterpret_cast<const unsigned short*>(&v);
    uint4 o;
    unsigned short* po = reinterpret_cast<unsigned short*>(&o);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bool hit = !((found >> j) & 1u) && __heq(__ushort_as_half(pv[j]), __ushort_as_half(pm[j]));
      po[j] = hit ? pg[j] : (unsigned short)0;
      found |= (hit ? 1u : 0u) << j;
    }
    reinterpret_cast<uint4*>(gx)[base + off[q]] = o;
  }
}

}  // namespace gvf

extern "C" GVF_API int gvf_bias_relu_nhwc_f16(void* x, const float* bias, long long pixels, int C, void* stream) {
  if (!x || !bias || pixels <= 0 || C <= 0 || (C % 8) || ((uintptr_t)x & 15)) return GVF_ERR_INVALID;
  const long long nvec = pixels * (C / 8);
  gvf::bias_relu_nhwc_kernel<<<(unsigned)((nvec + 255) / 256), 256, 0, (cudaStream_t)stream>>>((__half*)x, bias, nvec, C / 8);
  return cudaGetLastError() == cudaSuccess ? GVF_OK : GVF_ERR_CUDA;
}
extern "C" GVF_API int gvf_maxpool2_nhwc_f16(const void* x, void* y, int N, int H, int W, int C, void* stream) {
  if (!x || !y || N <= 0 || H <= 0 || W <= 0 || (H % 2) || (W % 2) || C <= 0 || (C % 8)) return GVF_ERR_INVALID;
  if (((uintptr_t)x | (uintptr_t)y) & 15) return GVF_ERR_INVALID;
  const long long total = (long long)N * (H / 2) * (W / 2) * (C / 8);
  gvf::maxpool2_nhwc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const __half*)x, (__half*)y, N, H / 2,
                                                                                             W / 2, C / 8);
  return cudaGetLastError() == cudaSuccess ? GVF_OK : GVF_ERR_CUDA;
}
extern "C" GVF_API int gvf_maxpool2_nhwc_bwd_f16(const void* x, const void* y, const void* gy, void* gx, int N, int H, int W, int C,
                                                 void* stream) {
  if (!x || !y || !gy || !gx || N <= 0 || H <= 0 || W <= 0 || (H % 2) || (W % 2) || C <= 0 || (C % 8)) return GVF_ERR_INVALID;
  if (((uintptr_t)x | (uintptr_t)y | (uintptr_t)gy | (uintptr_t)gx) & 15) return GVF_ERR_INVALID;
  const long long total = (long long)N * (H / 2) * (W / 2) * (C / 8);
  gvf::maxpool2_nhwc_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const __half*)x, (const __half*)y, (const __half*)gy, (__half*)gx, N, H / 2, W / 2, C / 8);
  return cudaGetLastError() == cudaSuccess ? GVF_OK : GVF_ERR_CUDA;
}

extern "C" GVF_API int gvf_lpips_tap_blocks(int HW) {
  int b = (HW + 63) / 64;                       // 8 pixels per pass and block, at least 8 passes
  return b < 1 ? 1 : (b > 296 ? 296 : b);
}
// partial fp32 [N, gvf_lpips_tap_blocks(HW)]: un-normalised block sums; d[n] = partial[n].sum() / HW
extern "C" GVF_API int gvf_lpips_tap_fwd(const void* fx, const void* fy, const float* w, int N, int HW, int C, float* partial,
                                         void* stream) {
  if (!fx || !fy || !w || !partial || N <= 0 || HW <= 0 || N > 65535) return GVF_ERR_INVALID;
  if (((uintptr_t)fx | (uintptr_t)fy) & 3) return GVF_ERR_INVALID;
  const dim3 grid(gvf_lpips_tap_blocks(HW), N);
  cudaStream_t st = (cudaStream_t)stream;
#define GVF_LP(CC) if (C == CC) { gvf::lpips_tap_fwd_kernel<CC><<<grid, 256, 0, st>>>((const __half*)fx, (const __half*)fy, w, HW, partial); \
                                  return cudaGetLastError() == cudaSuccess ? GVF_OK : GVF_ERR_CUDA; }
  GVF_LP(64) GVF_LP(128) GVF_LP(256) GVF_LP(512)
#undef GVF_LP
  return GVF_ERR_UNSUPPORTED;
}
// gout fp32 [N] = d loss / d d[n]; gfx fp16 [N, HW, C]
extern "C" GVF_API int gvf_lpips_tap_bwd(const void* fx, const void* fy, const float* w, const float* gout, int N, int HW, int C,
                                         void* gfx, void* stream) {
  if (!fx || !fy || !w || !gout || !gfx || N <= 0 || HW <= 0 || N > 65535) return GVF_ERR_INVALID;
  if (((uintptr_t)fx | (uintptr_t)fy | (uintptr_t)gfx) & 3) return GVF_ERR_INVALID;
  const dim3 grid(gvf_lpips_tap_blocks(HW), N);
  cudaStream_t st = (cudaStream_t)stream;
#define GVF_LP(CC) if (C == CC) { gvf::lpips_tap_bwd_kernel<CC><<<grid, 256, 0, st>>>((const __half*)fx, (const __half*)fy, w, gout, HW, (__half*)gfx); \
                                  return cudaGetLastError() == cudaSuccess ? GVF_OK : GVF_ERR_CUDA; }
  GVF_LP(64) GVF_LP(128) GVF_LP(256) GVF_LP(512)
#undef GVF_LP
  return GVF_ERR_UNSUPPORTED;
}

extern "C" GVF_API size_t gvf_ssim_l1_workspace_bytes(int planes, int H, int W) {
  if (planes <= 0 || H <= 0 || W <= 0) return 0;
  const size_t gx = (W + gvf::kTile - 1) / gvf::kTile, gy = (H + gvf::kTile - 1) / gvf::kTile;
  return (size_t)planes * gx * gy * 2 * sizeof(float);
}

extern "C" GVF_API int gvf_ssim_l1_fwd(const float* img1, const float* img2, int planes, int H, int W, float* workspace,
                                       size_t workspace_bytes, float* sums, float* dmaps, void* stream) {
  if (!img1 || !img2 || !workspace || !sums || planes <= 0 || H <= 0 || W <= 0 || planes > 65535) return GVF_ERR_INVALID;
  if (workspace_bytes < gvf_ssim_l1_workspace_bytes(planes, H, W)) return GVF_ERR_WORKSPACE;
  static const gvf::SsimWindow win = gvf::make_window();
  const dim3 grid((W + gvf::kTile - 1) / gvf::kTile, (H + gvf::kTile - 1) / gvf::kTile, planes);
  cudaStream_t st = (cudaStream_t)stream;
  gvf::ssim_l1_fwd_kernel<<<grid, 256, 0, st>>>(img1, img2, H, W, win, workspace, dmaps, planes);
  gvf::ssim_l1_finalize_kernel<<<planes, 256, 0, st>>>(workspace, (int)(grid.x * grid.y), sums);
  return cudaGetLastError() == cudaSuccess ? GVF_OK : GVF_ERR_CUDA;
}

extern "C" GVF_API int gvf_ssim_l1_bwd(const float* img1, const float* img2, const float* dmaps, int planes, int H, int W,
                                       const float* coef_ssim, const float* coef_l1, float* grad_img1, void* stream) {
  if (!img1 || !img2 || !dmaps || !coef_ssim || !coef_l1 || !grad_img1 || planes <= 0 || H <= 0 || W <= 0 ||
      planes > 65535)
    return GVF_ERR_INVALID;
  static const gvf::SsimWindow win = gvf::make_window();
  const dim3 grid((W + gvf::kTile - 1) / gvf::kTile, (H + gvf::kTile - 1) / gvf::kTile, planes);
  gvf::ssim_l1_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(img1, img2, dmaps, H, W, win, coef_ssim, coef_l1,
                                                                  grad_img1, planes);
  return cudaGetLastError() == cudaSuccess ? GVF_OK : GVF_ERR_CUDA;
}

extern "C" GVF_API int gvf_knn(const float* queries, const float* refs, int B, int P1, int P2, const long long* lengths1,
                               const long long* lengths2, int K, float* dists, long long* idx, void* stream) {
  if (!queries || !refs || !dists || !idx || B <= 0 || P1 <= 0 || P2 <= 0 || K <= 0) return GVF_ERR_INVALID;
  if (K > gvf::kKnnMax || B > 65535) return GVF_ERR_UNSUPPORTED;
  const dim3 grid((P1 + 255) / 256, B);
  cudaStream_t st = (cudaStream_t)stream;
  if (K <= 4) gvf::knn_kernel<4><<<grid, 256, 0, st>>>(queries, refs, P1, P2, lengths1, lengths2, K, dists, idx);
  else if (K <= 8) gvf::knn_kernel<8><<<grid, 256, 0, st>>>(queries, refs, P1, P2, lengths1, lengths2, K, dists, idx);
  else gvf::knn_kernel<16><<<grid, 256, 0, st>>>(queries, refs, P1, P2, lengths1, lengths2, K, dists, idx);
  return cudaGetLastError() == cudaSuccess ? GVF_OK : GVF_ERR_CUDA;
}

extern "C" GVF_API int gvf_knn_interp_deltas(const float* dists, const long long* idx, const float* static_pc,
                                             const float* moving_pc, const long long* lengths1, int B, int P1, int P2,
                                             int T, int K, int adaptive_radius, float beta, float* est, void* stream) {
  if (!dists || !idx || !static_pc || !moving_pc || !est || B <= 0 || P1 <= 0 || P2 <= 0 || T <= 0 || K <= 0)
    return GVF_ERR_INVALID;
  if (K > gvf::kKnnMax || B > 65535) return GVF_ERR_UNSUPPORTED;
  const dim3 grid((P1 + 255) / 256, B);
  gvf::knn_interp_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dists, idx, static_pc, moving_pc, lengths1, P1, P2, T,
                                                                 K, adaptive_radius, beta, est);
  return cudaGetLastError() == cudaSuccess ? GVF_OK : GVF_ERR_CUDA;
}
