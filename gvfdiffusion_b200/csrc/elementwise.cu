// elementwise.cu -- the small fused kernels around the GEMM / attention tiles of the DiT and
// motion-VAE paths: LayerNorm+adaLN modulate, per-head q/k RMS-norm, timestep embedding + all
// adaLN modulation vectors of a forward in two launches, small-K linears (K = 14 / 16),
// position embeddings, GEGLU, the final layer and the DPM-Solver++ state update.
//
// Each kernel names the reference lines it replaces.  Rounding points follow the fp16
// autocast the reference runs under (inference_dpm_latent.py:122-125): Linear outputs are
// fp16, LayerNorm / residual stream / sampler state fp32.  HBM-bound, one pass over the data.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gvf_b200.h"
#include "launch.h"
#include "tc_common.cuh"

namespace gvf {

__device__ __forceinline__ float r16f(float x) { return __half2float(__float2half_rn(x)); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float silu(float x) { return x / (1.0f + __expf(-x)); }

// ---------------------------------------------------------------------------------------
// LayerNorm (no affine or affine) + optional adaLN modulate -> fp16 GEMM operand.
// reference model/dit.py:246-247,254-255,263,268,273-274 (norm1..5 + modulate) and
// model/autoencoder.py:73-88 (PreNorm).  One warp per row, row kept in registers.
// ACT: SiLU on the normalised / modulated value before the single fp16 rounding (the `norm -> SiLU -> conv` pairs of the
// TRELLIS SparseResBlock3d, trellis/models/structured_latent_flow.py:57-62).
template <typename TIn, int C, bool VEC, bool ACT = false>
__global__ void __launch_bounds__(256) ln_mod_kernel(const TIn* __restrict__ x, __half* __restrict__ out,
                                                     int M, float eps, const float* __restrict__ w,
                                                     const float* __restrict__ bvec,
                                                     const __half* __restrict__ shift,
                                                     const __half* __restrict__ scale, int mod_stride,
                                                     int rows_per_batch) {
  constexpr int PER = C / 32;
  const int lane = threadIdx.x & 31;
  tc::pdl_wait();
  // grid-stride over rows: the grid is at most one resident wave (148 SMs x 8 CTAs), so there is no
  // partially filled second wave (12288 rows = 1536 CTAs of 8 rows were 1.3 waves)
  for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); row < M; row += gridDim.x * 8) {
  float v[PER];
  const TIn* xr = x + (size_t)row * C;
  if constexpr (!VEC) {
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      if constexpr (sizeof(TIn) == 4) v[i] = xr[i * 32 + lane];
      else v[i] = __half2float(xr[i * 32 + lane]);
    }
  } else
#pragma unroll
  for (int i = 0; i < PER / 4; ++i) {
    const int c = (i * 32 + lane) * 4;
    if constexpr (sizeof(TIn) == 4) {
      const float4 t = *reinterpret_cast<const float4*>(xr + c);
      v[i * 4] = t.x; v[i * 4 + 1] = t.y; v[i * 4 + 2] = t.z; v[i * 4 + 3] = t.w;
    } else {
      const uint2 t = *reinterpret_cast<const uint2*>(xr + c);
      const __half2 a = *reinterpret_cast<const __half2*>(&t.x), b2 = *reinterpret_cast<const __half2*>(&t.y);
      v[i * 4] = __low2float(a); v[i * 4 + 1] = __high2float(a);
      v[i * 4 + 2] = __low2float(b2); v[i * 4 + 3] = __high2float(b2);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) s += v[i];
  const float mean = warp_sum(s) * (1.0f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) { const float d = v[i] - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / C) + eps);
  const int b = (shift || scale) ? row / rows_per_batch : 0;
  __half* orow = out + (size_t)row * C;
  if constexpr (!VEC) {
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int c = i * 32 + lane;
      float y = (v[i] - mean) * rstd;
      if (w) y = y * w[c] + bvec[c];
      if (scale) y = y * (1.0f + __half2float(scale[(size_t)b * mod_stride + c])) +
                     __half2float(shift[(size_t)b * mod_stride + c]);
      if constexpr (ACT) y = y / (1.0f + __expf(-y));
      orow[c] = __float2half_rn(y);
    }
  } else
#pragma unroll
  for (int i = 0; i < PER / 4; ++i) {
    const int c = (i * 32 + lane) * 4;
    __align__(8) __half h[4];
#pragma unroll
    float wv[4] = {1.f, 1.f, 1.f, 1.f}, bv[4] = {0.f, 0.f, 0.f, 0.f}, sc[4] = {0.f, 0.f, 0.f, 0.f}, sh[4] = {0.f, 0.f, 0.f, 0.f};
    if (w) {
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(w + c)), b4 = __ldg(reinterpret_cast<const float4*>(bvec + c));
      wv[0] = w4.x; wv[1] = w4.y; wv[2] = w4.z; wv[3] = w4.w;
      bv[0] = b4.x; bv[1] = b4.y; bv[2] = b4.z; bv[3] = b4.w;
    }
    if (scale) {                                     // mod_stride and the modulation offsets are multiples of 4
      const uint2 s2 = *reinterpret_cast<const uint2*>(scale + (size_t)b * mod_stride + c);
      const uint2 h2 = *reinterpret_cast<const uint2*>(shift + (size_t)b * mod_stride + c);
      const __half* sp = reinterpret_cast<const __half*>(&s2);
      const __half* hp = reinterpret_cast<const __half*>(&h2);
#pragma unroll
      for (int t = 0; t < 4; ++t) { sc[t] = __half2float(sp[t]); sh[t] = __half2float(hp[t]); }
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float y = (v[i * 4 + t] - mean) * rstd;
      if (w) y = y * wv[t] + bv[t];
      if (scale) y = y * (1.0f + sc[t]) + sh[t];
      if constexpr (ACT) y = y / (1.0f + __expf(-y));
      h[t] = __float2half_rn(y);
    }
    *reinterpret_cast<uint2*>(orow + c) = *reinterpret_cast<uint2*>(h);
  }
  }
}

// Two rows per warp IN FLIGHT (the vector path of ln_mod_kernel, same per-row arithmetic: identical bits): a warp issues the
// loads of both of its rows before it reduces either, so that 6144 warps cover the DiT's 12288 rows in one pass instead of
// 30 % of 9472 warps walking a second row.  An experiment that LOST its A/B (see g_ln_two_rows below); opt-in only.
template <typename TIn, int C>
__global__ void __launch_bounds__(256) ln_mod2_kernel(const TIn* __restrict__ x, __half* __restrict__ out, int M, float eps,
                                                      const float* __restrict__ w, const float* __restrict__ bvec,
                                                      const __half* __restrict__ shift, const __half* __restrict__ scale,
                                                      int mod_stride, int rows_per_batch) {
  constexpr int PER = C / 32, RPW = 2;
  const int lane = threadIdx.x & 31;
  tc::pdl_wait();
  for (int row0 = (blockIdx.x * 8 + (threadIdx.x >> 5)) * RPW; row0 < M; row0 += gridDim.x * 8 * RPW) {
    float v[RPW][PER];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int row = row0 + r < M ? row0 + r : M - 1;
      const TIn* xr = x + (size_t)row * C;
#pragma unroll
      for (int i = 0; i < PER / 4; ++i) {
        const int c = (i * 32 + lane) * 4;
        if constexpr (sizeof(TIn) == 4) {
          const float4 t = *reinterpret_cast<const float4*>(xr + c);
          v[r][i * 4] = t.x; v[r][i * 4 + 1] = t.y; v[r][i * 4 + 2] = t.z; v[r][i * 4 + 3] = t.w;
        } else {
          const uint2 t = *reinterpret_cast<const uint2*>(xr + c);
          const __half2 a = *reinterpret_cast<const __half2*>(&t.x), b2 = *reinterpret_cast<const __half2*>(&t.y);
          v[r][i * 4] = __low2float(a); v[r][i * 4 + 1] = __high2float(a);
          v[r][i * 4 + 2] = __low2float(b2); v[r][i * 4 + 3] = __high2float(b2);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int row = row0 + r;
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < PER; ++i) s += v[r][i];
      const float mean = warp_sum(s) * (1.0f / C);
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < PER; ++i) { const float d = v[r][i] - mean; q += d * d; }
      const float rstd = rsqrtf(warp_sum(q) * (1.0f / C) + eps);
      if (row >= M) continue;
      const int b = (shift || scale) ? row / rows_per_batch : 0;
      __half* orow = out + (size_t)row * C;
#pragma unroll
      for (int i = 0; i < PER / 4; ++i) {
        const int c = (i * 32 + lane) * 4;
        __align__(8) __half h[4];
        float wv[4] = {1.f, 1.f, 1.f, 1.f}, bv[4] = {0.f, 0.f, 0.f, 0.f}, sc[4] = {0.f, 0.f, 0.f, 0.f}, sh[4] = {0.f, 0.f, 0.f, 0.f};
        if (w) {
          const float4 w4 = __ldg(reinterpret_cast<const float4*>(w + c)), b4 = __ldg(reinterpret_cast<const float4*>(bvec + c));
          wv[0] = w4.x; wv[1] = w4.y; wv[2] = w4.z; wv[3] = w4.w;
          bv[0] = b4.x; bv[1] = b4.y; bv[2] = b4.z; bv[3] = b4.w;
        }
        if (scale) {
          const uint2 s2 = *reinterpret_cast<const uint2*>(scale + (size_t)b * mod_stride + c);
          const uint2 h2 = *reinterpret_cast<const uint2*>(shift + (size_t)b * mod_stride + c);
          const __half* sp = reinterpret_cast<const __half*>(&s2);
          const __half* hp = reinterpret_cast<const __half*>(&h2);
#pragma unroll
          for (int t = 0; t < 4; ++t) { sc[t] = __half2float(sp[t]); sh[t] = __half2float(hp[t]); }
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          float y = (v[r][i * 4 + t] - mean) * rstd;
          if (w) y = y * wv[t] + bv[t];
          if (scale) y = y * (1.0f + sc[t]) + sh[t];
          h[t] = __float2half_rn(y);
        }
        *reinterpret_cast<uint2*>(orow + c) = *reinterpret_cast<uint2*>(h);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// MultiHeadRMSNorm on q and k in place (reference model/attention/modules.py:8-15,122-125):
// x <- fp16( x.float() / max(||x||, 1e-12) * gamma[h] * sqrt(d) ).  buf rows of `ld` halfs;
// q heads start at column 0, k heads at column k_off.  One thread per (row, head, q|k).
template <int D>
__global__ void __launch_bounds__(256) rmsnorm_heads_kernel(__half* __restrict__ buf, long long rows, int ld,
                                                            int H, int k_off,
                                                            const float* __restrict__ gq,
                                                            const float* __restrict__ gk) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= rows * H * 2) return;
  const int which = (int)(gid % 2);
  const int h = (int)((gid / 2) % H);
  const long long row = gid / (2 * H);
  __half* p = buf + row * ld + (which ? k_off : 0) + h * D;
  const float* g = (which ? gk : gq) + h * D;
  float v[D];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < D / 8; ++i) {
    const uint4 t = *reinterpret_cast<const uint4*>(p + i * 8);
    const __half* hh = reinterpret_cast<const __half*>(&t);
#pragma unroll
    for (int j = 0; j < 8; ++j) { v[i * 8 + j] = __half2float(hh[j]); ss += v[i * 8 + j] * v[i * 8 + j]; }
  }
  const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
  const float sq = sqrtf((float)D);
#pragma unroll
  for (int i = 0; i < D / 8; ++i) {
    __align__(16) __half o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = __float2half_rn(v[i * 8 + j] * inv * g[i * 8 + j] * sq);
    *reinterpret_cast<uint4*>(p + i * 8) = *reinterpret_cast<uint4*>(o);
  }
}

// Head dim 64 (the TRELLIS structured-latent flow blocks, 3656 x 3072 rows: the thread-per-head form above reads 128 B per
// thread with a 128 B stride between lanes -- 47 us for 22 MB): eight lanes per head, one 16 B piece each, so a warp reads
// 512 contiguous bytes; the sum of squares is reduced over the eight lanes by shuffles.
__global__ void __launch_bounds__(256) rmsnorm_heads64_kernel(__half* __restrict__ buf, long long rows, int ld, int H, int k_off,
                                                              const float* __restrict__ gq, const float* __restrict__ gk) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = rows * H * 2 * 8;
  const bool valid = gid < total;
  const long long g = valid ? gid : 0;
  const int part = (int)(g & 7);
  const int h = (int)((g >> 3) % H);
  const int which = (int)(((g >> 3) / H) & 1);
  const long long row = (g >> 3) / (2 * H);
  __half* p = buf + row * ld + (which ? k_off : 0) + h * 64 + part * 8;
  const float* gm = (which ? gk : gq) + h * 64 + part * 8;
  const uint4 t = *reinterpret_cast<const uint4*>(p);
  const __half* hh = reinterpret_cast<const __half*>(&t);
  float v[8];
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { v[j] = __half2float(hh[j]); ss += v[j] * v[j]; }
  ss += __shfl_xor_sync(0xffffffffu, ss, 1);
  ss += __shfl_xor_sync(0xffffffffu, ss, 2);
  ss += __shfl_xor_sync(0xffffffffu, ss, 4);
  const float inv = 8.0f / fmaxf(sqrtf(ss), 1e-12f);          // sqrt(64) / max(||x||, eps)
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(gm)), g1 = __ldg(reinterpret_cast<const float4*>(gm + 4));
  const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
  __align__(16) __half o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = __float2half_rn(v[j] * inv * gv[j]);
  if (valid) *reinterpret_cast<uint4*>(p) = *reinterpret_cast<uint4*>(o);
}

// ---------------------------------------------------------------------------------------
// TimestepEmbedder (reference model/dit.py:59-100): t[B] -> silu(t_emb)[B,C] as fp16 values.
// One CTA per batch element, C threads.  W0 [C,256], W2 [C,C] fp16; biases fp32 (fp16-valued).
__global__ void __launch_bounds__(1024) temb_kernel(const float* __restrict__ t, const __half* __restrict__ W0,
                                                    const float* __restrict__ b0, const __half* __restrict__ W2,
                                                    const float* __restrict__ b2, int C, int F,
                                                    __half* __restrict__ temb_out, __half* __restrict__ silu_out) {
  extern __shared__ float sh[];
  float* emb = sh;        // F
  float* hid = sh + F;    // C
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
  const float tv = t[b];
  const int half = F / 2;
  for (int i = tid; i < half; i += blockDim.x) {
    const float freq = expf(-9.210340371976184f * (float)i / (float)half);
    const float a = tv * freq;
    emb[i] = r16f(cosf(a));
    emb[half + i] = r16f(sinf(a));
  }
  __syncthreads();
  // one warp per output row, 16 B weight loads (the scalar version of these two mat-vecs took 64 us per NFE
  // on its single SM); F and C are multiples of 8
  auto dot_row = [&](const __half* wr, const float* x, int n) {
    float acc = 0.f;
    for (int i = lane * 8; i < n; i += 256) {
      const uint4 tw = __ldg(reinterpret_cast<const uint4*>(wr + i));
      const __half2* h2 = reinterpret_cast<const __half2*>(&tw);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float2 f = __half22float2(h2[u]);
        acc = fmaf(f.x, x[i + 2 * u], acc);
        acc = fmaf(f.y, x[i + 2 * u + 1], acc);
      }
    }
    return warp_sum(acc);
  };
#pragma unroll 4
  for (int j = w; j < C; j += nw) {
    const float acc = dot_row(W0 + (size_t)j * F, emb, F);
    if (lane == 0) hid[j] = r16f(silu(r16f(acc + b0[j])));
  }
  __syncthreads();
#pragma unroll 4
  for (int j = w; j < C; j += nw) {
    const float acc = dot_row(W2 + (size_t)j * C, hid, C);
    if (lane == 0) {
      const float te = r16f(acc + b2[j]);
      temb_out[(size_t)b * C + j] = __float2half_rn(te);
      silu_out[(size_t)b * C + j] = __float2half_rn(silu(te));
    }
  }
}

// All adaLN modulation vectors of one forward: out[b, r] = fp16( W[r,:] . s[b,:] + bias[r] ),
// W = the 12 x (adaLN_modulation.1 | adaLN_modulation_temporal.1) + final adaLN rows
// concatenated [R, C] (reference model/dit.py:240-242,299).  One warp per output row.
template <int MAXB>
__global__ void __launch_bounds__(256) mod_gemv_kernel(const __half* __restrict__ W, const float* __restrict__ bias,
                                                       const __half* __restrict__ s, int B, int R, int C,
                                                       __half* __restrict__ out) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  float acc[MAXB];
#pragma unroll
  for (int b = 0; b < MAXB; ++b) acc[b] = 0.f;
  const __half* wr = W + (size_t)r * C;
  for (int c = lane * 8; c < C; c += 256) {
    const uint4 t = *reinterpret_cast<const uint4*>(wr + c);
    const __half* wh = reinterpret_cast<const __half*>(&t);
#pragma unroll
    for (int b = 0; b < MAXB; ++b) {
      if (b < B) {
        const uint4 u = *reinterpret_cast<const uint4*>(s + (size_t)b * C + c);
        const __half* sh = reinterpret_cast<const __half*>(&u);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[b] += __half2float(wh[k]) * __half2float(sh[k]);
      }
    }
  }
#pragma unroll
  for (int b = 0; b < MAXB; ++b) {
    if (b < B) {
      const float v = warp_sum(acc[b]);
      if (lane == 0) out[(size_t)b * R + r] = __float2half_rn(v + bias[r]);
    }
  }
}

// ---------------------------------------------------------------------------------------
// Small-K linear on CUDA cores: y[m, n] = fp16( x[m,:K] . W[n,:K] + b[n] ) (+ add[m % add_rows, n])
// used for input_layer (16->512, + APE; reference model/dit.py:457,470-472), static_cond_proj
// (14->512, :465), VAE proj (16->768, autoencoder.py:585) and gs_embedding (14->768, :389).
// x fp32 (rounded to fp16 on load like autocast does), W fp16 [N,K], out fp32 or fp16.
// RPB rows per block: a thread loads its weight row once per block, so 8 rows (the first form) re-read the 16 KB of weights
// 1536 times at M = 12288 with 32 B-strided loads and ran 48 us for a 25 MB pass; 64 rows per block for large M make the
// pass bandwidth-shaped.  Per-output arithmetic (k ascending, then bias, rounding, + add) is unchanged: identical bits.
template <typename TOut, int RPB>
__global__ void __launch_bounds__(256) small_linear_kernel(const float* __restrict__ x, int ldx,
                                                           const __half* __restrict__ W,
                                                           const float* __restrict__ bias, int M, int N,
                                                           int K, const float* __restrict__ add,
                                                           int add_rows, TOut* __restrict__ out) {
  __shared__ float xs[RPB][32];
  const int row0 = blockIdx.y * RPB;
  for (int i = threadIdx.x; i < RPB * K; i += 256) {
    const int r = i / K, k = i - r * K;
    xs[r][k] = (row0 + r < M) ? r16f(x[(size_t)(row0 + r) * ldx + k]) : 0.f;
  }
  __syncthreads();
  const int n = blockIdx.x * 256 + threadIdx.x;
  if (n >= N) return;
  float wv[32];
  for (int k = 0; k < K; ++k) wv[k] = __half2float(W[(size_t)n * K + k]);
  const float bn = bias ? bias[n] : 0.f;
  int ar = row0 % add_rows;
  // four rows per trip: the four `add` loads are issued before any of them is needed (the loop is latency-bound at the
  // ~20 warps per SM a 384-block grid leaves resident), per-output arithmetic unchanged
  constexpr int U = RPB >= 4 ? 4 : 1;
  for (int r = 0; r < RPB && row0 + r < M; r += U) {
    float av[U];
    int arr = ar;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      av[u] = (add && row0 + r + u < M) ? __ldg(add + (size_t)arr * N + n) : 0.f;
      if (++arr == add_rows) arr = 0;
    }
    ar = arr;
    float acc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) acc[u] = 0.f;
    for (int k = 0; k < K; ++k) {
#pragma unroll
      for (int u = 0; u < U; ++u) acc[u] += xs[r + u][k] * wv[k];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (row0 + r + u >= M) break;
      float y = r16f(acc[u] + bn);
      if (add) y += av[u];
      if constexpr (sizeof(TOut) == 4) out[(size_t)(row0 + r + u) * N + n] = y;
      else out[(size_t)(row0 + r + u) * N + n] = __float2half_rn(y);
    }
  }
}

// AbsolutePositionEmbedder (reference model/dit.py:16-56): xyz [R,3] -> [R, C] fp32:
// per coordinate [sin(fd), cos(fd)], fd = C/3/2, freq_j = 10000^(-j/fd); zero padded to C.
__global__ void ape_kernel(const float* __restrict__ xyz, int R, int C, float* __restrict__ out) {
  const int fd = C / 3 / 2;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)R * C) return;
  const int r = (int)(gid / C), c = (int)(gid - (long long)r * C);
  float v = 0.f;
  if (c < 6 * fd) {
    const int coord = c / (2 * fd), w = c - coord * 2 * fd;
    const int j = w < fd ? w : w - fd;
    const float freq = 1.0f / powf(10000.0f, (float)j / (float)fd);
    const float a = xyz[(size_t)r * 3 + coord] * freq;
    v = w < fd ? sinf(a) : cosf(a);
  }
  out[gid] = v;
}

// VAE query embedding (reference model/autoencoder.py:250-301,389-391,560):
//   out = fp16( LN_1e-6( LN_1e-5(gs[q]) + LN_1e-5(PointEmbed(q.xyz)) ) )
// gs [Q, C] fp16 = Linear(14->C)(q) (small_linear above).  PointEmbed runs under fp16 autocast
// in the reference (einsum is autocast-to-fp16): coordinate, omega and their product are
// rounded to fp16 before sin/cos, whose results are rounded to fp16 too.
// omega_j = 10000^(-j / (E/2)), E = C/6, computed in float64 then rounded.  Warp per query.
// FINAL = false: the sum LN(gs) + LN(PE(xyz)) itself (the encoder's token embedding, model/autoencoder.py:529-533,
// which enters a residual stream before its PreNorm); `rows_per_xyz` consecutive output rows share one xyz row
// (the T frames of a point).
template <int C, bool FINAL = true>
__global__ void __launch_bounds__(256) query_embed_kernel(const float* __restrict__ queries, int ldq,
                                                          const __half* __restrict__ gs, int Q,
                                                          __half* __restrict__ out, const int* __restrict__ xyz_row = nullptr) {
  constexpr int PER = C / 32, E = C / 6;
  const int qi = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (qi >= Q) return;
  float g[PER], p[PER];
  float sg = 0.f, sp = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = i * 32 + lane;
    g[i] = __half2float(gs[(size_t)qi * C + c]);
    const int coord = c / (2 * E), w = c - coord * 2 * E;
    const int j = w < E ? w : w - E;
    const double om = 1.0 / pow(10000.0, (double)j / ((double)E / 2.0));
    const size_t qrow = xyz_row ? (size_t)__ldg(xyz_row + qi) : (size_t)qi;
    const float x16 = r16f(queries[qrow * ldq + coord]);
    const float a = r16f(x16 * r16f((float)om));
    p[i] = r16f(w < E ? sinf(a) : cosf(a));
    sg += g[i];
    sp += p[i];
  }
  const float mg = warp_sum(sg) / C, mp = warp_sum(sp) / C;
  float vg = 0.f, vp = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) { vg += (g[i] - mg) * (g[i] - mg); vp += (p[i] - mp) * (p[i] - mp); }
  const float rg = rsqrtf(warp_sum(vg) / C + 1e-5f), rp = rsqrtf(warp_sum(vp) / C + 1e-5f);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) { g[i] = (g[i] - mg) * rg + (p[i] - mp) * rp; s += g[i]; }
  if constexpr (!FINAL) {                          // fp32: LayerNorm outputs stay fp32 under autocast, and so does their sum
    float* o32 = reinterpret_cast<float*>(out);
#pragma unroll
    for (int i = 0; i < PER; ++i) o32[(size_t)qi * C + i * 32 + lane] = g[i];
    return;
  }
  const float m = warp_sum(s) / C;
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) v += (g[i] - m) * (g[i] - m);
  const float rs = rsqrtf(warp_sum(v) / C + 1e-6f);
#pragma unroll
  for (int i = 0; i < PER; ++i) out[(size_t)qi * C + i * 32 + lane] = __float2half_rn((g[i] - m) * rs);
}

// GEGLU (reference model/autoencoder.py:90-93): h [M, 2F] fp16 -> out [M, F] = a * gelu_erf(gate)
__global__ void __launch_bounds__(256) geglu_kernel(const __half* __restrict__ h, long long M, int F,
                                                    __half* __restrict__ out) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n8 = M * (F / 8);
  if (gid >= n8) return;
  const long long row = gid / (F / 8);
  const int c = (int)(gid - row * (F / 8)) * 8;
  const uint4 a = *reinterpret_cast<const uint4*>(h + row * 2 * F + c);
  const uint4 g = *reinterpret_cast<const uint4*>(h + row * 2 * F + F + c);
  const __half* ah = reinterpret_cast<const __half*>(&a);
  const __half* gh = reinterpret_cast<const __half*>(&g);
  __align__(16) __half o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float x = __half2float(gh[j]);
    const float ge = r16f(0.5f * x * (1.0f + erff(x * 0.7071067811865476f)));
    o[j] = __float2half_rn(__half2float(ah[j]) * ge);
  }
  *reinterpret_cast<uint4*>(out + row * F + c) = *reinterpret_cast<uint4*>(o);
}

__global__ void __launch_bounds__(256) cast_f16_kernel(const float* __restrict__ x, long long n,
                                                       __half* __restrict__ out) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    const float4 a = *reinterpret_cast<const float4*>(x + i), b = *reinterpret_cast<const float4*>(x + i + 4);
    __align__(16) __half h[8] = {__float2half_rn(a.x), __float2half_rn(a.y), __float2half_rn(a.z),
                                 __float2half_rn(a.w), __float2half_rn(b.x), __float2half_rn(b.y),
                                 __float2half_rn(b.z), __float2half_rn(b.w)};
    *reinterpret_cast<uint4*>(out + i) = *reinterpret_cast<uint4*>(h);
  } else {
    for (long long j = i; j < n; ++j) out[j] = __float2half_rn(x[j]);
  }
}

// ---------------------------------------------------------------------------------------
// FinalLayer (reference model/dit.py:287-303): LN -> modulate(shift, scale) -> Linear(C -> O<=16)
// X fp32 [M,C] -> out fp32 [M,O] (fp16-valued, as the autocast module returns).  Warp per row.
// Four rows per warp: a weight value loaded once serves four rows (the one-row form issued 256 loads + 80 shuffles per row
// and lane: 46 us for a 25 MB pass).  The O = 16 dot products of a row are reduced together by a transposed butterfly --
// at offset 16 every lane keeps half of the outputs and hands the other half to its partner, and so on: 8 + 4 + 2 + 1 + 1
// shuffles instead of 80 -- which adds the same pairs in the same order as sixteen warp_sum trees: identical bits.
template <int C, int O>
__global__ void __launch_bounds__(256) final_layer_kernel(const float* __restrict__ x, int M,
                                                          const __half* __restrict__ shift,
                                                          const __half* __restrict__ scale, int mod_stride,
                                                          int rows_per_batch, const __half* __restrict__ W,
                                                          const float* __restrict__ bias,
                                                          float* __restrict__ out) {
  static_assert(O == 16, "the transposed reduction is written for 16 outputs");
  constexpr int PER = C / 32, RW = 4;
  const int lane = threadIdx.x & 31;
  const int row0 = (blockIdx.x * 8 + (threadIdx.x >> 5)) * RW;
  if (row0 >= M) return;
  float v[RW][PER];
#pragma unroll
  for (int r = 0; r < RW; ++r) {
    const int row = row0 + r < M ? row0 + r : M - 1;           // tail rows: recompute the last row, not stored
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) { v[r][i] = x[(size_t)row * C + i * 32 + lane]; s += v[r][i]; }
    const float mean = warp_sum(s) / C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) q += (v[r][i] - mean) * (v[r][i] - mean);
    const float rstd = rsqrtf(warp_sum(q) / C + 1e-6f);
    const int b = row / rows_per_batch;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int c = i * 32 + lane;
      const float y = (v[r][i] - mean) * rstd * (1.0f + __half2float(scale[(size_t)b * mod_stride + c])) +
                      __half2float(shift[(size_t)b * mod_stride + c]);
      v[r][i] = r16f(y);
    }
  }
  float acc[RW][O];
#pragma unroll
  for (int r = 0; r < RW; ++r)
#pragma unroll
    for (int o = 0; o < O; ++o) acc[r][o] = 0.f;
#pragma unroll
  for (int o = 0; o < O; ++o) {
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const float w = __half2float(W[(size_t)o * C + i * 32 + lane]);
#pragma unroll
      for (int r = 0; r < RW; ++r) acc[r][o] += v[r][i] * w;
    }
  }
  // lane l ends with output o(l) = 8 [l & 16] + 4 [l & 8] + 2 [l & 4] + [l & 2]
  const int o_mine = ((lane & 16) ? 8 : 0) + ((lane & 8) ? 4 : 0) + ((lane & 4) ? 2 : 0) + ((lane & 2) ? 1 : 0);
  const float bo = bias[o_mine];
#pragma unroll
  for (int r = 0; r < RW; ++r) {
#pragma unroll
    for (int off = 16, n = O; off >= 2; off >>= 1, n >>= 1) {
      const bool upper = (lane & off) != 0;
#pragma unroll
      for (int j = 0; j < n / 2; ++j) {
        const float send = upper ? acc[r][j] : acc[r][j + n / 2];
        const float keep = upper ? acc[r][j + n / 2] : acc[r][j];
        acc[r][j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      }
    }
    const float tot = acc[r][0] + __shfl_xor_sync(0xffffffffu, acc[r][0], 1);
    if (!(lane & 1) && row0 + r < M) out[(size_t)(row0 + r) * O + o_mine] = r16f(tot + bo);
  }
}

// ---------------------------------------------------------------------------------------
// DPM-Solver++ data prediction with the reference's model wrapper folded in
// (reference model/dpmsolver.py:284-300 v -> eps, :328-347 3-way CFG, :450-459 x0):
//   eps_k = alpha * v_k + sigma * x ;  eps = eps_fu + s1 (eps_u - eps_fu) + s2 (eps_c - eps_u)
//   x0 = (x - sigma * eps) / alpha
// v: [branches, n] (branch order full-uncond, image-uncond, cond; 1 branch = cond only).
__global__ void __launch_bounds__(256) dpm_x0_kernel(const float* __restrict__ x, const float* __restrict__ v,
                                                     long long n, int branches, int vpred, float alpha,
                                                     float sigma, float s1, float s2, float* __restrict__ x0) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float xv = x[i];
  float eps;
  if (branches == 1) {
    eps = vpred ? alpha * v[i] + sigma * xv : v[i];
  } else {
    const float efu = vpred ? alpha * v[i] + sigma * xv : v[i];
    const float eu = vpred ? alpha * v[n + i] + sigma * xv : v[n + i];
    const float ec = vpred ? alpha * v[2 * n + i] + sigma * xv : v[2 * n + i];
    eps = efu + s1 * (eu - efu) + s2 * (ec - eu);
  }
  x0[i] = (xv - sigma * eps) / alpha;
}

__global__ void __launch_bounds__(256) dpm_error_sq_kernel(const float* __restrict__ xh, const float* __restrict__ xl,
                                                           const float* __restrict__ xp, long long n_per_batch,
                                                           float atol, float rtol, float* __restrict__ E2) {
  const int b = blockIdx.y;
  const float* h = xh + (size_t)b * n_per_batch;
  const float* l = xl + (size_t)b * n_per_batch;
  const float* p = xp + (size_t)b * n_per_batch;
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_per_batch;
       i += (long long)gridDim.x * blockDim.x) {
    const float delta = fmaxf(atol, rtol * fmaxf(fabsf(l[i]), fabsf(p[i])));
    const float e = (h[i] - l[i]) / delta;
    acc += e * e;
  }
  acc = warp_sum(acc);
  __shared__ float ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += ws[i];
    atomicAdd(E2 + b, t);
  }
}

// out = a*x + b*y + c*z  (y, z optional) -- first / second order updates written exactly in the
// reference's evaluation order by the host (model/dpmsolver.py:589-597, 843-848).
//   first order : x_t = (sig_t/sig_s) x - (alpha_t phi1) m0
//   second order: x_t = (sig_t/sig_0) x - (alpha_t phi1) m0 - 0.5 (alpha_t phi1) * ((1/r0) (m0 - m1))
__global__ void __launch_bounds__(256) dpm_update_kernel(const float* __restrict__ x,
                                                         const float* __restrict__ m0,
                                                         const float* __restrict__ m1, long long n,
                                                         float cx, float cm, float inv_r0, int order,
                                                         float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float r = cx * x[i] - cm * m0[i];
  if (order == 2) {
    const float d1 = inv_r0 * (m0[i] - m1[i]);
    r = r - 0.5f * cm * d1;
  }
  out[i] = r;
}

// out = x * a[c] + b[c]  over the last dim (latent de-normalisation, inference_dpm_latent.py:250)
__global__ void __launch_bounds__(256) affine_lastdim_kernel(const float* __restrict__ x, long long n, int C,
                                                             const float* __restrict__ a,
                                                             const float* __restrict__ b, float as, float bs,
                                                             float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = (int)(i % C);
  out[i] = x[i] * (a ? a[c] : as) + (b ? b[c] : bs);
}

// Euler step of the TRELLIS flow samplers (trellis/pipelines/samplers/flow_euler.py:36-77 + the guidance mixins):
//   v = (1 + s) v_cond - s v_neg (when v_neg is given);  x_prev = x - (t - t_prev) v;
//   x0 = (1 - sigma_min) x - (sigma_min + (1 - sigma_min) t) v
// Every product and difference is rounded on its own (no FMA contraction), in the reference's operation order, so the
// result equals the torch expressions bit for bit.
// The scalar coefficients arrive as floats the HOST rounded from its double expressions (1 + s, 1 - sigma_min,
// sigma_min + (1 - sigma_min) t), exactly what torch does with a Python scalar operand.
__global__ void __launch_bounds__(256) flow_euler_kernel(const float* __restrict__ x, const float* __restrict__ v,
                                                         const float* __restrict__ vn, long long n, float s1, float s,
                                                         float dt, float c0, float c1, float* __restrict__ xp,
                                                         float* __restrict__ x0) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float vv = v[i];
  if (vn) vv = __fsub_rn(__fmul_rn(s1, vv), __fmul_rn(s, vn[i]));
  const float xi = x[i];
  xp[i] = __fsub_rn(xi, __fmul_rn(dt, vv));
  if (x0) x0[i] = __fsub_rn(__fmul_rn(c0, xi), __fmul_rn(c1, vv));
}

}  // namespace gvf

using namespace gvf;
#define ST(s) ((cudaStream_t)(s))
#define RET() return cudaGetLastError() == cudaSuccess ? GVF_OK : GVF_ERR_CUDA

extern "C" {

// MEASURED (tools/ln_bench.py, 50 launches per graph): one row per warp 6.52 us (5.8 TB/s) at 12288 x 512, two rows in
// flight 7.48 us; 8.66 vs 10.66 at width 768, 5.00 vs 6.48 at 3656 x 1024 -- the one-row kernel already runs at the HBM
// rate and the wider per-warp footprint only costs occupancy.  Opt-in (gvf_ln_set_two_rows(1)), default off.
static int g_ln_two_rows = 0;
static int ln_mod_launch(const void* x, int x_is_f16, void* out, int M, int C, float eps, const float* w, const float* b,
                         const void* shift, const void* scale, int mod_stride, int rows_per_batch, int act, void* stream) {
  if (!x || !out || M <= 0) return GVF_ERR_INVALID;
  if ((w == nullptr) != (b == nullptr) || (shift == nullptr) != (scale == nullptr)) return GVF_ERR_INVALID;
  const int rpb = rows_per_batch > 0 ? rows_per_batch : M;
  // the vector path reads 4 modulation / affine values at a time
  if (shift && ((mod_stride % 4) || ((uintptr_t)shift & 7) || ((uintptr_t)scale & 7))) return GVF_ERR_INVALID;
  if (w && (((uintptr_t)w | (uintptr_t)b) & 15)) return GVF_ERR_INVALID;
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  if (!act && g_ln_two_rows && M >= 2048 && (C == 512 || C == 768 || C == 1024)) {
    const int want2 = (M + 15) / 16, cap2 = num_sms * 8;
    const dim3 grid2(want2 < cap2 ? want2 : cap2);
#define LN2(T, CC) launch_pdl(ln_mod2_kernel<T, CC>, grid2, dim3(256), 0, ST(stream), (const T*)x, (__half*)out, M, eps, w, b, \
                              (const __half*)shift, (const __half*)scale, mod_stride, rpb)
    if (C == 512) { if (x_is_f16) LN2(__half, 512); else LN2(float, 512); }
    else if (C == 768) { if (x_is_f16) LN2(__half, 768); else LN2(float, 768); }
    else { if (x_is_f16) LN2(__half, 1024); else LN2(float, 1024); }
#undef LN2
    RET();
  }
  const int want = (M + 7) / 8, cap = num_sms * 8;
  const dim3 grid(want < cap ? want : cap);
#define LN_CASE(T, CC, V, A)                                                                          \
  launch_pdl(ln_mod_kernel<T, CC, V, A>, grid, dim3(256), 0, ST(stream), (const T*)x, (__half*)out, M, eps, w, b, \
             (const __half*)shift, (const __half*)scale, mod_stride, rpb)
#define LN_BOTH(CC, V)                              \
  if (C == CC && !act) {                            \
    if (x_is_f16) LN_CASE(__half, CC, V, false);    \
    else LN_CASE(float, CC, V, false);              \
    RET();                                          \
  }
#define LN_ACT(CC, V)                               \
  if (C == CC && act == 1) {                        \
    if (x_is_f16) LN_CASE(__half, CC, V, true);     \
    else LN_CASE(float, CC, V, true);               \
    RET();                                          \
  }
  LN_BOTH(512, true)
  LN_BOTH(768, true)
  LN_BOTH(64, false)
  LN_BOTH(96, false)
  LN_BOTH(192, false)
  LN_BOTH(128, true)
  LN_BOTH(384, true)
  LN_BOTH(256, true)
  LN_BOTH(1024, true)
  // the ResBlock widths of the structured-latent flow model (io 64 / 128, their skip concatenations, model 1024 / 2048)
  LN_ACT(32, false)
  LN_ACT(64, false)
  LN_ACT(128, true)
  LN_ACT(512, true)
  LN_ACT(256, true)
  LN_ACT(1024, true)
  LN_ACT(2048, true)
#undef LN_ACT
#undef LN_BOTH
#undef LN_CASE
  return GVF_ERR_UNSUPPORTED;
}

GVF_API void gvf_ln_set_two_rows(int on) { g_ln_two_rows = on ? 1 : 0; }

GVF_API int gvf_ln_mod_f16(const void* x, int x_is_f16, void* out, int M, int C, float eps,
                           const float* w, const float* b, const void* shift, const void* scale,
                           int mod_stride, int rows_per_batch, void* stream) {
  return ln_mod_launch(x, x_is_f16, out, M, C, eps, w, b, shift, scale, mod_stride, rows_per_batch, 0, stream);
}

GVF_API int gvf_ln_mod_act_f16(const void* x, int x_is_f16, void* out, int M, int C, float eps,
                               const float* w, const float* b, const void* shift, const void* scale,
                               int mod_stride, int rows_per_batch, int act, void* stream) {
  if (act != 0 && act != 1) return GVF_ERR_INVALID;
  return ln_mod_launch(x, x_is_f16, out, M, C, eps, w, b, shift, scale, mod_stride, rows_per_batch, act, stream);
}

GVF_API int gvf_rmsnorm_heads_f16(void* buf, long long rows, int ld, int H, int D, int k_off,
                                  const float* gamma_q, const float* gamma_k, void* stream) {
  if (!buf || !gamma_q || !gamma_k || rows <= 0 || (ld % 8) || (k_off % 8)) return GVF_ERR_INVALID;
  const long long n = rows * H * 2;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (D == 32) rmsnorm_heads_kernel<32><<<blocks, 256, 0, ST(stream)>>>((__half*)buf, rows, ld, H, k_off, gamma_q, gamma_k);
  else if (D == 64 && !(((uintptr_t)gamma_q | (uintptr_t)gamma_k) & 15))
    rmsnorm_heads64_kernel<<<(unsigned)((n * 8 + 255) / 256), 256, 0, ST(stream)>>>((__half*)buf, rows, ld, H, k_off, gamma_q, gamma_k);
  else if (D == 64) rmsnorm_heads_kernel<64><<<blocks, 256, 0, ST(stream)>>>((__half*)buf, rows, ld, H, k_off, gamma_q, gamma_k);
  else return GVF_ERR_UNSUPPORTED;
  RET();
}

GVF_API int gvf_dit_modulation(const float* t, int B, int C, int F, const void* W0, const float* b0,
                               const void* W2, const float* b2, const void* Wmod, const float* bmod,
                               int R, void* temb, void* silu_temb, void* mod_out, void* stream) {
  if (!t || !W0 || !W2 || !Wmod || !mod_out || B <= 0 || B > 8 || C > 1024 || (C % 32)) return GVF_ERR_INVALID;
  temb_kernel<<<B, 1024, (F + C) * sizeof(float), ST(stream)>>>(t, (const __half*)W0, b0, (const __half*)W2, b2, C,
                                                             F, (__half*)temb, (__half*)silu_temb);
  if (cudaGetLastError() != cudaSuccess) return GVF_ERR_CUDA;
  mod_gemv_kernel<8><<<(R + 7) / 8, 256, 0, ST(stream)>>>((const __half*)Wmod, bmod, (const __half*)silu_temb, B,
                                                          R, C, (__half*)mod_out);
  RET();
}

GVF_API int gvf_small_linear(const float* x, int ldx, const void* W, const float* bias, int M, int N,
                             int K, const float* add, int add_rows, void* out, int out_is_f16,
                             void* stream) {
  if (!x || !W || !out || K > 32 || K <= 0 || M <= 0 || N <= 0) return GVF_ERR_INVALID;
  const int ar = add_rows > 0 ? add_rows : 1;
  if (M >= 4096) {
    const dim3 grid((N + 255) / 256, (M + 63) / 64);
    if (out_is_f16)
      small_linear_kernel<__half, 64><<<grid, 256, 0, ST(stream)>>>(x, ldx, (const __half*)W, bias, M, N, K, add, ar, (__half*)out);
    else
      small_linear_kernel<float, 64><<<grid, 256, 0, ST(stream)>>>(x, ldx, (const __half*)W, bias, M, N, K, add, ar, (float*)out);
    RET();
  }
  const dim3 grid((N + 255) / 256, (M + 7) / 8);
  if (out_is_f16)
    small_linear_kernel<__half, 8><<<grid, 256, 0, ST(stream)>>>(x, ldx, (const __half*)W, bias, M, N, K, add, ar, (__half*)out);
  else
    small_linear_kernel<float, 8><<<grid, 256, 0, ST(stream)>>>(x, ldx, (const __half*)W, bias, M, N, K, add, ar, (float*)out);
  RET();
}

GVF_API int gvf_ape(const float* xyz, int R, int C, float* out, void* stream) {
  if (!xyz || !out || R <= 0 || C < 6) return GVF_ERR_INVALID;
  const long long n = (long long)R * C;
  ape_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>(xyz, R, C, out);
  RET();
}

GVF_API int gvf_vae_query_embed(const float* queries, int ldq, const void* gs, int Q, int C, void* out,
                                void* stream) {
  if (!queries || !gs || !out || Q <= 0) return GVF_ERR_INVALID;
  const dim3 grid((Q + 7) / 8);
  if (C == 768) query_embed_kernel<768><<<grid, 256, 0, ST(stream)>>>(queries, ldq, (const __half*)gs, Q, (__half*)out);
  else if (C == 96) query_embed_kernel<96><<<grid, 256, 0, ST(stream)>>>(queries, ldq, (const __half*)gs, Q, (__half*)out);
  else if (C == 384) query_embed_kernel<384><<<grid, 256, 0, ST(stream)>>>(queries, ldq, (const __half*)gs, Q, (__half*)out);
  else if (C == 192) query_embed_kernel<192><<<grid, 256, 0, ST(stream)>>>(queries, ldq, (const __half*)gs, Q, (__half*)out);
  else return GVF_ERR_UNSUPPORTED;
  RET();
}

GVF_API int gvf_vae_embed_sum(const float* xyz, int ldq, const int* xyz_row, const void* lin, int R, int C, void* out,
                              void* stream) {
  if (!xyz || !lin || !out || R <= 0) return GVF_ERR_INVALID;
  const dim3 grid((R + 7) / 8);
  if (C == 768) query_embed_kernel<768, false><<<grid, 256, 0, ST(stream)>>>(xyz, ldq, (const __half*)lin, R, (__half*)out, xyz_row);
  else if (C == 96) query_embed_kernel<96, false><<<grid, 256, 0, ST(stream)>>>(xyz, ldq, (const __half*)lin, R, (__half*)out, xyz_row);
  else if (C == 384) query_embed_kernel<384, false><<<grid, 256, 0, ST(stream)>>>(xyz, ldq, (const __half*)lin, R, (__half*)out, xyz_row);
  else if (C == 192) query_embed_kernel<192, false><<<grid, 256, 0, ST(stream)>>>(xyz, ldq, (const __half*)lin, R, (__half*)out, xyz_row);
  else return GVF_ERR_UNSUPPORTED;
  RET();
}

// DiagonalGaussianDistribution (model/autoencoder.py:304-326): logvar clamped to [-30, 20]; sample = mean + exp(logvar / 2)
// * noise; kl[b] = 0.5 * mean over the entry's elements of (mean^2 + var - 1 - logvar).  One CTA per batch entry.
__global__ void __launch_bounds__(256) diag_gaussian_kernel(const float* __restrict__ mean, const float* __restrict__ logvar,
                                                            const float* __restrict__ noise, long long per_batch,
                                                            float* __restrict__ sample, float* __restrict__ kl) {
  __shared__ float red[8];
  const long long base = (long long)blockIdx.x * per_batch;
  float acc = 0.f;
  for (long long i = threadIdx.x; i < per_batch; i += 256) {
    const float m = mean[base + i];
    const float lv = fminf(fmaxf(logvar[base + i], -30.0f), 20.0f);
    if (sample) sample[base + i] = m + expf(0.5f * lv) * (noise ? noise[base + i] : 0.f);
    acc += m * m + expf(lv) - 1.0f - lv;
  }
  acc = gvf::warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w];
    if (kl) kl[blockIdx.x] = 0.5f * s / (float)per_batch;
  }
}

// backward of the above: d mean = d sample + d kl[b] * mean / n,  d logvar = (d sample * noise * std / 2 + d kl[b] * (var - 1)
// / (2 n)) inside the clamp range, 0 outside (n = elements per batch entry)
__global__ void __launch_bounds__(256) diag_gaussian_bwd_kernel(const float* __restrict__ mean, const float* __restrict__ logvar,
                                                                const float* __restrict__ noise,
                                                                const float* __restrict__ dsample,
                                                                const float* __restrict__ dkl, long long per_batch,
                                                                long long total, float* __restrict__ dmean,
                                                                float* __restrict__ dlogvar) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float gk = dkl ? dkl[i / per_batch] / (float)per_batch : 0.f;
  const float gs = dsample ? dsample[i] : 0.f;
  const float lv0 = logvar[i];
  const bool inside = lv0 >= -30.0f && lv0 <= 20.0f;
  const float lv = fminf(fmaxf(lv0, -30.0f), 20.0f);
  dmean[i] = gs + gk * mean[i];
  dlogvar[i] = inside ? gs * (noise ? noise[i] : 0.f) * 0.5f * expf(0.5f * lv) + 0.5f * gk * (expf(lv) - 1.0f) : 0.f;
}

GVF_API int gvf_diag_gaussian_bwd(const float* mean, const float* logvar, const float* noise, const float* dsample,
                                  const float* dkl, int B, long long per_batch, float* dmean, float* dlogvar, void* stream) {
  if (!mean || !logvar || !dmean || !dlogvar || B <= 0 || per_batch <= 0) return GVF_ERR_INVALID;
  const long long total = (long long)B * per_batch;
  diag_gaussian_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ST(stream)>>>(mean, logvar, noise, dsample, dkl, per_batch,
                                                                                  total, dmean, dlogvar);
  RET();
}

GVF_API int gvf_diag_gaussian(const float* mean, const float* logvar, const float* noise, int B, long long per_batch,
                              float* sample, float* kl, void* stream) {
  if (!mean || !logvar || B <= 0 || per_batch <= 0) return GVF_ERR_INVALID;
  diag_gaussian_kernel<<<B, 256, 0, ST(stream)>>>(mean, logvar, noise, per_batch, sample, kl);
  RET();
}

GVF_API int gvf_geglu_f16(const void* h, long long M, int F, void* out, void* stream) {
  if (!h || !out || M <= 0 || (F % 8)) return GVF_ERR_INVALID;
  const long long n8 = M * (F / 8);
  geglu_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, ST(stream)>>>((const __half*)h, M, F, (__half*)out);
  RET();
}

GVF_API int gvf_cast_f32_f16(const float* x, long long n, void* out, void* stream) {
  if (!x || !out || n <= 0) return GVF_ERR_INVALID;
  const long long th = (n + 7) / 8;
  cast_f16_kernel<<<(unsigned)((th + 255) / 256), 256, 0, ST(stream)>>>(x, n, (__half*)out);
  RET();
}

GVF_API int gvf_dit_final_layer(const float* x, int M, int C, int O, const void* shift, const void* scale,
                                int mod_stride, int rows_per_batch, const void* W, const float* bias,
                                float* out, void* stream) {
  if (!x || !shift || !scale || !W || !bias || !out || M <= 0 || rows_per_batch <= 0) return GVF_ERR_INVALID;
  const dim3 grid((M + 31) / 32);
  if (C == 512 && O == 16)
    final_layer_kernel<512, 16><<<grid, 256, 0, ST(stream)>>>(x, M, (const __half*)shift, (const __half*)scale,
                                                              mod_stride, rows_per_batch, (const __half*)W, bias, out);
  else if (C == 128 && O == 16)
    final_layer_kernel<128, 16><<<grid, 256, 0, ST(stream)>>>(x, M, (const __half*)shift, (const __half*)scale,
                                                              mod_stride, rows_per_batch, (const __half*)W, bias, out);
  else if (C == 64 && O == 16)
    final_layer_kernel<64, 16><<<grid, 256, 0, ST(stream)>>>(x, M, (const __half*)shift, (const __half*)scale,
                                                             mod_stride, rows_per_batch, (const __half*)W, bias, out);
  else return GVF_ERR_UNSUPPORTED;
  RET();
}

GVF_API int gvf_dpm_x0(const float* x, const float* v, long long n, int branches, int model_type,
                       float alpha, float sigma, float s1, float s2, float* x0, void* stream) {
  if (!x || !v || !x0 || n <= 0 || (branches != 1 && branches != 3)) return GVF_ERR_INVALID;
  if (model_type != 0 && model_type != 1) return GVF_ERR_UNSUPPORTED;
  dpm_x0_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>(x, v, n, branches, model_type, alpha, sigma,
                                                                     s1, s2, x0);
  RET();
}

GVF_API int gvf_flow_euler_step(const float* x, const float* v, const float* v_neg, long long n, double cfg_strength,
                                double t, double t_prev, double sigma_min, float* x_prev, float* x0, void* stream) {
  if (!x || !v || !x_prev || n <= 0) return GVF_ERR_INVALID;
  flow_euler_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>(
      x, v, v_neg, n, (float)(1.0 + cfg_strength), (float)cfg_strength, (float)(t - t_prev), (float)(1.0 - sigma_min),
      (float)(sigma_min + (1.0 - sigma_min) * t), x_prev, x0);
  RET();
}

GVF_API int gvf_dpm_error_sq(const float* x_higher, const float* x_lower, const float* x_prev, int B,
                             long long n_per_batch, float atol, float rtol, float* E2, void* stream) {
  if (!x_higher || !x_lower || !x_prev || !E2 || B <= 0 || n_per_batch <= 0) return GVF_ERR_INVALID;
  const unsigned bx = (unsigned)((n_per_batch + 256 * 8 - 1) / (256 * 8));
  dpm_error_sq_kernel<<<dim3(bx > 1024 ? 1024 : bx, B), 256, 0, ST(stream)>>>(x_higher, x_lower, x_prev, n_per_batch,
                                                                               atol, rtol, E2);
  RET();
}

GVF_API int gvf_dpm_update(const float* x, const float* m0, const float* m1, long long n, float cx, float cm,
                           float inv_r0, int order, float* out, void* stream) {
  if (!x || !m0 || !out || n <= 0 || (order == 2 && !m1) || order < 1 || order > 2) return GVF_ERR_INVALID;
  dpm_update_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>(x, m0, m1, n, cx, cm, inv_r0, order, out);
  RET();
}

GVF_API int gvf_affine_lastdim(const float* x, long long n, int C, const float* a, const float* b, float as,
                               float bs, float* out, void* stream) {
  if (!x || !out || n <= 0 || C <= 0) return GVF_ERR_INVALID;
  affine_lastdim_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>(x, n, C, a, b, as, bs, out);
  RET();
}

}  // extern "C"
