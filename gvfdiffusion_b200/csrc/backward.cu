// backward.cu -- the non-GEMM backward kernels of the motion-VAE decoder's training step (BASELINE configs[2] / [4]:
// reference train_vae.py:293-353 back-propagates the render / interpolation losses through
// `GSKLTemporalVariationalAutoEncoder.decode`, model/autoencoder.py:552-609).  What autograd derives there for
// LayerNorm (PreNorm, :73-88), GEGLU (:90-93), the K <= 32 Linears (proj :585, gs_embedding :389, to_outputs :574), the
// query embedding (:250-301,389-391,560) and the bias gradients is written here as one HBM pass each; the dense
// dgrad / wgrad contractions run on the tcgen05 GEMM (csrc/gemm.cu) over transposed operands (transpose_kernel below),
// attention in csrc/attn_bwd.cu.  Activation gradients are fp16 like the activations they belong to (the reference's
// autocast regime); weight / bias gradients are produced in fp32.  All reductions are two-stage and deterministic.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gvf_b200.h"

namespace gvf {

__device__ __forceinline__ float bw_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float bw_r16(float x) { return __half2float(__float2half_rn(x)); }

// ------------------------------------------------------------------------------------------------ transpose
// in [R, C] (row stride ld_in) -> out [C, Rp] (row stride ld_out >= Rp), Rp = R rounded up to the caller's padding:
// columns r in [R, ld_out) of every output row are zero-filled (the GEMM reduces over them).  64 x 64 tiles.
__global__ void __launch_bounds__(256) transpose_f16_kernel(const __half* __restrict__ in, int R, int C, long long ld_in,
                                                            __half* __restrict__ out, long long ld_out) {
  __shared__ __half tile[64][66];
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
  const bool vec_in = ((ld_in & 1) == 0) && ((reinterpret_cast<uintptr_t>(in) & 3) == 0);
#pragma unroll
  for (int rr = ty; rr < 64; rr += 8) {
    const int r = r0 + rr, c = c0 + 2 * tx;
    __half2 v = __floats2half2_rn(0.f, 0.f);
    if (r < R) {
      if (vec_in && c + 1 < C) v = *reinterpret_cast<const __half2*>(in + (long long)r * ld_in + c);
      else {
        if (c < C) v.x = in[(long long)r * ld_in + c];
        if (c + 1 < C) v.y = in[(long long)r * ld_in + c + 1];
      }
    }
    tile[rr][2 * tx] = v.x;
    tile[rr][2 * tx + 1] = v.y;
  }
  __syncthreads();
  const bool vec_out = ((ld_out & 1) == 0) && ((reinterpret_cast<uintptr_t>(out) & 3) == 0);
#pragma unroll
  for (int cc = ty; cc < 64; cc += 8) {
    const int c = c0 + cc, r = r0 + 2 * tx;
    if (c >= C) continue;
    __half2 v;
    v.x = tile[2 * tx][cc];
    v.y = tile[2 * tx + 1][cc];
    if (vec_out && r + 1 < ld_out) *reinterpret_cast<__half2*>(out + (long long)c * ld_out + r) = v;
    else {
      if (r < ld_out) out[(long long)c * ld_out + r] = v.x;
      if (r + 1 < ld_out) out[(long long)c * ld_out + r + 1] = v.y;
    }
  }
}

// ------------------------------------------------------------------------------------------------ column sums
// out[n] = sum_m x[m, n] (bias gradients).  Stage 1: CTA = 256 columns x a slab of rows -> partial [slabs, N].
template <typename T>
__global__ void __launch_bounds__(256) colsum_partial_kernel(const T* __restrict__ x, long long M, int N, long long ld,
                                                             int rows_per_slab, float* __restrict__ partial) {
  const int n = blockIdx.x * 256 + threadIdx.x;
  if (n >= N) return;
  const long long m0 = (long long)blockIdx.y * rows_per_slab;
  const long long m1 = m0 + rows_per_slab < M ? m0 + rows_per_slab : M;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  long long m = m0;
  for (; m + 3 < m1; m += 4) {
    s0 += (float)x[m * ld + n];
    s1 += (float)x[(m + 1) * ld + n];
    s2 += (float)x[(m + 2) * ld + n];
    s3 += (float)x[(m + 3) * ld + n];
  }
  for (; m < m1; ++m) s0 += (float)x[m * ld + n];
  partial[(long long)blockIdx.y * N + n] = (s0 + s1) + (s2 + s3);
}
__global__ void __launch_bounds__(256) reduce_slabs_kernel(const float* __restrict__ partial, int slabs, long long n,
                                                           float* __restrict__ out, int accumulate) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int k = 0; k < slabs; ++k) s += partial[(long long)k * n + i];
  out[i] = accumulate ? out[i] + s : s;
}

// fp16, N % 8 == 0: one launch.  CTA = 256 columns (32 x 16 B vectors) x 8 row lanes over a slab of rows; the slab
// partials go to the workspace and the LAST CTA of a column block to finish (ticket counter, self-resetting) adds them
// in slab order -- the same deterministic sum as the two-launch form without its second launch (~7 us each, 150 bias
// gradients per static-VAE step).  Not re-entrant across streams (neither is the shared workspace).
__device__ unsigned int g_colsum_tickets[256];
__global__ void __launch_bounds__(256) colsum_f16_fused_kernel(const __half* __restrict__ x, long long M, int N,
                                                               long long ld, int rows_per_slab,
                                                               float* __restrict__ partial, float* __restrict__ out,
                                                               int accumulate) {
  __shared__ float red[8][256 + 8];
  __shared__ unsigned int s_ticket;
  const int vc = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int n0 = blockIdx.x * 256 + vc * 8;
  const long long m0 = (long long)blockIdx.y * rows_per_slab;
  const long long m1 = m0 + rows_per_slab < M ? m0 + rows_per_slab : M;
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f;
  if (n0 < N) {
#pragma unroll 4
    for (long long m = m0 + rl; m < m1; m += 8) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + m * ld + n0));
      const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[2 * j] += __low2float(h[j]);
        s[2 * j + 1] += __high2float(h[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[rl][vc * 8 + j] = s[j];
  __syncthreads();
  const int n = blockIdx.x * 256 + threadIdx.x;
  if (n < N) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) t += red[r][threadIdx.x];
    partial[(long long)blockIdx.y * N + n] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_ticket = atomicAdd(&g_colsum_tickets[blockIdx.x], 1u);
  __syncthreads();
  if (s_ticket != gridDim.y - 1) return;
  __threadfence();
  if (n < N) {
    float t = 0.f;
    for (unsigned k = 0; k < gridDim.y; ++k) t += __ldcg(partial + (long long)k * N + n);
    out[n] = accumulate ? out[n] + t : t;
  }
  if (threadIdx.x == 0) g_colsum_tickets[blockIdx.x] = 0u;
}

// ------------------------------------------------------------------------------------------------ LayerNorm backward
// y = (x - mean) * rstd (no affine, PreNorm):  dx = rstd * (dy - mean(dy) - yhat * mean(dy * yhat)) (+ dres).
// x fp16 or fp32 [M, C]; dy, dres, dx fp16.  One warp per row, row in registers.
template <typename TIn, int C>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const TIn* __restrict__ x, const __half* __restrict__ dy,
                                                     const __half* __restrict__ dres, __half* __restrict__ dx, int M,
                                                     float eps) {
  constexpr int PER = C / 32;
  const int lane = threadIdx.x & 31;
  for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); row < M; row += gridDim.x * 8) {
    float v[PER], g[PER];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      v[i] = (float)x[(size_t)row * C + i * 32 + lane];
      g[i] = __half2float(dy[(size_t)row * C + i * 32 + lane]);
      s += v[i];
    }
    const float mean = bw_warp_sum(s) / C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) { v[i] -= mean; q += v[i] * v[i]; }
    const float rstd = rsqrtf(bw_warp_sum(q) / C + eps);
    float sg = 0.f, sgy = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) { v[i] *= rstd; sg += g[i]; sgy += g[i] * v[i]; }
    const float mg = bw_warp_sum(sg) / C, mgy = bw_warp_sum(sgy) / C;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      float r = rstd * (g[i] - mg - v[i] * mgy);
      if (dres) r += __half2float(dres[(size_t)row * C + i * 32 + lane]);
      dx[(size_t)row * C + i * 32 + lane] = __float2half_rn(r);
    }
  }
}

// ------------------------------------------------------------------------------------------------ GEGLU backward
// h [M, 2F] = [a | g] fp16, dG [M, F] fp16 -> dh [M, 2F]:  da = dG * gelu(g),  dg = dG * a * gelu'(g),
// gelu'(x) = Phi(x) + x phi(x)  (exact erf GELU, model/autoencoder.py:90-93).
__global__ void __launch_bounds__(256) geglu_bwd_kernel(const __half* __restrict__ h, const __half* __restrict__ dG,
                                                        long long M, int F, __half* __restrict__ dh) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n8 = M * (F / 8);
  if (gid >= n8) return;
  const long long row = gid / (F / 8);
  const int c = (int)(gid - row * (F / 8)) * 8;
  const uint4 a = *reinterpret_cast<const uint4*>(h + row * 2 * F + c);
  const uint4 g = *reinterpret_cast<const uint4*>(h + row * 2 * F + F + c);
  const uint4 d = *reinterpret_cast<const uint4*>(dG + row * F + c);
  const __half* ah = reinterpret_cast<const __half*>(&a);
  const __half* gh = reinterpret_cast<const __half*>(&g);
  const __half* dh8 = reinterpret_cast<const __half*>(&d);
  __align__(16) __half oa[8], og[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float x = __half2float(gh[j]), av = __half2float(ah[j]), dv = __half2float(dh8[j]);
    const float cdf = 0.5f * (1.0f + erff(x * 0.7071067811865476f));
    const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
    oa[j] = __float2half_rn(dv * bw_r16(x * cdf));
    og[j] = __float2half_rn(dv * av * (cdf + x * pdf));
  }
  *reinterpret_cast<uint4*>(dh + row * 2 * F + c) = *reinterpret_cast<uint4*>(oa);
  *reinterpret_cast<uint4*>(dh + row * 2 * F + F + c) = *reinterpret_cast<uint4*>(og);
}

// ------------------------------------------------------------------------------------------------ get_gaussian_tensor
// Backward of gvf_gaussian_tensor (reference train_vae.py:466-472 under autograd: the activated canonical Gaussians are
// the motion VAE's decoder queries, so the deformation losses reach the static VAE through them).  g [P, 14] =
// [xyz3 | rgb3 | opacity1 | scale3 | rot4] -> gradients of the raw tensors.
__global__ void __launch_bounds__(256) gaussian_tensor_bwd_kernel(const gvf_raster_params prm, int P,
                                                                  const float* __restrict__ scaling,
                                                                  const float* __restrict__ rotation,
                                                                  const float* __restrict__ opacity,
                                                                  const float* __restrict__ g, float* __restrict__ d_xyz,
                                                                  float* __restrict__ d_dc, float* __restrict__ d_scaling,
                                                                  float* __restrict__ d_rotation,
                                                                  float* __restrict__ d_opacity) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const float* gi = g + (size_t)i * 14;
  const float k2 = prm.min_kernel * prm.min_kernel;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    d_xyz[(size_t)i * 3 + c] = gi[c] * prm.aabb[3 + c];
    d_dc[(size_t)i * 3 + c] = gi[3 + c];
    const float r = scaling[(size_t)i * 3 + c] + prm.scale_bias;
    // s = softplus(r) | exp(r);  ds/dr = sigmoid(r) | s;  out = sqrt(s^2 + k2): d out / ds = s / out
    const float s = prm.softplus ? (r > 20.f ? r : log1pf(expf(r))) : expf(r);
    const float ds = prm.softplus ? 1.0f / (1.0f + expf(-r)) : s;
    const float o = sqrtf(s * s + k2);
    d_scaling[(size_t)i * 3 + c] = gi[7 + c] * (o > 0.f ? s / o : 0.f) * ds;
  }
  const float op = 1.0f / (1.0f + expf(-(opacity[i] + prm.opacity_bias)));
  d_opacity[i] = gi[6] * op * (1.0f - op);
  float q[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) q[c] = rotation[(size_t)i * 4 + c] + (c == 0 ? 1.0f : 0.0f);
  const float n = fmaxf(sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]), 1e-12f);
  float dot = 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c) dot += gi[10 + c] * q[c];
  // y = q / n:  dq = (g - y (g . y)) / n
#pragma unroll
  for (int c = 0; c < 4; ++c) d_rotation[(size_t)i * 4 + c] = (gi[10 + c] - q[c] * dot / (n * n)) / n;
}

// ------------------------------------------------------------------------------------------------ GELU (tanh)
// The sparse trunk's MLP (sparse_transformer.py: nn.GELU(approximate="tanh")).  Inference fuses it into the fc1 epilogue;
// the training forward keeps the pre-activation and applies the same tanh.approx form here.
__device__ __forceinline__ float bw_tanh(float u) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  return t;
}
__global__ void __launch_bounds__(256) gelu_tanh_kernel(const __half* __restrict__ h, long long n8, __half* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const uint4 v = reinterpret_cast<const uint4*>(h)[i];
  const __half* hv = reinterpret_cast<const __half*>(&v);
  __align__(16) __half o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float x = __half2float(hv[j]);
    const float u = x * fmaf(0.7978845608028654f * 0.044715f, x * x, 0.7978845608028654f);
    const float hx = 0.5f * x;
    o[j] = __float2half_rn(fmaf(hx, bw_tanh(u), hx));
  }
  reinterpret_cast<uint4*>(out)[i] = *reinterpret_cast<uint4*>(o);
}
// d/dx [0.5 x (1 + tanh u)], u = k0 (x + k1 x^3):  0.5 (1 + t) + 0.5 x (1 - t^2) k0 (1 + 3 k1 x^2)
__global__ void __launch_bounds__(256) gelu_tanh_bwd_kernel(const __half* __restrict__ h, const __half* __restrict__ dy,
                                                            long long n8, __half* __restrict__ dh) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const uint4 v = reinterpret_cast<const uint4*>(h)[i], g = reinterpret_cast<const uint4*>(dy)[i];
  const __half* hv = reinterpret_cast<const __half*>(&v);
  const __half* gv = reinterpret_cast<const __half*>(&g);
  __align__(16) __half o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float x = __half2float(hv[j]);
    const float k0 = 0.7978845608028654f, k1 = 0.044715f;
    const float t = bw_tanh(k0 * (x + k1 * x * x * x));
    const float d = 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * k0 * (1.0f + 3.0f * k1 * x * x);
    o[j] = __float2half_rn(__half2float(gv[j]) * d);
  }
  reinterpret_cast<uint4*>(dh)[i] = *reinterpret_cast<uint4*>(o);
}

// ------------------------------------------------------------------------------------------------ K <= 32 Linears
// dx[m, k] = sum_n dy[m, n] W[n, k]   (y = x W^T + b with K <= 32 inputs).  Warp per row; W fp16 [N, K].
template <typename TDy>
__global__ void __launch_bounds__(256) small_linear_bwd_input_kernel(const TDy* __restrict__ dy, long long ld,
                                                                     const __half* __restrict__ W, long long M, int N,
                                                                     int K, float* __restrict__ dx, int ldx) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  float acc[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) acc[k] = 0.f;
  for (int n = lane; n < N; n += 32) {
    const float g = (float)dy[row * ld + n];
    const __half* w = W + (size_t)n * K;
#pragma unroll
    for (int k = 0; k < 32; ++k)
      if (k < K) acc[k] = fmaf(g, __half2float(w[k]), acc[k]);
  }
#pragma unroll
  for (int k = 0; k < 32; ++k)
    if (k < K) {
      const float s = bw_warp_sum(acc[k]);
      if (lane == 0) dx[row * ldx + k] = s;
    }
}

// partial[slab][k][n] = sum_{m in slab} x[m, k] * y[m, n]   (k < K <= 16; thread = two adjacent columns, x rows broadcast
// from shared memory as float4).  Serves dW^T of proj / gs_embedding (x = the layer input, y = dy fp16) and dW of
// to_outputs (x = dOut fp32 [M, 14], y = the layer input fp16).
template <typename TY>
__global__ void __launch_bounds__(256) skinny_outer_partial_kernel(const float* __restrict__ x, int ldx, int K,
                                                                   const TY* __restrict__ y, long long ldy, long long M,
                                                                   int N, int rows_per_slab, float* __restrict__ partial) {
  __shared__ __align__(16) float sx[64][16];
  const int n = (blockIdx.x * 256 + threadIdx.x) * 2;
  const long long m0 = (long long)blockIdx.y * rows_per_slab;
  const long long m1 = m0 + rows_per_slab < M ? m0 + rows_per_slab : M;
  const bool vec = (sizeof(TY) == 2) && ((ldy & 1) == 0) && ((reinterpret_cast<uintptr_t>(y) & 3) == 0) && n + 1 < N;
  float a0[16], a1[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) { a0[k] = 0.f; a1[k] = 0.f; }
  for (long long mb = m0; mb < m1; mb += 64) {
    __syncthreads();
    for (int t = threadIdx.x; t < 64 * 16; t += 256) {
      const int r = t >> 4, k = t & 15;
      sx[r][k] = (mb + r < m1 && k < K) ? x[(mb + r) * ldx + k] : 0.f;
    }
    __syncthreads();
    if (n < N) {
      const int lim = (int)(m1 - mb < 64 ? m1 - mb : 64);
#pragma unroll 2
      for (int r = 0; r < lim; ++r) {
        float y0, y1 = 0.f;
        if constexpr (sizeof(TY) == 2) {
          if (vec) {
            const float2 v = __half22float2(*reinterpret_cast<const __half2*>(y + (mb + r) * ldy + n));
            y0 = v.x; y1 = v.y;
          } else {
            y0 = __half2float(y[(mb + r) * ldy + n]);
            if (n + 1 < N) y1 = __half2float(y[(mb + r) * ldy + n + 1]);
          }
        } else {
          y0 = y[(mb + r) * ldy + n];
          if (n + 1 < N) y1 = y[(mb + r) * ldy + n + 1];
        }
        const float4* xr = reinterpret_cast<const float4*>(sx[r]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 xv = xr[q];
          a0[4 * q] = fmaf(xv.x, y0, a0[4 * q]);         a1[4 * q] = fmaf(xv.x, y1, a1[4 * q]);
          a0[4 * q + 1] = fmaf(xv.y, y0, a0[4 * q + 1]); a1[4 * q + 1] = fmaf(xv.y, y1, a1[4 * q + 1]);
          a0[4 * q + 2] = fmaf(xv.z, y0, a0[4 * q + 2]); a1[4 * q + 2] = fmaf(xv.z, y1, a1[4 * q + 2]);
          a0[4 * q + 3] = fmaf(xv.w, y0, a0[4 * q + 3]); a1[4 * q + 3] = fmaf(xv.w, y1, a1[4 * q + 3]);
        }
      }
    }
  }
  if (n < N) {
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if (k < K) {
        partial[((long long)blockIdx.y * K + k) * N + n] = a0[k];
        if (n + 1 < N) partial[((long long)blockIdx.y * K + k) * N + n + 1] = a1[k];
      }
  }
}

// The same for fp16 y with N % 8 == 0 (the [B*T*Q, 768] activations of the decoder's output side, 0.6 GB per pass):
// thread = 8 adjacent columns (one 16 B load per row) x one of RG row groups, 8 x K accumulators in registers, the row
// groups of a CTA are summed through shared memory.  FMA-bound at K = 14 (4.2 GFMA for the 393 216 x 768 pass).
template <int RG>
__global__ void __launch_bounds__(384) skinny_outer_vec8_kernel(const float* __restrict__ x, int ldx, int K,
                                                                const __half* __restrict__ y, long long ldy, long long M,
                                                                int N, int rows_per_slab, float* __restrict__ partial) {
  extern __shared__ float sred[];                       // [RG][16][8 * CT] for the final reduction, also x staging
  __shared__ __align__(16) float sx[64][16];
  const int CT = N >> 3;                                // column threads per row
  const int c = threadIdx.x % CT, rg = threadIdx.x / CT;
  const bool active = rg < RG;
  const long long m0 = (long long)blockIdx.x * rows_per_slab;
  const long long m1 = m0 + rows_per_slab < M ? m0 + rows_per_slab : M;
  float acc[16][8];
#pragma unroll
  for (int k = 0; k < 16; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[k][j] = 0.f;
  for (long long mb = m0; mb < m1; mb += 64) {
    __syncthreads();
    for (int t = threadIdx.x; t < 64 * 16; t += blockDim.x) {
      const int r = t >> 4, k = t & 15;
      sx[r][k] = (mb + r < m1 && k < K) ? x[(mb + r) * ldx + k] : 0.f;
    }
    __syncthreads();
    if (active) {
      const int lim = (int)(m1 - mb < 64 ? m1 - mb : 64);
#pragma unroll 2
      for (int r = rg; r < lim; r += RG) {
        const uint4 v = *reinterpret_cast<const uint4*>(y + (mb + r) * ldy + 8 * c);
        const __half2* h2 = reinterpret_cast<const __half2*>(&v);
        float yv[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(h2[j]); yv[2 * j] = f.x; yv[2 * j + 1] = f.y; }
        const float4* xr = reinterpret_cast<const float4*>(sx[r]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 xv = xr[q];
          const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
          for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[4 * q + e][j] = fmaf(xs[e], yv[j], acc[4 * q + e][j]);
        }
      }
    }
  }
  // sum the row groups: [rg][k][n] in shared memory, then thread (c, rg = 0) adds them up
  __syncthreads();
  if (active) {
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if (k < K)
#pragma unroll
        for (int j = 0; j < 8; ++j) sred[((size_t)rg * K + k) * N + 8 * c + j] = acc[k][j];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < K * N; t += blockDim.x) {
    float v = 0.f;
    for (int g = 0; g < RG; ++g) v += sred[(size_t)g * K * N + t];
    partial[(size_t)blockIdx.x * K * N + t] = v;
  }
}

// out[m, n] = sum_k x[m, k] Wt[k, n]   (K <= 16, fp32 operands; out fp16 or fp32): rank-K expansion, e.g. the gradient of
// to_outputs' input, d lat = d out [M, 14] @ W [14, 768].  Thread = two adjacent columns (weights in registers), CTA =
// 64 rows staged in shared memory; one coalesced 4 B (8 B) store per thread and row.
template <typename TOut>
__global__ void __launch_bounds__(512) skinny_expand_kernel(const float* __restrict__ x, int ldx, int K,
                                                            const float* __restrict__ Wt, long long M, int N,
                                                            TOut* __restrict__ out, long long ldo) {
  __shared__ float sx[64][16];
  const long long m0 = (long long)blockIdx.x * 64;
  for (int t = threadIdx.x; t < 64 * 16; t += blockDim.x) {
    const int r = t >> 4, k = t & 15;
    sx[r][k] = (m0 + r < M && k < K) ? x[(m0 + r) * ldx + k] : 0.f;
  }
  __syncthreads();
  const int lim = (int)(M - m0 < 64 ? M - m0 : 64);
  for (int n = 2 * threadIdx.x; n < N; n += 2 * blockDim.x) {
    float w0[16], w1[16];
    const bool two = n + 1 < N;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      w0[k] = k < K ? Wt[(size_t)k * N + n] : 0.f;
      w1[k] = (k < K && two) ? Wt[(size_t)k * N + n + 1] : 0.f;
    }
    for (int r = 0; r < lim; ++r) {
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) { a0 = fmaf(sx[r][k], w0[k], a0); a1 = fmaf(sx[r][k], w1[k], a1); }
      TOut* o = out + (m0 + r) * ldo + n;
      if constexpr (sizeof(TOut) == 2) {
        if (two && ((ldo & 1) == 0)) *reinterpret_cast<__half2*>(o) = __floats2half2_rn(a0, a1);
        else { o[0] = __float2half_rn(a0); if (two) o[1] = __float2half_rn(a1); }
      } else {
        o[0] = a0;
        if (two) o[1] = a1;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ query embedding
// Backward of gvf_vae_query_embed: out = LN_1e-6( LN_1e-5(gs) + LN_1e-5(PE(xyz)) ), PE as in query_embed_kernel
// (fp16-rounded argument and value).  d_out fp16 [Q, C] -> d_gs fp16 [Q, C], d_xyz fp32 [Q, 3] (the derivative of the
// sin / cos features with respect to the coordinate; the fp16 roundings are treated as identity, like autograd does).
template <int C>
__global__ void __launch_bounds__(256) query_embed_bwd_kernel(const float* __restrict__ queries, int ldq,
                                                              const __half* __restrict__ gs,
                                                              const __half* __restrict__ dout, int Q,
                                                              __half* __restrict__ dgs, float* __restrict__ dxyz,
                                                              int ld_dxyz, int accumulate) {
  constexpr int PER = C / 32, E = C / 6;
  const int qi = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (qi >= Q) return;
  float g[PER], p[PER], dpa[PER], u[PER], go[PER];
  float sg = 0.f, sp = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = i * 32 + lane;
    g[i] = __half2float(gs[(size_t)qi * C + c]);
    go[i] = __half2float(dout[(size_t)qi * C + c]);
    const int coord = c / (2 * E), w = c - coord * 2 * E;
    const int j = w < E ? w : w - E;
    const float om = bw_r16((float)(1.0 / pow(10000.0, (double)j / ((double)E / 2.0))));
    const float x16 = bw_r16(queries[(size_t)qi * ldq + coord]);
    const float a = bw_r16(x16 * om);
    float sn, cs;
    sincosf(a, &sn, &cs);
    p[i] = bw_r16(w < E ? sn : cs);
    dpa[i] = (w < E ? cs : -sn) * om;            // d p / d coordinate
    sg += g[i];
    sp += p[i];
  }
  const float mg = bw_warp_sum(sg) / C, mp = bw_warp_sum(sp) / C;
  float vg = 0.f, vp = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) { g[i] -= mg; p[i] -= mp; vg += g[i] * g[i]; vp += p[i] * p[i]; }
  const float rg = rsqrtf(bw_warp_sum(vg) / C + 1e-5f), rp = rsqrtf(bw_warp_sum(vp) / C + 1e-5f);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) { g[i] *= rg; p[i] *= rp; u[i] = g[i] + p[i]; s += u[i]; }     // g, p = normalised
  const float m = bw_warp_sum(s) / C;
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) { u[i] -= m; v += u[i] * u[i]; }
  const float rs = rsqrtf(bw_warp_sum(v) / C + 1e-6f);
  // outer LayerNorm
  float a0 = 0.f, a1 = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) { u[i] *= rs; a0 += go[i]; a1 += go[i] * u[i]; }
  const float m0 = bw_warp_sum(a0) / C, m1 = bw_warp_sum(a1) / C;
  float b0 = 0.f, b1 = 0.f, c0 = 0.f, c1 = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    go[i] = rs * (go[i] - m0 - u[i] * m1);        // d(g' + p')
    b0 += go[i]; b1 += go[i] * g[i];
    c0 += go[i]; c1 += go[i] * p[i];
  }
  const float gb0 = bw_warp_sum(b0) / C, gb1 = bw_warp_sum(b1) / C, pc1 = bw_warp_sum(c1) / C;
  (void)c0;
  float dx[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = i * 32 + lane;
    dgs[(size_t)qi * C + c] = __float2half_rn(rg * (go[i] - gb0 - g[i] * gb1));
    const float dpe = rp * (go[i] - gb0 - p[i] * pc1);
    const int coord = c / (2 * E);
    if (coord < 3) {
#pragma unroll
      for (int t = 0; t < 3; ++t)
        if (coord == t) dx[t] += dpe * dpa[i];
    }
  }
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    const float r = bw_warp_sum(dx[t]);
    if (lane == 0) {
      float* dst = dxyz + (size_t)qi * ld_dxyz + t;
      *dst = accumulate ? *dst + r : r;
    }
  }
}

}  // namespace gvf

using namespace gvf;
#define ST(s) ((cudaStream_t)(s))
#define RET() return cudaGetLastError() == cudaSuccess ? GVF_OK : GVF_ERR_CUDA

extern "C" {

GVF_API int gvf_transpose_f16(const void* in, int R, int C, long long ld_in, void* out, long long ld_out, void* stream) {
  if (!in || !out || R <= 0 || C <= 0 || ld_in < C || ld_out < R) return GVF_ERR_INVALID;
  const dim3 grid((C + 63) / 64, (int)((ld_out + 63) / 64));
  transpose_f16_kernel<<<grid, 256, 0, ST(stream)>>>((const __half*)in, R, C, ld_in, (__half*)out, ld_out);
  RET();
}

GVF_API size_t gvf_colsum_workspace_bytes(long long M, int N, int K) {
  const int slabs = 512;
  return (size_t)slabs * (size_t)(K > 0 ? K : 1) * (size_t)N * sizeof(float);
}

static int slab_rows(long long M, int* slabs) {
  long long rps = (M + 255) / 256;
  rps = ((rps + 63) / 64) * 64;                  // whole 64-row staging batches
  *slabs = (int)((M + rps - 1) / rps);
  return (int)rps;
}

GVF_API int gvf_colsum(const void* x, int x_is_f16, long long M, int N, long long ld, float* workspace,
                       size_t workspace_bytes, float* out, int accumulate, void* stream) {
  if (!x || !workspace || !out || M <= 0 || N <= 0 || ld < N) return GVF_ERR_INVALID;
  int slabs;
  const int rps = slab_rows(M, &slabs);
  if (workspace_bytes < (size_t)slabs * N * sizeof(float)) return GVF_ERR_WORKSPACE;
  const dim3 grid((N + 255) / 256, slabs);
  if (x_is_f16 && (N % 8) == 0 && (ld % 8) == 0 && ((uintptr_t)x & 15) == 0 && grid.x <= 256) {
    colsum_f16_fused_kernel<<<grid, 256, 0, ST(stream)>>>((const __half*)x, M, N, ld, rps, workspace, out, accumulate);
    RET();
  }
  if (x_is_f16) colsum_partial_kernel<__half><<<grid, 256, 0, ST(stream)>>>((const __half*)x, M, N, ld, rps, workspace);
  else colsum_partial_kernel<float><<<grid, 256, 0, ST(stream)>>>((const float*)x, M, N, ld, rps, workspace);
  reduce_slabs_kernel<<<(N + 255) / 256, 256, 0, ST(stream)>>>(workspace, slabs, N, out, accumulate);
  RET();
}

GVF_API int gvf_ln_bwd_f16(const void* x, int x_is_f16, const void* dy, const void* dres, void* dx, int M, int C,
                           float eps, void* stream) {
  if (!x || !dy || !dx || M <= 0) return GVF_ERR_INVALID;
  int grid = (M + 7) / 8;
  if (grid > 148 * 8) grid = 148 * 8;
#define GVF_LNB(CC)                                                                                               \
  if (C == CC) {                                                                                                  \
    if (x_is_f16)                                                                                                 \
      ln_bwd_kernel<__half, CC><<<grid, 256, 0, ST(stream)>>>((const __half*)x, (const __half*)dy,                 \
                                                              (const __half*)dres, (__half*)dx, M, eps);          \
    else                                                                                                          \
      ln_bwd_kernel<float, CC><<<grid, 256, 0, ST(stream)>>>((const float*)x, (const __half*)dy,                   \
                                                             (const __half*)dres, (__half*)dx, M, eps);           \
    RET();                                                                                                        \
  }
  GVF_LNB(768) GVF_LNB(128) GVF_LNB(96) GVF_LNB(192) GVF_LNB(384) GVF_LNB(512) GVF_LNB(1024)
#undef GVF_LNB
  return GVF_ERR_UNSUPPORTED;
}

GVF_API int gvf_geglu_bwd_f16(const void* h, const void* dG, long long M, int F, void* dh, void* stream) {
  if (!h || !dG || !dh || M <= 0 || (F % 8)) return GVF_ERR_INVALID;
  const long long n8 = M * (F / 8);
  geglu_bwd_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, ST(stream)>>>((const __half*)h, (const __half*)dG, M, F,
                                                                        (__half*)dh);
  RET();
}

GVF_API int gvf_gaussian_tensor_bwd(const gvf_raster_params* prm, int P, const float* scaling, const float* rotation,
                                    const float* opacity, const float* g, float* d_xyz, float* d_dc, float* d_scaling,
                                    float* d_rotation, float* d_opacity, void* stream) {
  if (!prm || !scaling || !rotation || !opacity || !g || !d_xyz || !d_dc || !d_scaling || !d_rotation || !d_opacity || P <= 0)
    return GVF_ERR_INVALID;
  gaussian_tensor_bwd_kernel<<<(P + 255) / 256, 256, 0, ST(stream)>>>(*prm, P, scaling, rotation, opacity, g, d_xyz, d_dc,
                                                                     d_scaling, d_rotation, d_opacity);
  RET();
}

GVF_API int gvf_gelu_tanh_f16(const void* h, long long n, void* out, void* stream) {
  if (!h || !out || n <= 0 || (n % 8)) return GVF_ERR_INVALID;
  gelu_tanh_kernel<<<(unsigned)((n / 8 + 255) / 256), 256, 0, ST(stream)>>>((const __half*)h, n / 8, (__half*)out);
  RET();
}
GVF_API int gvf_gelu_tanh_bwd_f16(const void* h, const void* dy, long long n, void* dh, void* stream) {
  if (!h || !dy || !dh || n <= 0 || (n % 8)) return GVF_ERR_INVALID;
  gelu_tanh_bwd_kernel<<<(unsigned)((n / 8 + 255) / 256), 256, 0, ST(stream)>>>((const __half*)h, (const __half*)dy, n / 8,
                                                                              (__half*)dh);
  RET();
}

GVF_API int gvf_small_linear_bwd_input(const void* dy, int dy_is_f16, long long ld, const void* W, long long M, int N,
                                       int K, float* dx, int ldx, void* stream) {
  if (!dy || !W || !dx || M <= 0 || N <= 0 || K <= 0 || K > 32 || ldx < K) return GVF_ERR_INVALID;
  const unsigned blocks = (unsigned)((M + 7) / 8);
  if (dy_is_f16)
    small_linear_bwd_input_kernel<__half><<<blocks, 256, 0, ST(stream)>>>((const __half*)dy, ld, (const __half*)W, M, N, K,
                                                                         dx, ldx);
  else
    small_linear_bwd_input_kernel<float><<<blocks, 256, 0, ST(stream)>>>((const float*)dy, ld, (const __half*)W, M, N, K,
                                                                        dx, ldx);
  RET();
}

GVF_API int gvf_skinny_outer(const float* x, int ldx, int K, const void* y, int y_is_f16, long long ldy, long long M, int N,
                             float* workspace, size_t workspace_bytes, float* out, int accumulate, void* stream) {
  if (!x || !y || !workspace || !out || M <= 0 || N <= 0 || K <= 0 || K > 16 || ldx < K || ldy < N) return GVF_ERR_INVALID;
  int slabs;
  if (y_is_f16 && (N % 8) == 0 && N / 8 <= 96 && M >= 4096 && (ldy % 8) == 0 && ((uintptr_t)y & 15) == 0) {
    // big fp16 passes: 8 columns per thread, ~2 CTAs per SM
    constexpr int RG = 4;
    long long rps = (M + 295) / 296;
    rps = ((rps + 63) / 64) * 64;
    slabs = (int)((M + rps - 1) / rps);
    if (workspace_bytes < (size_t)slabs * K * N * sizeof(float)) return GVF_ERR_WORKSPACE;
    const int threads = (N / 8) * RG;
    const size_t smem = (size_t)RG * K * N * sizeof(float);
    static bool configured = false;
    if (!configured) {
      if (cudaFuncSetAttribute(skinny_outer_vec8_kernel<RG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) !=
          cudaSuccess)
        return GVF_ERR_CUDA;
      configured = true;
    }
    if (smem <= 200 * 1024 && threads <= 384) {
      skinny_outer_vec8_kernel<RG><<<slabs, threads, smem, ST(stream)>>>(x, ldx, K, (const __half*)y, ldy, M, N, (int)rps,
                                                                         workspace);
      const long long n = (long long)K * N;
      reduce_slabs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>(workspace, slabs, n, out, accumulate);
      RET();
    }
  }
  const int rps = slab_rows(M, &slabs);
  if (workspace_bytes < (size_t)slabs * K * N * sizeof(float)) return GVF_ERR_WORKSPACE;
  const dim3 grid((N + 511) / 512, slabs);
  if (y_is_f16)
    skinny_outer_partial_kernel<__half><<<grid, 256, 0, ST(stream)>>>(x, ldx, K, (const __half*)y, ldy, M, N, rps, workspace);
  else
    skinny_outer_partial_kernel<float><<<grid, 256, 0, ST(stream)>>>(x, ldx, K, (const float*)y, ldy, M, N, rps, workspace);
  const long long n = (long long)K * N;
  reduce_slabs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>(workspace, slabs, n, out, accumulate);
  RET();
}

GVF_API int gvf_skinny_expand(const float* x, int ldx, int K, const float* Wt, long long M, int N, void* out,
                              int out_is_f16, long long ldo, void* stream) {
  if (!x || !Wt || !out || M <= 0 || N <= 0 || K <= 0 || K > 16 || ldx < K || ldo < N) return GVF_ERR_INVALID;
  if (out_is_f16 && (((uintptr_t)out & 3) != 0)) return GVF_ERR_INVALID;
  int threads = (N + 1) / 2;
  threads = threads > 512 ? 512 : ((threads + 31) / 32) * 32;
  const unsigned blocks = (unsigned)((M + 63) / 64);
  if (out_is_f16) skinny_expand_kernel<__half><<<blocks, threads, 0, ST(stream)>>>(x, ldx, K, Wt, M, N, (__half*)out, ldo);
  else skinny_expand_kernel<float><<<blocks, threads, 0, ST(stream)>>>(x, ldx, K, Wt, M, N, (float*)out, ldo);
  RET();
}

GVF_API int gvf_vae_query_embed_bwd(const float* queries, int ldq, const void* gs, const void* dout, int Q, int C,
                                    void* dgs, float* dxyz, int ld_dxyz, int accumulate, void* stream) {
  if (!queries || !gs || !dout || !dgs || !dxyz || Q <= 0 || ld_dxyz < 3) return GVF_ERR_INVALID;
  const dim3 grid((Q + 7) / 8);
  if (C == 768)
    query_embed_bwd_kernel<768><<<grid, 256, 0, ST(stream)>>>(queries, ldq, (const __half*)gs, (const __half*)dout, Q,
                                                              (__half*)dgs, dxyz, ld_dxyz, accumulate);
  else if (C == 96)
    query_embed_bwd_kernel<96><<<grid, 256, 0, ST(stream)>>>(queries, ldq, (const __half*)gs, (const __half*)dout, Q,
                                                             (__half*)dgs, dxyz, ld_dxyz, accumulate);
  else if (C == 384)
    query_embed_bwd_kernel<384><<<grid, 256, 0, ST(stream)>>>(queries, ldq, (const __half*)gs, (const __half*)dout, Q,
                                                              (__half*)dgs, dxyz, ld_dxyz, accumulate);
  else if (C == 192)
    query_embed_bwd_kernel<192><<<grid, 256, 0, ST(stream)>>>(queries, ldq, (const __half*)gs, (const __half*)dout, Q,
                                                              (__half*)dgs, dxyz, ld_dxyz, accumulate);
  else return GVF_ERR_UNSUPPORTED;
  RET();
}

}  // extern "C"
