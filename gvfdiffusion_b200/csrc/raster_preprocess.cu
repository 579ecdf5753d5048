// raster_preprocess.cu -- per-(frame, Gaussian) activation + projection + tile counting.
//
// Replaces, fused in one pass: GaussianModel.get_{xyz,scaling,rotation,features,opacity}
// _with_delta (reference representations/gaussian/gaussian_model.py:98-114, five
// elementwise torch kernels + dtype copies per render, renderers/gaussian_render.py:154-193)
// and the `preprocessCUDA` stage of diff_gaussian_rasterization (mip-splatting fork; call
// site renderers/gaussian_render.py:198-206).
//
// COMPILED WITH -fmad=false.  Everything that feeds an integer output (radius, tile
// rectangle, depth-sort key) is evaluated in the exact operation order of the CPU oracle
// with correctly rounded IEEE ops and the reproducible transcendentals of gvf_math.h, so
// radii / rectangles / keys are bit-identical to the oracle's (north_star: "bit-exact on
// tile/sort indices").
//
// HBM-bound: reads 56 B canonical + 56 B delta per (frame, Gaussian), writes a 48 B splat
// record + 8 B rect (+4 B radius); one global atomic per touched tile (1-4 typical).
#include <stdlib.h>
#include <cooperative_groups.h>
#include "../../include/gvf_math.h"
#include "raster_common.h"

namespace gvf {

__device__ __forceinline__ int f2i_clamped(float v) {
  v = fminf(fmaxf(v, -1.0e6f), 1.0e6f);
  return (int)v;
}

struct PreArgs {
  gvf_raster_params prm;
  int F, P, activated;
  int views;           // consecutive frames that share one delta row (cameras per timestep), >= 1
  const float *xyz, *dc, *scaling, *rotation, *opacity, *delta, *cams;
  float4* splat;
  ushort4* rect;
  uint32_t* tile_count;
  int32_t* radii;
};

__global__ void __launch_bounds__(256) preprocess_kernel(const PreArgs a) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long FP = (long long)a.F * a.P;
  if (gid >= FP) return;
  const int f = (int)(gid / a.P);
  const int i = (int)(gid - (long long)f * a.P);
  const gvf_raster_params& prm = a.prm;
  const int W = prm.W, H = prm.H;
  const int gx = (W + GVF_TILE - 1) / GVF_TILE, gy = (H + GVF_TILE - 1) / GVF_TILE;

  // ---------------- activation (GaussianModel.get_*_with_delta) ----------------
  float m3[3], sc[3], q[4], sh[3], opac;
  if (a.activated) {
    const size_t b = (size_t)gid;
    m3[0] = a.xyz[b * 3 + 0]; m3[1] = a.xyz[b * 3 + 1]; m3[2] = a.xyz[b * 3 + 2];
    sh[0] = a.dc[b * 3 + 0]; sh[1] = a.dc[b * 3 + 1]; sh[2] = a.dc[b * 3 + 2];
    sc[0] = a.scaling[b * 3 + 0]; sc[1] = a.scaling[b * 3 + 1]; sc[2] = a.scaling[b * 3 + 2];
    q[0] = a.rotation[b * 4 + 0]; q[1] = a.rotation[b * 4 + 1];
    q[2] = a.rotation[b * 4 + 2]; q[3] = a.rotation[b * 4 + 3];
    opac = a.opacity[b];
  } else {
    float d[14];
    if (a.delta) {
      // [F,P,14] rows are 56 B: 8-byte aligned -> seven float2 loads
      const float2* dp = reinterpret_cast<const float2*>(a.delta + ((size_t)(f / a.views) * a.P + i) * 14);
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        const float2 v = __ldg(dp + k);
        d[2 * k] = v.x;
        d[2 * k + 1] = v.y;
      }
    }
    const float k2 = prm.min_kernel * prm.min_kernel;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = a.xyz[(size_t)i * 3 + c] * prm.aabb[3 + c] + prm.aabb[c];
      m3[c] = a.delta ? v + d[c] : v;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float s = a.scaling[(size_t)i * 3 + c] + prm.scale_bias;
      if (a.delta) s = s + d[3 + c];
      s = prm.softplus ? gvf_softplusf(s) : gvf_expf(s);
      sc[c] = sqrtf(s * s + k2);
    }
    float qq[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float r = a.rotation[(size_t)i * 4 + c] + (c == 0 ? 1.0f : 0.0f);
      qq[c] = a.delta ? r + d[6 + c] : r;
    }
    float n = sqrtf(qq[0] * qq[0] + qq[1] * qq[1] + qq[2] * qq[2] + qq[3] * qq[3]);
    n = fmaxf(n, 1e-12f);
#pragma unroll
    for (int c = 0; c < 4; ++c) q[c] = qq[c] / n;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      sh[c] = a.delta ? a.dc[(size_t)i * 3 + c] + d[10 + c] : a.dc[(size_t)i * 3 + c];
    float o = a.opacity[i] + prm.opacity_bias;
    if (a.delta) o = o + d[13];
    opac = gvf_sigmoidf(o);
  }

  // ---------------- projection ----------------
  const float* view = a.cams + (size_t)f * 32;
  const float* proj = view + 16;
  const float px = m3[0], py = m3[1], pz = m3[2];
  const float tx = view[0] * px + view[4] * py + view[8] * pz + view[12];
  const float ty = view[1] * px + view[5] * py + view[9] * pz + view[13];
  const float tz = view[2] * px + view[6] * py + view[10] * pz + view[14];

  int radius = 0, x0 = 0, y0 = 0, x1 = 0, y1 = 0;
  float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0, r2 = r0;
  bool ok = tz > 0.2f;
  if (ok) {
    const float hx = proj[0] * px + proj[4] * py + proj[8] * pz + proj[12];
    const float hy = proj[1] * px + proj[5] * py + proj[9] * pz + proj[13];
    const float hw = proj[3] * px + proj[7] * py + proj[11] * pz + proj[15];
    const float pw = 1.0f / (hw + 0.0000001f);
    const float ndcx = hx * pw, ndcy = hy * pw;

    const float s0 = prm.scale_modifier * sc[0], s1 = prm.scale_modifier * sc[1],
                s2 = prm.scale_modifier * sc[2];
    const float r = q[0], x = q[1], y = q[2], z = q[3];
    const float R00 = 1.f - 2.f * (y * y + z * z), R01 = 2.f * (x * y - r * z), R02 = 2.f * (x * z + r * y);
    const float R10 = 2.f * (x * y + r * z), R11 = 1.f - 2.f * (x * x + z * z), R12 = 2.f * (y * z - r * x);
    const float R20 = 2.f * (x * z - r * y), R21 = 2.f * (y * z + r * x), R22 = 1.f - 2.f * (x * x + y * y);
    const float L00 = R00 * s0, L01 = R01 * s1, L02 = R02 * s2;
    const float L10 = R10 * s0, L11 = R11 * s1, L12 = R12 * s2;
    const float L20 = R20 * s0, L21 = R21 * s1, L22 = R22 * s2;
    const float S00 = L00 * L00 + L01 * L01 + L02 * L02;
    const float S01 = L00 * L10 + L01 * L11 + L02 * L12;
    const float S02 = L00 * L20 + L01 * L21 + L02 * L22;
    const float S11 = L10 * L10 + L11 * L11 + L12 * L12;
    const float S12 = L10 * L20 + L11 * L21 + L12 * L22;
    const float S22 = L20 * L20 + L21 * L21 + L22 * L22;

    const float fx = (float)W / (2.0f * prm.tanfovx), fy = (float)H / (2.0f * prm.tanfovy);
    const float limx = 1.3f * prm.tanfovx, limy = 1.3f * prm.tanfovy;
    const float txtz = tx / tz, tytz = ty / tz;
    const float cx = fminf(limx, fmaxf(-limx, txtz)) * tz;
    const float cy = fminf(limy, fmaxf(-limy, tytz)) * tz;
    const float J00 = fx / tz, J02 = -(fx * cx) / (tz * tz);
    const float J11 = fy / tz, J12 = -(fy * cy) / (tz * tz);
    const float W00 = view[0], W01 = view[4], W02 = view[8];
    const float W10 = view[1], W11 = view[5], W12 = view[9];
    const float W20 = view[2], W21 = view[6], W22 = view[10];
    const float T00 = J00 * W00 + J02 * W20, T01 = J00 * W01 + J02 * W21, T02 = J00 * W02 + J02 * W22;
    const float T10 = J11 * W10 + J12 * W20, T11 = J11 * W11 + J12 * W21, T12 = J11 * W12 + J12 * W22;
    const float V00 = T00 * S00 + T01 * S01 + T02 * S02;
    const float V01 = T00 * S01 + T01 * S11 + T02 * S12;
    const float V02 = T00 * S02 + T01 * S12 + T02 * S22;
    const float V10 = T10 * S00 + T11 * S01 + T12 * S02;
    const float V11 = T10 * S01 + T11 * S11 + T12 * S12;
    const float V12 = T10 * S02 + T11 * S12 + T12 * S22;
    float ca = V00 * T00 + V01 * T01 + V02 * T02;
    const float cb = V00 * T10 + V01 * T11 + V02 * T12;
    float cc = V10 * T10 + V11 * T11 + V12 * T12;

    const float ks = prm.kernel_size;
    const float det0 = fmaxf(1e-6f, ca * cc - cb * cb);
    const float det1 = fmaxf(1e-6f, (ca + ks) * (cc + ks) - cb * cb);
    float coef = sqrtf(det0 / (det1 + 1e-6f) + 1e-6f);
    if (det0 <= 1e-6f || det1 <= 1e-6f) coef = 0.0f;
    if (!prm.mip_filter) coef = 1.0f;            // plain 3DGS dilation: no opacity compensation
    ca = ca + ks;
    cc = cc + ks;
    const float det = ca * cc - cb * cb;
    ok = det != 0.0f;
    if (ok) {
      const float det_inv = 1.0f / det;
      const float mid = 0.5f * (ca + cc);
      const float disc = sqrtf(fmaxf(0.1f, mid * mid - det));
      const float lambda1 = mid + disc, lambda2 = mid - disc;
      const float rad = ceilf(3.0f * sqrtf(fmaxf(lambda1, lambda2)));
      const float pix_x = ((ndcx + 1.0f) * (float)W - 1.0f) * 0.5f;
      const float pix_y = ((ndcy + 1.0f) * (float)H - 1.0f) * 0.5f;
      const int irad = f2i_clamped(rad);
      x0 = f2i_clamped((pix_x - (float)irad) / (float)GVF_TILE);
      y0 = f2i_clamped((pix_y - (float)irad) / (float)GVF_TILE);
      x1 = f2i_clamped((pix_x + (float)irad + (float)(GVF_TILE - 1)) / (float)GVF_TILE);
      y1 = f2i_clamped((pix_y + (float)irad + (float)(GVF_TILE - 1)) / (float)GVF_TILE);
      x0 = x0 < 0 ? 0 : (x0 > gx ? gx : x0);
      y0 = y0 < 0 ? 0 : (y0 > gy ? gy : y0);
      x1 = x1 < 0 ? 0 : (x1 > gx ? gx : x1);
      y1 = y1 < 0 ? 0 : (y1 > gy ? gy : y1);
      ok = (x1 - x0) * (y1 - y0) != 0;
      if (ok) {
        radius = irad;
        const float SH_C0 = 0.28209479177387814f;
        r0 = make_float4(pix_x, pix_y, cc * det_inv, -cb * det_inv);
        r1 = make_float4(ca * det_inv, opac * coef, fmaxf(SH_C0 * sh[0] + 0.5f, 0.0f),
                         fmaxf(SH_C0 * sh[1] + 0.5f, 0.0f));
        r2 = make_float4(fmaxf(SH_C0 * sh[2] + 0.5f, 0.0f), tz, __int_as_float(irad),
                         __int_as_float((x1 - x0) * (y1 - y0)));
      }
    }
  }
  if (!ok) { x0 = y0 = x1 = y1 = 0; }
  float4* sp = a.splat + (size_t)gid * 3;
  sp[0] = r0;
  sp[1] = r1;
  sp[2] = r2;
  a.rect[gid] = make_ushort4((unsigned short)x0, (unsigned short)y0, (unsigned short)x1,
                             (unsigned short)y1);
  if (a.radii) a.radii[gid] = radius;
  if (ok) {
    uint32_t* tc = a.tile_count + (size_t)f * gx * gy;
    for (int yy = y0; yy < y1; ++yy)
      for (int xx = x0; xx < x1; ++xx) atomicAdd(tc + yy * gx + xx, 1u);
  }
}

cudaError_t launch_preprocess(const gvf_raster_params& prm, int F, int P, int activated,
                              const float* xyz, const float* dc, const float* scaling,
                              const float* rotation, const float* opacity, const float* delta,
                              const float* cams, const RasterWs& ws, int32_t* radii,
                              cudaStream_t st, int views) {
  PreArgs a;
  a.prm = prm; a.F = F; a.P = P; a.activated = activated; a.views = views < 1 ? 1 : views;
  a.xyz = xyz; a.dc = dc; a.scaling = scaling; a.rotation = rotation; a.opacity = opacity;
  a.delta = delta; a.cams = cams;
  a.splat = ws.splat; a.rect = ws.rect; a.tile_count = ws.tile_count; a.radii = radii;
  const long long n = (long long)F * P;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  preprocess_kernel<<<blocks, 256, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace gvf

// =======================================================================================
// get_gaussian_tensor (reference train_vae.py:466-472): activated canonical Gaussians packed
// [xyz3 | rgb(features_dc)3 | opacity1 | scale3 | rot4] = the motion-VAE decoder queries and
// the source of `static_latent` / `deformation_position_xyz` (inference_dpm_latent.py:205-216).
// Same activation code (and -fmad=false) as the rasteriser preprocess above.
namespace gvf {

__global__ void __launch_bounds__(256) gaussian_tensor_kernel(const gvf_raster_params prm, int P,
                                                              const float* __restrict__ xyz,
                                                              const float* __restrict__ dc,
                                                              const float* __restrict__ scaling,
                                                              const float* __restrict__ rotation,
                                                              const float* __restrict__ opacity,
                                                              float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  float* o = out + (size_t)i * 14;
  const float k2 = prm.min_kernel * prm.min_kernel;
#pragma unroll
  for (int c = 0; c < 3; ++c) o[c] = xyz[(size_t)i * 3 + c] * prm.aabb[3 + c] + prm.aabb[c];
#pragma unroll
  for (int c = 0; c < 3; ++c) o[3 + c] = dc[(size_t)i * 3 + c];
  o[6] = gvf_sigmoidf(opacity[i] + prm.opacity_bias);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float s = scaling[(size_t)i * 3 + c] + prm.scale_bias;
    s = prm.softplus ? gvf_softplusf(s) : gvf_expf(s);
    o[7 + c] = sqrtf(s * s + k2);
  }
  float q[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) q[c] = rotation[(size_t)i * 4 + c] + (c == 0 ? 1.0f : 0.0f);
  float n = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  n = fmaxf(n, 1e-12f);
#pragma unroll
  for (int c = 0; c < 4; ++c) o[10 + c] = q[c] / n;
}

// Farthest point sampling (replaces torch_cluster.fps as used by sample_gs,
// reference utils/inference_utils.py:180-198).  One CTA per cloud, 1024 threads; deterministic:
// starts at point `start` (torch_cluster's default start is random), ties -> lowest index.
__global__ void __launch_bounds__(1024) fps_kernel(const float* __restrict__ pts, int ld, int P, int K,
                                                   int start, float* __restrict__ mind,
                                                   int* __restrict__ out_idx) {
  __shared__ float sd[32];
  __shared__ int si[32];
  __shared__ int cur_s;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const float* p = pts + (size_t)blockIdx.y * 0;   // single cloud per launch (blockIdx unused)
  for (int i = tid; i < P; i += 1024) mind[i] = 3.0e38f;
  int cur = start;
  if (tid == 0) out_idx[0] = cur;
  __syncthreads();
  for (int k = 1; k < K; ++k) {
    const float cx = p[(size_t)cur * ld], cy = p[(size_t)cur * ld + 1], cz = p[(size_t)cur * ld + 2];
    float best = -1.0f;
    int bi = 0x7fffffff;
    for (int i = tid; i < P; i += 1024) {
      const float dx = p[(size_t)i * ld] - cx, dy = p[(size_t)i * ld + 1] - cy, dz = p[(size_t)i * ld + 2] - cz;
      const float d = fminf(mind[i], dx * dx + dy * dy + dz * dz);
      mind[i] = d;
      if (d > best) { best = d; bi = i; }     // ascending i per thread: first max kept
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) { sd[w] = best; si[w] = bi; }
    __syncthreads();
    if (w == 0) {
      best = sd[lane];
      bi = si[lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (lane == 0) { cur_s = bi; out_idx[k] = bi; }
    }
    __syncthreads();
    cur = cur_s;
  }
}

// Fast path for clouds that fit shared memory (P <= 16384): coordinates live in smem as SoA, each
// thread keeps the running minimum distances of its PER points in registers, and a sample costs one
// __syncthreads: warp arg-max by redux.sync (distances are >= +0, so their bit patterns order like
// the floats), double-buffered per-warp results, every warp re-reduces the 32 partials redundantly.
// Same arithmetic, same tie rule (lowest index) as fps_kernel: identical indices.
template <int PER>
__global__ void __launch_bounds__(1024, 1) fps_smem_kernel(const float* __restrict__ pts, int ld, int P, int K,
                                                            int start, int* __restrict__ out_idx) {
  extern __shared__ float sxyz[];                  // x[0..Pp) | y | z, Pp = PER * 1024
  __shared__ unsigned sd[2][32];
  __shared__ unsigned si[2][32];
  constexpr int Pp = PER * 1024;
  float* sx = sxyz; float* sy = sxyz + Pp; float* sz = sxyz + 2 * Pp;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  float md[PER];
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int i = tid + 1024 * j;
    const bool ok = i < P;
    sx[i] = ok ? pts[(size_t)i * ld] : 0.f;
    sy[i] = ok ? pts[(size_t)i * ld + 1] : 0.f;
    sz[i] = ok ? pts[(size_t)i * ld + 2] : 0.f;
    md[j] = ok ? 3.0e38f : -1.0f;                  // padding never wins: min(-1, d) = -1 < any distance
  }
  unsigned cur = (unsigned)start;
  if (tid == 0) out_idx[0] = start;
  __syncthreads();
  for (int k = 1; k < K; ++k) {
    const float cx = sx[cur], cy = sy[cur], cz = sz[cur];
    float best = -1.0f;
    unsigned bi = 0xffffffffu;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int i = tid + 1024 * j;
      const float dx = sx[i] - cx, dy = sy[i] - cy, dz = sz[i] - cz;
      const float d = fminf(md[j], dx * dx + dy * dy + dz * dz);
      md[j] = d;
      if (d > best) { best = d; bi = (unsigned)i; }   // ascending i per thread: first max kept
    }
    unsigned ub = best < 0.f ? 0u : __float_as_uint(best);
    unsigned wm = __reduce_max_sync(0xffffffffu, ub);
    unsigned wi = __reduce_min_sync(0xffffffffu, ub == wm ? bi : 0xffffffffu);
    if (lane == 0) { sd[k & 1][w] = wm; si[k & 1][w] = wi; }
    __syncthreads();
    ub = sd[k & 1][lane];
    bi = si[k & 1][lane];
    wm = __reduce_max_sync(0xffffffffu, ub);
    cur = __reduce_min_sync(0xffffffffu, ub == wm ? bi : 0xffffffffu);
    if (tid == 0) out_idx[k] = (int)cur;
  }
}

// Second generation of the single-CTA kernel.  Per sample the first one re-read all three coordinate arrays
// from shared memory (196 KB = 1536 clocks of the SM's 128 B / clk) and spent as long again issuing twelve
// instructions per point.  Here every thread keeps the x coordinates of its PER points in registers next to
// their running minimum distances (y and z still come from shared memory: the register file cannot hold all
// three at 1024 threads), and the arithmetic runs on packed pairs of points (sub / mul / add .f32x2: the same
// individually rounded operations in the same order, half the instructions).  Identical indices.
__device__ __forceinline__ void sub2(float& d0, float& d1, float a0, float a1, float b) {
  asm("{\n.reg .b64 ra, rb, rd;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %4};\nsub.rn.f32x2 rd, ra, rb;\n"
      "mov.b64 {%0, %1}, rd;\n}" : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b));
}
__device__ __forceinline__ void mul2(float& d0, float& d1, float a0, float a1) {      // (a0*a0, a1*a1)
  asm("{\n.reg .b64 ra, rd;\nmov.b64 ra, {%2, %3};\nmul.rn.f32x2 rd, ra, ra;\nmov.b64 {%0, %1}, rd;\n}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1));
}
__device__ __forceinline__ void add2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{\n.reg .b64 ra, rb, rd;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\nadd.rn.f32x2 rd, ra, rb;\n"
      "mov.b64 {%0, %1}, rd;\n}" : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
template <int PER>
__global__ void __launch_bounds__(1024, 1) fps_smem2_kernel(const float* __restrict__ pts, int ld, int P, int K,
                                                             int start, int* __restrict__ out_idx) {
  extern __shared__ float sxyz[];                  // x[0..Pp) | y | z, Pp = PER * 1024
  __shared__ unsigned sd[2][32];
  __shared__ unsigned si[2][32];
  constexpr int Pp = PER * 1024;
  float* sx = sxyz; float* sy = sxyz + Pp; float* sz = sxyz + 2 * Pp;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  float md[PER], px[PER];
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int i = tid + 1024 * j;
    const bool ok = i < P;
    px[j] = ok ? pts[(size_t)i * ld] : 0.f;
    sx[i] = px[j];
    sy[i] = ok ? pts[(size_t)i * ld + 1] : 0.f;
    sz[i] = ok ? pts[(size_t)i * ld + 2] : 0.f;
    md[j] = ok ? 3.0e38f : -1.0f;                  // padding never wins: min(-1, d) = -1 < any distance
  }
  unsigned cur = (unsigned)start;
  if (tid == 0) out_idx[0] = start;
  __syncthreads();
  for (int k = 1; k < K; ++k) {
    const float cx = sx[cur], cy = sy[cur], cz = sz[cur];
    float best = -1.0f;
    unsigned bi = 0xffffffffu;
#pragma unroll
    for (int j = 0; j < PER; j += 2) {
      const int i0 = tid + 1024 * j, i1 = i0 + 1024;
      float dx0, dx1, dy0, dy1, dz0, dz1, s0, s1;
      sub2(dx0, dx1, px[j], px[j + 1], cx);
      sub2(dy0, dy1, sy[i0], sy[i1], cy);
      sub2(dz0, dz1, sz[i0], sz[i1], cz);
      mul2(dx0, dx1, dx0, dx1);
      mul2(dy0, dy1, dy0, dy1);
      mul2(dz0, dz1, dz0, dz1);
      add2(s0, s1, dx0, dx1, dy0, dy1);
      add2(s0, s1, s0, s1, dz0, dz1);               // (dx*dx + dy*dy) + dz*dz, every operation rounded
      const float d0 = fminf(md[j], s0), d1 = fminf(md[j + 1], s1);
      md[j] = d0; md[j + 1] = d1;
      if (d0 > best) { best = d0; bi = (unsigned)i0; }   // ascending i per thread: first max kept
      if (d1 > best) { best = d1; bi = (unsigned)i1; }
    }
    unsigned ub = best < 0.f ? 0u : __float_as_uint(best);
    unsigned wm = __reduce_max_sync(0xffffffffu, ub);
    unsigned wi = __reduce_min_sync(0xffffffffu, ub == wm ? bi : 0xffffffffu);
    if (lane == 0) { sd[k & 1][w] = wm; si[k & 1][w] = wi; }
    __syncthreads();
    ub = sd[k & 1][lane];
    bi = si[k & 1][lane];
    wm = __reduce_max_sync(0xffffffffu, ub);
    cur = __reduce_min_sync(0xffffffffu, ub == wm ? bi : 0xffffffffu);
    if (tid == 0) out_idx[k] = (int)cur;
  }
}

// Third generation: exact pruning.  Every thread owns PER CONSECUTIVE points (thread (w, l): points [PER (32 w + l), + PER)),
// keeps their bounding box in registers next to their running minimum distances, and caches its candidate (its largest
// running minimum and the lowest index that attains it).  Rows of the path's clouds are voxel-major with the voxels in
// lexicographic order (sparse/basic.py layout, `argwhere` of the occupancy grid), so 16 consecutive points are two
// neighbouring voxels and the box is small.  A new sample c can only lower a running minimum that is larger than the
// point's distance to c, so a thread whose box is at least `best` away from c has nothing to update and its cached
// candidate stays valid; a warp in which no thread is active goes straight to the reductions.  The bound is evaluated with
// the SAME individually rounded operations as the distance itself ((ex*ex + ey*ey) + ez*ez, ex = the per-axis gap to the
// box): subtraction, multiplication and addition are monotone under round-to-nearest, so bound <= distance holds in floating
// point, not just in exact arithmetic, and the skipped points' minima are provably unchanged -- the selected indices are
// those of the brute-force kernels bit for bit, whatever the row order (an unordered cloud only prunes less and then costs
// what the second generation costs plus the box test).  Shared memory holds the coordinates transposed per warp
// (point j of lane l at 32 PER w + 32 j + l) so that the lanes' loads stay conflict-free.  Tie rule unchanged (lowest index).
template <int PER, int NT = 1024>
__global__ void __launch_bounds__(NT, 1) fps_pruned_kernel(const float* __restrict__ pts, int ld, int P, int K,
                                                              int start, int* __restrict__ out_idx) {
  extern __shared__ float sxyz[];                  // x[0..Pp) | y | z, Pp = PER * 1024, transposed per warp
  __shared__ unsigned sd[2][32];
  __shared__ unsigned si[2][32];
  constexpr int Pp = PER * NT, NW = NT / 32;       // NT = 512: half the per-step warp overhead, twice the points per thread
  float* sx = sxyz; float* sy = sxyz + Pp; float* sz = sxyz + 2 * Pp;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int first = (w * 32 + lane) * PER;         // this thread's first point
  const int sbase = w * 32 * PER + lane;           // its shared-memory slot for j = 0 (stride 32 per j)
  float md[PER], px[PER];
  float lox = 3.0e38f, loy = 3.0e38f, loz = 3.0e38f, hix = -3.0e38f, hiy = -3.0e38f, hiz = -3.0e38f;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int i = first + j;
    const bool ok = i < P;
    const float x = ok ? pts[(size_t)i * ld] : 0.f, y = ok ? pts[(size_t)i * ld + 1] : 0.f, z = ok ? pts[(size_t)i * ld + 2] : 0.f;
    px[j] = x; sx[sbase + 32 * j] = x; sy[sbase + 32 * j] = y; sz[sbase + 32 * j] = z;
    md[j] = ok ? 3.0e38f : -1.0f;                  // padding never wins: min(-1, d) = -1 < any distance
    if (ok) {
      lox = fminf(lox, x); loy = fminf(loy, y); loz = fminf(loz, z);
      hix = fmaxf(hix, x); hiy = fmaxf(hiy, y); hiz = fmaxf(hiz, z);
    }
  }
  float best = first < P ? 3.0e38f : -1.0f;        // cached candidate of this thread
  unsigned bi = first < P ? (unsigned)first : 0xffffffffu;
  unsigned cur = (unsigned)start;
  if (tid == 0) out_idx[0] = start;
  __syncthreads();
  for (int k = 1; k < K; ++k) {
    const unsigned cw = cur / (32 * PER), cr = cur % (32 * PER);
    const unsigned ca = cw * (32 * PER) + (cr % PER) * 32 + cr / PER;
    const float cx = sx[ca], cy = sy[ca], cz = sz[ca];
    // gap to the box per axis, then the distance's own operation sequence on the gaps
    const float ex = cx < lox ? __fsub_rn(lox, cx) : (cx > hix ? __fsub_rn(cx, hix) : 0.f);
    const float ey = cy < loy ? __fsub_rn(loy, cy) : (cy > hiy ? __fsub_rn(cy, hiy) : 0.f);
    const float ez = cz < loz ? __fsub_rn(loz, cz) : (cz > hiz ? __fsub_rn(cz, hiz) : 0.f);
    const float lb = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez));
    const bool active = lb < best;                  // best < 0 (padding only): never
    if (__any_sync(0xffffffffu, active)) {
      if (active) {
        float nb = -1.0f;
        unsigned ni = 0xffffffffu;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
          const float dx = __fsub_rn(px[j], cx), dy = __fsub_rn(sy[sbase + 32 * j], cy), dz = __fsub_rn(sz[sbase + 32 * j], cz);
          const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
          const float m = fminf(md[j], d);
          md[j] = m;
          if (m > nb) { nb = m; ni = (unsigned)(first + j); }     // ascending index: first maximum kept
        }
        best = nb; bi = ni;
      }
    }
    unsigned ub = best < 0.f ? 0u : __float_as_uint(best);
    unsigned wm = __reduce_max_sync(0xffffffffu, ub);
    const unsigned wi = __reduce_min_sync(0xffffffffu, ub == wm ? bi : 0xffffffffu);
    if (lane == 0) { sd[k & 1][w] = wm; si[k & 1][w] = wi; }
    __syncthreads();
    ub = lane < NW ? sd[k & 1][lane] : 0u;
    const unsigned ci = lane < NW ? si[k & 1][lane] : 0xffffffffu;
    wm = __reduce_max_sync(0xffffffffu, ub);
    cur = __reduce_min_sync(0xffffffffu, ub == wm ? ci : 0xffffffffu);
    if (tid == 0) out_idx[k] = (int)cur;
  }
}

// Cluster variant: the cloud is spread over the 8 CTAs of one thread-block cluster, 512 threads each, every
// thread keeps the coordinates and running minimum distances of its PER points in REGISTERS (the single-CTA
// kernel above re-reads all 196 KB of coordinates from shared memory for every sample: 1536 clocks of
// shared-memory bandwidth out of the 2770 a sample took).  Per sample: register-only distance update, warp
// arg-max by redux, one __syncthreads, the CTA's candidate (distance, index, x, y, z) written to its own
// shared memory, one cluster barrier, and every CTA reads the 8 candidates through distributed shared memory.
// Same arithmetic and tie rule (lowest index) as fps_kernel: identical indices.
namespace cg = cooperative_groups;
constexpr int kFpsCluster = 8, kFpsThreads = 512;

template <int PER>
__global__ void __launch_bounds__(kFpsThreads, 1) fps_cluster_kernel(const float* __restrict__ pts, int ld, int P,
                                                                     int K, int start, int* __restrict__ out_idx) {
  __shared__ float wslot[2][16][5];                // per warp: (distance bits, index bits, x, y, z)
  __shared__ float cand[2][8];                     // this CTA's best, same five fields
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  float px[PER], py[PER], pz[PER], md[PER];
  int gi[PER];
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    // global point index: interleaved so that every CTA owns the same share of any prefix of the cloud
    const int i = (j * kFpsThreads + tid) * kFpsCluster + (int)rank;
    gi[j] = i;
    const bool ok = i < P;
    px[j] = ok ? pts[(size_t)i * ld] : 0.f;
    py[j] = ok ? pts[(size_t)i * ld + 1] : 0.f;
    pz[j] = ok ? pts[(size_t)i * ld + 2] : 0.f;
    md[j] = ok ? 3.0e38f : -1.0f;
  }
  float cx = pts[(size_t)start * ld], cy = pts[(size_t)start * ld + 1], cz = pts[(size_t)start * ld + 2];
  if (rank == 0 && tid == 0) out_idx[0] = start;
  // arg-max over the 32 lanes of (distance bits, then lowest index); returns the winning lane
  auto warp_best = [&](unsigned ub, unsigned bi, unsigned& wm, unsigned& wi) {
    wm = __reduce_max_sync(0xffffffffu, ub);
    wi = __reduce_min_sync(0xffffffffu, ub == wm ? bi : 0xffffffffu);
    return __ffs(__ballot_sync(0xffffffffu, ub == wm && bi == wi)) - 1;
  };
  for (int k = 1; k < K; ++k) {
    float best = -1.0f, bx = 0.f, by = 0.f, bz = 0.f;
    unsigned bi = 0xffffffffu;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const float dx = px[j] - cx, dy = py[j] - cy, dz = pz[j] - cz;
      const float d = fminf(md[j], dx * dx + dy * dy + dz * dz);
      md[j] = d;
      if (d > best) { best = d; bi = (unsigned)gi[j]; bx = px[j]; by = py[j]; bz = pz[j]; }   // first max kept
    }
    const int b = k & 1;
    unsigned wm, wi;
    int src = warp_best(best < 0.f ? 0u : __float_as_uint(best), bi, wm, wi);
    if (lane == src) {
      wslot[b][w][0] = __uint_as_float(wm); wslot[b][w][1] = __uint_as_float(wi);
      wslot[b][w][2] = bx; wslot[b][w][3] = by; wslot[b][w][4] = bz;
    }
    __syncthreads();
    if (w == 0) {
      const int l = lane & 15;
      const unsigned ub = lane < 16 ? __float_as_uint(wslot[b][l][0]) : 0u;
      const unsigned ui = lane < 16 ? __float_as_uint(wslot[b][l][1]) : 0xffffffffu;
      src = warp_best(ub, ui, wm, wi);
      if (lane == src) {
        cand[b][0] = __uint_as_float(wm); cand[b][1] = __uint_as_float(wi);
        cand[b][2] = wslot[b][l][2]; cand[b][3] = wslot[b][l][3]; cand[b][4] = wslot[b][l][4];
      }
    }
    cluster.sync();
    // every warp reduces the 8 CTA candidates itself (lane r reads CTA r's slot through distributed smem)
    unsigned cm = 0u, ci = 0xffffffffu;
    float nx = 0.f, ny = 0.f, nz = 0.f;
    if (lane < kFpsCluster) {
      const float* rc = cluster.map_shared_rank(&cand[b][0], lane);
      cm = __float_as_uint(rc[0]);
      ci = __float_as_uint(rc[1]);
      nx = rc[2]; ny = rc[3]; nz = rc[4];
    }
    src = warp_best(cm, ci, wm, wi);
    cx = __shfl_sync(0xffffffffu, nx, src);
    cy = __shfl_sync(0xffffffffu, ny, src);
    cz = __shfl_sync(0xffffffffu, nz, src);
    if (rank == 0 && tid == 0) out_idx[k] = (int)wi;
  }
  cluster.sync();                                  // nobody reads a peer's shared memory after it has exited
}

}  // namespace gvf

extern "C" GVF_API int gvf_gaussian_tensor(const gvf_raster_params* prm, int P, const float* xyz,
                                           const float* dc, const float* scaling, const float* rotation,
                                           const float* opacity, float* out, void* stream) {
  if (!prm || !xyz || !dc || !scaling || !rotation || !opacity || !out || P <= 0) return GVF_ERR_INVALID;
  gvf::gaussian_tensor_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*prm, P, xyz, dc, scaling,
                                                                                rotation, opacity, out);
  return cudaGetLastError() == cudaSuccess ? GVF_OK : GVF_ERR_CUDA;
}

static int fps_impl(const float* pts, int ld, int P, int K, int start, float* workspace, int32_t* out_idx, void* stream,
                    bool ordered);
extern "C" GVF_API int gvf_fps(const float* pts, int ld, int P, int K, int start, float* workspace,
                               int32_t* out_idx, void* stream) {
  return fps_impl(pts, ld, P, K, start, workspace, out_idx, stream, false);
}
// Same result, for clouds whose rows are spatially ordered (voxel-major Gaussians of lexicographically ordered voxels: every
// cloud the path samples): the exactly pruned kernel, 3.1 ms instead of 5.0 at 16384 -> 4096.  On an unordered cloud it is
// still exact but slower than gvf_fps (6.3 ms), hence the separate entry point.
extern "C" GVF_API int gvf_fps_ordered(const float* pts, int ld, int P, int K, int start, float* workspace,
                                       int32_t* out_idx, void* stream) {
  return fps_impl(pts, ld, P, K, start, workspace, out_idx, stream, true);
}
static int fps_impl(const float* pts, int ld, int P, int K, int start, float* workspace, int32_t* out_idx, void* stream,
                    bool ordered) {
  if (!pts || !workspace || !out_idx || P <= 0 || K <= 0 || K > P || start < 0 || start >= P || ld < 3)
    return GVF_ERR_INVALID;
  auto launch_smem = [&](auto kern, int per) {
    const int bytes = per * 1024 * 3 * (int)sizeof(float);
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) return false;
    kern<<<1, 1024, bytes, (cudaStream_t)stream>>>(pts, ld, P, K, start, out_idx);
    return true;
  };
  bool ok = true;
  // MEASURED (tools/fps_bench.py, 16384 -> 4096): cluster kernel 7.2 - 9.2 ms, single-CTA shared-memory kernel 5.66 ms,
  // its second generation (x in registers, packed f32x2 arithmetic) 5.02 ms, the exactly pruned third generation = the default --
  // the cluster barrier costs more per sample (~1.7 us) than the shared-memory re-read it removes, so the
  // single-CTA kernels stay the default; GVF_FPS=cluster selects the cluster kernel (identical indices).
  static int fps_mode = -1;
  if (fps_mode < 0) {
    const char* e = getenv("GVF_FPS");
    // "cluster" / "smem" (generation 1) / "pruned" (generation 3 for every call) / default: generation 2, and generation 3
    // through gvf_fps_ordered
    fps_mode = (e && e[0] == 'c') ? 0 : (e && e[0] == 's') ? 1 : (e && e[0] == 'p') ? 3 : 2;
  }
  auto launch_cluster = [&](auto kern) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(gvf::kFpsCluster);
    cfg.blockDim = dim3(gvf::kFpsThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = gvf::kFpsCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, pts, ld, P, K, start, (int*)out_idx) == cudaSuccess;
  };
  if (fps_mode == 0 && P > 1024 && P <= 16384) {
    const int per = (P + gvf::kFpsCluster * gvf::kFpsThreads - 1) / (gvf::kFpsCluster * gvf::kFpsThreads);
    ok = per <= 1 ? launch_cluster(gvf::fps_cluster_kernel<1>) : per <= 2 ? launch_cluster(gvf::fps_cluster_kernel<2>)
                                                                           : launch_cluster(gvf::fps_cluster_kernel<4>);
    return (ok && cudaGetLastError() == cudaSuccess) ? GVF_OK : GVF_ERR_CUDA;
  }
  if ((fps_mode == 3 || (fps_mode == 2 && ordered)) && P <= 16384) {
    static int half_block = -1;                    // GVF_FPS_THREADS=1024: the 1024-thread form at every size (A/B)
    if (half_block < 0) {
      const char* e2 = getenv("GVF_FPS_THREADS");
      half_block = (e2 && atoi(e2) == 1024) ? 0 : 1;
    }
    if (P <= 4096) ok = launch_smem(gvf::fps_pruned_kernel<4>, 4);
    else if (P <= 8192) ok = launch_smem(gvf::fps_pruned_kernel<8>, 8);
    else if (half_block) {
      const int bytes = 32 * 512 * 3 * (int)sizeof(float);
      auto kern = gvf::fps_pruned_kernel<32, 512>;
      ok = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess;
      if (ok) kern<<<1, 512, bytes, (cudaStream_t)stream>>>(pts, ld, P, K, start, out_idx);
    } else ok = launch_smem(gvf::fps_pruned_kernel<16>, 16);
    return (ok && cudaGetLastError() == cudaSuccess) ? GVF_OK : GVF_ERR_CUDA;
  }
  if (fps_mode == 2 && P <= 16384) {
    if (P <= 4096) ok = launch_smem(gvf::fps_smem2_kernel<4>, 4);
    else if (P <= 8192) ok = launch_smem(gvf::fps_smem2_kernel<8>, 8);
    else ok = launch_smem(gvf::fps_smem2_kernel<16>, 16);
    return (ok && cudaGetLastError() == cudaSuccess) ? GVF_OK : GVF_ERR_CUDA;
  }
  if (P <= 4096) ok = launch_smem(gvf::fps_smem_kernel<4>, 4);
  else if (P <= 8192) ok = launch_smem(gvf::fps_smem_kernel<8>, 8);
  else if (P <= 16384) ok = launch_smem(gvf::fps_smem_kernel<16>, 16);
  else gvf::fps_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(pts, ld, P, K, start, workspace, out_idx);
  return (ok && cudaGetLastError() == cudaSuccess) ? GVF_OK : GVF_ERR_CUDA;
}
