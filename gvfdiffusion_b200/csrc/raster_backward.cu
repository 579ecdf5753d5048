// raster_backward.cu -- backward of the canonical+delta Gaussian rasteriser.
//
// Replaces `GaussianRasterizer.backward` of diff_gaussian_rasterization (mip-splatting fork;
// reached through autograd from reference train_vae.py:313-352, renderers/gaussian_render.py:198-206)
// fused with the backward of GaussianModel.get_*_with_delta
// (representations/gaussian/gaussian_model.py:98-114), so gradients land directly on
// (delta_pc, _xyz, _features_dc, _scaling, _rotation, _opacity).  Batched over frames.
//
//   blend_backward   one CTA per (frame, tile): walks the tile's sorted list back to front,
//                    recomputes alpha, carries the suffix colour, reduces the nine per-splat
//                    gradients (pix xy, conic abc, opacity', rgb) over the warp with shuffles and
//                    issues one atomicAdd per warp and value (upstream: one per thread).
//   preprocess_backward  one thread per (frame, Gaussian): conic -> cov2D (+ mip opacity
//                    compensation) -> (Sigma, view-space mean) -> (scale, quaternion, mean) ->
//                    activations -> raw parameters and delta.
// Conventions that are a choice follow upstream: min(0.99, .) passes gradients through.
#include <stdlib.h>

#include "../../include/gvf_math.h"
#include "raster_common.h"

namespace gvf {

constexpr int kDs = 12;   // floats per (frame, Gaussian) in the dsplat accumulator (9 used)

struct BlendBwdArgs {
  int F, P, H, W, gx, gy;
  float bg0, bg1, bg2;
  const float4* splat;
  const uint32_t* tile_start;
  const uint32_t* point_list;
  long long cap;
  const float2* subpixel_offset;
  const float* final_T;
  const uint32_t* n_contrib;
  const float* dL_drgba;    // [F,4,H,W]
  float* dsplat;            // [F*P*kDs]
};

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(GVF_TILE_PIX) blend_backward_kernel(const BlendBwdArgs a) {
  __shared__ float4 sA[GVF_TILE_PIX];
  __shared__ float4 sB[GVF_TILE_PIX];
  __shared__ float sC[GVF_TILE_PIX];
  __shared__ uint32_t sId[GVF_TILE_PIX];
  const int T = a.gx * a.gy;
  const int tile = blockIdx.x;
  const int f = tile / T, t = tile - f * T;
  const int tyi = t / a.gx, txi = t - tyi * a.gx;
  const int tid = threadIdx.x, lane = tid & 31;
  long long s = a.tile_start[tile], e = a.tile_start[tile + 1];
  if (s > a.cap) s = a.cap;
  if (e > a.cap) e = a.cap;
  const int n = (int)(e - s);
  if (n == 0) return;
  const int px = txi * GVF_TILE + (tid & (GVF_TILE - 1));
  const int py = tyi * GVF_TILE + (tid >> 4);
  const bool inside = px < a.W && py < a.H;
  const size_t HW = (size_t)a.H * a.W;
  const size_t pid = (size_t)py * a.W + px;
  float pfx = (float)px, pfy = (float)py;
  float g0 = 0.f, g1 = 0.f, g2 = 0.f, Tc = 1.f, suffix = 0.f;
  int last = 0;
  if (inside) {
    if (a.subpixel_offset) {
      const float2 o = a.subpixel_offset[pid];
      pfx += o.x;
      pfy += o.y;
    }
    const float* g = a.dL_drgba + (size_t)f * 4 * HW + pid;
    g0 = g[0]; g1 = g[HW]; g2 = g[2 * HW];
    const float gA = g[3 * HW];
    Tc = a.final_T[(size_t)f * HW + pid];
    last = (int)a.n_contrib[(size_t)f * HW + pid];
    suffix = Tc * (a.bg0 * g0 + a.bg1 * g1 + a.bg2 * g2 - gA);   // d(C + T bg)/dT and A = 1 - T
  }
  const float4* sp = a.splat + (size_t)f * a.P * 3;
  float* ds = a.dsplat + (size_t)f * a.P * kDs;
  const int rounds = (n + GVF_TILE_PIX - 1) / GVF_TILE_PIX;
  for (int r = 0; r < rounds; ++r) {
    __syncthreads();
    const int pos = n - 1 - (r * GVF_TILE_PIX + tid);           // smem slot 0 = back of the list
    if (pos >= 0) {
      const uint32_t id = a.point_list[s + pos];
      const float4* rec = sp + (size_t)id * 3;
      sId[tid] = id;
      sA[tid] = __ldg(rec);
      sB[tid] = __ldg(rec + 1);
      sC[tid] = __ldg(reinterpret_cast<const float*>(rec + 2));
    }
    __syncthreads();
    const int m = min(GVF_TILE_PIX, n - r * GVF_TILE_PIX);
    for (int k = 0; k < m; ++k) {
      const int posk = n - 1 - (r * GVF_TILE_PIX + k);
      float v[9];
#pragma unroll
      for (int j = 0; j < 9; ++j) v[j] = 0.f;
      bool hit = false;
      if (inside && posk < last) {
        const float4 A = sA[k], B = sB[k];
        const float dx = A.x - pfx, dy = A.y - pfy;
        const float power = -0.5f * (A.z * dx * dx + B.x * dy * dy) - A.w * dx * dy;
        if (power <= 0.0f) {
          const float G = __expf(power);
          const float alpha = fminf(0.99f, B.y * G);
          if (alpha >= 1.0f / 255.0f) {
            hit = true;
            const float one_m = 1.0f - alpha;
            Tc = Tc / one_m;
            const float w = alpha * Tc;
            const float cb = sC[k];
            const float cdot = B.z * g0 + B.w * g1 + cb * g2;
            const float dL_dalpha = Tc * cdot - suffix / one_m;
            suffix += cdot * w;
            const float dpow = G * B.y * dL_dalpha;               // dL/dpower
            v[0] = dpow * (-A.z * dx - A.w * dy);                 // d/d px
            v[1] = dpow * (-B.x * dy - A.w * dx);                 // d/d py
            v[2] = dpow * (-0.5f * dx * dx);                      // d/d conic a
            v[3] = dpow * (-dx * dy);                             // d/d conic b
            v[4] = dpow * (-0.5f * dy * dy);                      // d/d conic c
            v[5] = G * dL_dalpha;                                 // d/d opacity'
            v[6] = w * g0; v[7] = w * g1; v[8] = w * g2;          // d/d rgb
          }
        }
      }
      if (__any_sync(0xffffffffu, hit)) {
#pragma unroll
        for (int j = 0; j < 9; ++j) v[j] = warp_sum_f(v[j]);
        if (lane == 0) {
          float* d = ds + (size_t)sId[k] * kDs;
#pragma unroll
          for (int j = 0; j < 9; ++j) atomicAdd(d + j, v[j]);
        }
      }
    }
  }
}

// Generation 2 of the blend backward (default; GVF_RASTER_BWD=v1 selects the kernel above).  The launch list
// showed generation 1 at 1.35 ms per 24 frames against 0.30 ms for the forward's sort_blend: it evaluated every
// (pixel, splat) pair of a tile (86 % of the (warp, splat) pairs are rejected by all 32 lanes) and paid nine
// five-step shuffle reductions plus nine atomics per surviving pair.  Here
//   * the forward's staging is reused verbatim -- conic in the log2 domain, opacity-aware sub-tile masks, 8 x 4
//     pixel warp footprint, MUFU.EX2 -- so a warp walks only the splats whose ellipse reaches its block and
//     takes bit-identical skip decisions to the forward pass (same p2, same alpha);
//   * list entries behind the last contributor of every pixel of the tile are not staged at all;
//   * the nine per-splat sums are reduced with a transposing butterfly (eight values in 9 shuffles instead of
//     40; the ninth in 5) and leave the warp as ONE atomic instruction: lane 4 j holds value j, lane 1 value 8,
//     36 contiguous bytes of the dsplat row.
__device__ __forceinline__ float ex2_approx_b(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(GVF_TILE_PIX) blend_backward2_kernel(const BlendBwdArgs a) {
  __shared__ float4 sA[GVF_TILE_PIX];    // px, py, -qa, -qb          (q = conic scaled to the log2 domain)
  __shared__ float2 sT[GVF_TILE_PIX];    // -qc, -tau
  __shared__ float4 sC[GVF_TILE_PIX];    // opacity', r, g, b
  __shared__ uint32_t sM[GVF_TILE_PIX];  // warp-block mask
  __shared__ uint32_t sId[GVF_TILE_PIX];
  __shared__ int sLast[8];
  const int T = a.gx * a.gy;
  const int tile = blockIdx.x;
  const int f = tile / T, t = tile - f * T;
  const int tyi = t / a.gx, txi = t - tyi * a.gx;
  const int tid = threadIdx.x, ln = tid & 31, wq = tid >> 5;
  long long s = a.tile_start[tile], e = a.tile_start[tile + 1];
  if (s > a.cap) s = a.cap;
  if (e > a.cap) e = a.cap;
  if (e == s) return;
  const int px = txi * GVF_TILE + (ln & 7) + 8 * (wq & 1);
  const int py = tyi * GVF_TILE + (ln >> 3) + 4 * (wq >> 1);
  const bool inside = px < a.W && py < a.H;
  const size_t HW = (size_t)a.H * a.W;
  const size_t pid = (size_t)py * a.W + px;
  float pfx = (float)px, pfy = (float)py;
  float g0 = 0.f, g1 = 0.f, g2 = 0.f, Tc = 1.f, suffix = 0.f;
  int last = 0;
  if (inside) {
    if (a.subpixel_offset) {
      const float2 o = a.subpixel_offset[pid];
      pfx += o.x;
      pfy += o.y;
    }
    const float* g = a.dL_drgba + (size_t)f * 4 * HW + pid;
    g0 = g[0]; g1 = g[HW]; g2 = g[2 * HW];
    const float gA = g[3 * HW];
    Tc = a.final_T[(size_t)f * HW + pid];
    last = (int)a.n_contrib[(size_t)f * HW + pid];
    suffix = Tc * (a.bg0 * g0 + a.bg1 * g1 + a.bg2 * g2 - gA);   // d(C + T bg)/dT and A = 1 - T
  }
  // entries at or behind n = max over the tile of `last` contribute to no pixel
  {
    const int wl = __reduce_max_sync(0xffffffffu, last);
    if (ln == 0) sLast[wq] = wl;
    __syncthreads();
    int m8 = sLast[ln & 7];
    m8 = __reduce_max_sync(0xffffffffu, m8);
    last = inside ? last : 0;
    if (m8 == 0) return;
    e = s + m8;
  }
  const int n = (int)(e - s);
  const float4* sp = a.splat + (size_t)f * a.P * 3;
  float* ds = a.dsplat + (size_t)f * a.P * kDs;
  const float tile_x0 = (float)(txi * GVF_TILE), tile_y0 = (float)(tyi * GVF_TILE);
  const float slack = a.subpixel_offset ? 1.01f : 0.01f;
  constexpr float kLn2 = 0.6931471805599453f;
  const int rounds = (n + GVF_TILE_PIX - 1) / GVF_TILE_PIX;
  for (int r = 0; r < rounds; ++r) {
    __syncthreads();
    const int pos = n - 1 - (r * GVF_TILE_PIX + tid);           // smem slot 0 = back of the list
    if (pos >= 0) {
      const uint32_t id = a.point_list[s + pos];
      const float4* rec = sp + (size_t)id * 3;
      const float4 r0 = __ldg(rec), r1 = __ldg(rec + 1);
      // identical to the staging of sort_blend_kernel (raster_blend.cu): same values, same decisions
      constexpr float kLog2e = 1.4426950408889634f;
      const float qa = 0.5f * kLog2e * r0.z, qb = kLog2e * r0.w, qc = 0.5f * kLog2e * r1.x;
      const float tau = __log2f(255.0f * r1.y) + 0.0145f;
      sId[tid] = id;
      sA[tid] = make_float4(r0.x, r0.y, -qa, -qb);
      sT[tid] = make_float2(-qc, -tau);
      sC[tid] = make_float4(r1.y, r1.z, r1.w, __ldg(reinterpret_cast<const float*>(rec + 2)));
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
    const int m = min(GVF_TILE_PIX, n - r * GVF_TILE_PIX);
    for (int g = 0; g < m; g += 32) {
      const uint32_t mk = (g + ln < m) ? sM[g + ln] : 0u;
      uint32_t bits = __ballot_sync(0xffffffffu, (mk >> wq) & 1u);
      while (bits) {
        const int k = g + __ffs(bits) - 1;
        bits &= bits - 1;
        const int posk = n - 1 - (r * GVF_TILE_PIX + k);
        float v[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) v[j] = 0.f;
        bool hit = false;
        if (posk < last) {
          const float4 A = sA[k];
          const float2 ct = sT[k];
          const float dx = A.x - pfx, dy = A.y - pfy;
          const float tt = fmaf(A.z, dx, A.w * dy);
          const float p2 = fmaf(dx, tt, (ct.x * dy) * dy);       // log2 of exp(power), as in the forward
          if (!(p2 > 0.0f || p2 < ct.y)) {
            const float4 B = sC[k];
            const float G = ex2_approx_b(p2);
            const float alpha = fminf(0.99f, B.x * G);
            if (alpha >= 1.0f / 255.0f) {
              hit = true;
              const float one_m = 1.0f - alpha;
              Tc = Tc / one_m;
              const float w = alpha * Tc;
              const float cdot = B.y * g0 + B.z * g1 + B.w * g2;
              const float dL_dalpha = Tc * cdot - suffix / one_m;
              suffix += cdot * w;
              const float dpow = G * B.x * dL_dalpha;             // dL/dpower, power = ln 2 * p2
              const float dl2 = dpow * kLn2;                      // conic = 2 ln2 qa, ln2 qb, 2 ln2 qc
              v[0] = dl2 * (2.0f * A.z * dx + A.w * dy);          // d/d px: dpow * (-a dx - b dy)
              v[1] = dl2 * (2.0f * ct.x * dy + A.w * dx);         // d/d py: dpow * (-c dy - b dx)
              v[2] = dpow * (-0.5f * dx * dx);                    // d/d conic a
              v[3] = dpow * (-dx * dy);                           // d/d conic b
              v[4] = dpow * (-0.5f * dy * dy);                    // d/d conic c
              v[5] = G * dL_dalpha;                               // d/d opacity'
              v[6] = w * g0; v[7] = w * g1; v[8] = w * g2;        // d/d rgb
            }
          }
        }
        if (__any_sync(0xffffffffu, hit)) {
          // transposing butterfly: after the offsets 16, 8, 4 a lane keeps one of the eight values
          // (index (ln >> 2) & 7), summed over the 8 lanes that differ in lane bits 4..2; offsets 2, 1 finish
          float u4[4], u2[2], u1;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const bool hi = ln & 16;
            const float keep = hi ? v[4 + j] : v[j], give = hi ? v[j] : v[4 + j];
            u4[j] = keep + __shfl_xor_sync(0xffffffffu, give, 16);
          }
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const bool hi = ln & 8;
            const float keep = hi ? u4[2 + j] : u4[j], give = hi ? u4[j] : u4[2 + j];
            u2[j] = keep + __shfl_xor_sync(0xffffffffu, give, 8);
          }
          {
            const bool hi = ln & 4;
            const float keep = hi ? u2[1] : u2[0], give = hi ? u2[0] : u2[1];
            u1 = keep + __shfl_xor_sync(0xffffffffu, give, 4);
          }
          u1 += __shfl_xor_sync(0xffffffffu, u1, 2);
          u1 += __shfl_xor_sync(0xffffffffu, u1, 1);
          const float v8 = warp_sum_f(v[8]);
          float* d = ds + (size_t)sId[k] * kDs;
          if ((ln & 3) == 0) atomicAdd(d + (ln >> 2), u1);
          else if (ln == 1) atomicAdd(d + 8, v8);
        }
      }
    }
  }
}

struct PreBwdArgs {
  gvf_raster_params prm;
  int F, P, activated;
  const float *xyz, *dc, *scaling, *rotation, *opacity, *delta, *cams;
  const float4* splat;
  const float* dsplat;
  float *g_xyz, *g_dc, *g_scaling, *g_rotation, *g_opacity;   // raw: [P,..] (atomic over frames); activated: [F,P,..]
  float* g_delta;      // [F,P,14] or null
  float* g_means2D;    // [F,P,2] or null
};

__global__ void __launch_bounds__(256) preprocess_backward_kernel(const PreBwdArgs a) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)a.F * a.P) return;
  const int f = (int)(gid / a.P);
  const int i = (int)(gid - (long long)f * a.P);
  const gvf_raster_params& prm = a.prm;
  const int W = prm.W, H = prm.H;
  const float4 rec2 = a.splat[gid * 3 + 2];
  const bool visible = __float_as_int(rec2.z) > 0;              // radius
  float gd[14];
#pragma unroll
  for (int k = 0; k < 14; ++k) gd[k] = 0.f;
  float gm2x = 0.f, gm2y = 0.f;
  // outputs wrt activated inputs
  float g_p[3] = {0, 0, 0}, g_sc[3] = {0, 0, 0}, g_q[4] = {0, 0, 0, 0}, g_sh[3] = {0, 0, 0}, g_opac = 0.f;
  // forward activation state (needed for the activation backward)
  float m3[3], sc[3], q[4], sh[3], opac, sp_raw[3] = {0, 0, 0}, vn = 1.f, vraw[4] = {1, 0, 0, 0};
  if (a.activated) {
    const size_t b = (size_t)gid;
    for (int c = 0; c < 3; ++c) { m3[c] = a.xyz[b * 3 + c]; sh[c] = a.dc[b * 3 + c]; sc[c] = a.scaling[b * 3 + c]; }
    for (int c = 0; c < 4; ++c) q[c] = a.rotation[b * 4 + c];
    opac = a.opacity[b];
  } else {
    float d[14];
#pragma unroll
    for (int k = 0; k < 14; ++k) d[k] = a.delta ? a.delta[(size_t)gid * 14 + k] : 0.f;
    const float k2 = prm.min_kernel * prm.min_kernel;
    for (int c = 0; c < 3; ++c) m3[c] = a.xyz[(size_t)i * 3 + c] * prm.aabb[3 + c] + prm.aabb[c] + d[c];
    for (int c = 0; c < 3; ++c) {
      const float u = a.scaling[(size_t)i * 3 + c] + prm.scale_bias + d[3 + c];
      sp_raw[c] = u;
      const float s = prm.softplus ? gvf_softplusf(u) : gvf_expf(u);
      sc[c] = sqrtf(s * s + k2);
    }
    for (int c = 0; c < 4; ++c) vraw[c] = a.rotation[(size_t)i * 4 + c] + (c == 0 ? 1.0f : 0.0f) + d[6 + c];
    vn = fmaxf(sqrtf(vraw[0] * vraw[0] + vraw[1] * vraw[1] + vraw[2] * vraw[2] + vraw[3] * vraw[3]), 1e-12f);
    for (int c = 0; c < 4; ++c) q[c] = vraw[c] / vn;
    for (int c = 0; c < 3; ++c) sh[c] = a.dc[(size_t)i * 3 + c] + d[10 + c];
    opac = gvf_sigmoidf(a.opacity[i] + prm.opacity_bias + d[13]);
  }

  if (visible) {
    const float* ds = a.dsplat + (size_t)gid * kDs;
    const float g_px = ds[0], g_py = ds[1], gA = ds[2], gB = ds[3], gC = ds[4], g_opp = ds[5];
    const float* view = a.cams + (size_t)f * 32;
    const float* proj = view + 16;
    const float px = m3[0], py = m3[1], pz = m3[2];
    const float tx = view[0] * px + view[4] * py + view[8] * pz + view[12];
    const float ty = view[1] * px + view[5] * py + view[9] * pz + view[13];
    const float tz = view[2] * px + view[6] * py + view[10] * pz + view[14];
    const float hx = proj[0] * px + proj[4] * py + proj[8] * pz + proj[12];
    const float hy = proj[1] * px + proj[5] * py + proj[9] * pz + proj[13];
    const float hw = proj[3] * px + proj[7] * py + proj[11] * pz + proj[15];
    const float pw = 1.0f / (hw + 0.0000001f);
    // ---- colour, opacity
    const float SH_C0 = 0.28209479177387814f;
    for (int c = 0; c < 3; ++c) g_sh[c] = (SH_C0 * sh[c] + 0.5f > 0.0f) ? SH_C0 * ds[6 + c] : 0.f;
    // ---- forward intermediates
    const float s0 = prm.scale_modifier * sc[0], s1 = prm.scale_modifier * sc[1], s2 = prm.scale_modifier * sc[2];
    const float r = q[0], x = q[1], y = q[2], z = q[3];
    const float R[9] = {1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
                        2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
                        2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y)};
    const float sv[3] = {s0, s1, s2};
    float L[9], S[9];
    for (int ii = 0; ii < 3; ++ii) for (int kk = 0; kk < 3; ++kk) L[ii * 3 + kk] = R[ii * 3 + kk] * sv[kk];
    for (int ii = 0; ii < 3; ++ii) for (int jj = 0; jj < 3; ++jj)
      S[ii * 3 + jj] = L[ii * 3] * L[jj * 3] + L[ii * 3 + 1] * L[jj * 3 + 1] + L[ii * 3 + 2] * L[jj * 3 + 2];
    const float fx = (float)W / (2.0f * prm.tanfovx), fy = (float)H / (2.0f * prm.tanfovy);
    const float limx = 1.3f * prm.tanfovx, limy = 1.3f * prm.tanfovy;
    const float txtz = tx / tz, tytz = ty / tz;
    const bool clx = txtz < -limx || txtz > limx, cly = tytz < -limy || tytz > limy;
    const float cxn = fminf(limx, fmaxf(-limx, txtz)), cyn = fminf(limy, fmaxf(-limy, tytz));
    const float cx = cxn * tz, cy = cyn * tz;
    const float J00 = fx / tz, J02 = -(fx * cx) / (tz * tz), J11 = fy / tz, J12 = -(fy * cy) / (tz * tz);
    const float Wv[9] = {view[0], view[4], view[8], view[1], view[5], view[9], view[2], view[6], view[10]};
    float Tm[6];
    for (int jj = 0; jj < 3; ++jj) {
      Tm[jj] = J00 * Wv[jj] + J02 * Wv[6 + jj];
      Tm[3 + jj] = J11 * Wv[3 + jj] + J12 * Wv[6 + jj];
    }
    float TS[6];   // Tm * S
    for (int ii = 0; ii < 2; ++ii) for (int jj = 0; jj < 3; ++jj)
      TS[ii * 3 + jj] = Tm[ii * 3] * S[jj] + Tm[ii * 3 + 1] * S[3 + jj] + Tm[ii * 3 + 2] * S[6 + jj];
    const float a0 = TS[0] * Tm[0] + TS[1] * Tm[1] + TS[2] * Tm[2];
    const float b = TS[0] * Tm[3] + TS[1] * Tm[4] + TS[2] * Tm[5];
    const float c0 = TS[3] * Tm[3] + TS[4] * Tm[4] + TS[5] * Tm[5];
    const float ks = prm.kernel_size;
    const float d0r = a0 * c0 - b * b, d1r = (a0 + ks) * (c0 + ks) - b * b;
    const float det0 = fmaxf(1e-6f, d0r), det1 = fmaxf(1e-6f, d1r);
    const bool mip = prm.mip_filter != 0;
    const bool coef_zero = mip && (det0 <= 1e-6f || det1 <= 1e-6f);
    const float coef = !mip ? 1.0f : (coef_zero ? 0.f : sqrtf(det0 / (det1 + 1e-6f) + 1e-6f));
    const float ca = a0 + ks, cc = c0 + ks;
    const float det = ca * cc - b * b;
    const float di2 = 1.0f / (det * det);
    // ---- conic -> filtered cov (a, b, c)
    float g_a = (gA * (-cc * cc) + gB * (b * cc) + gC * (-b * b)) * di2;
    float g_b = (gA * (2.f * b * cc) + gB * (-(ca * cc + b * b)) + gC * (2.f * ca * b)) * di2;
    float g_c = (gA * (-b * b) + gB * (ca * b) + gC * (-ca * ca)) * di2;
    // ---- mip opacity compensation
    g_opac = coef * g_opp;
    if (mip && !coef_zero) {
      const float g_coef = opac * g_opp;
      const float g_r = g_coef / (2.0f * coef);
      const float g_det0 = (d0r > 1e-6f) ? g_r / (det1 + 1e-6f) : 0.f;
      const float g_det1 = (d1r > 1e-6f) ? -g_r * det0 / ((det1 + 1e-6f) * (det1 + 1e-6f)) : 0.f;
      g_a += g_det0 * c0 + g_det1 * (c0 + ks);
      g_c += g_det0 * a0 + g_det1 * (a0 + ks);
      g_b += -2.f * b * (g_det0 + g_det1);
    }
    // ---- cov2D = Tm S Tm^T : dL/dS = Tm^T Gc Tm ; dL/dTm = 2 Gc Tm S
    const float Gc[4] = {g_a, 0.5f * g_b, 0.5f * g_b, g_c};
    float gS[9];
    for (int ii = 0; ii < 3; ++ii) for (int jj = 0; jj < 3; ++jj)
      gS[ii * 3 + jj] = Tm[ii] * (Gc[0] * Tm[jj] + Gc[1] * Tm[3 + jj]) + Tm[3 + ii] * (Gc[2] * Tm[jj] + Gc[3] * Tm[3 + jj]);
    float gT[6];
    for (int jj = 0; jj < 3; ++jj) {
      gT[jj] = 2.f * (Gc[0] * TS[jj] + Gc[1] * TS[3 + jj]);
      gT[3 + jj] = 2.f * (Gc[2] * TS[jj] + Gc[3] * TS[3 + jj]);
    }
    // ---- Tm = J Wv : dL/dJ = gT Wv^T
    const float gJ00 = gT[0] * Wv[0] + gT[1] * Wv[1] + gT[2] * Wv[2];
    const float gJ02 = gT[0] * Wv[6] + gT[1] * Wv[7] + gT[2] * Wv[8];
    const float gJ11 = gT[3] * Wv[3] + gT[4] * Wv[4] + gT[5] * Wv[5];
    const float gJ12 = gT[3] * Wv[6] + gT[4] * Wv[7] + gT[5] * Wv[8];
    const float tz2 = tz * tz, tz3 = tz2 * tz;
    float g_tx = 0.f, g_ty = 0.f, g_tz = -fx / tz2 * gJ00 - fy / tz2 * gJ11;
    if (!clx) { g_tx += -fx / tz2 * gJ02; g_tz += 2.f * fx * tx / tz3 * gJ02; }
    else g_tz += fx * cxn / tz2 * gJ02;                         // J02 = -fx * lim / tz
    if (!cly) { g_ty += -fy / tz2 * gJ12; g_tz += 2.f * fy * ty / tz3 * gJ12; }
    else g_tz += fy * cyn / tz2 * gJ12;
    for (int jj = 0; jj < 3; ++jj) g_p[jj] = Wv[jj] * g_tx + Wv[3 + jj] * g_ty + Wv[6 + jj] * g_tz;
    // ---- pixel position
    const float g_ndcx = g_px * 0.5f * (float)W, g_ndcy = g_py * 0.5f * (float)H;
    gm2x = g_ndcx; gm2y = g_ndcy;
    const float g_hx = g_ndcx * pw, g_hy = g_ndcy * pw, g_hw = -(g_ndcx * hx + g_ndcy * hy) * pw * pw;
    for (int jj = 0; jj < 3; ++jj) g_p[jj] += proj[4 * jj] * g_hx + proj[4 * jj + 1] * g_hy + proj[4 * jj + 3] * g_hw;
    // ---- S = L L^T, L = R diag(s)
    float gL[9];
    for (int ii = 0; ii < 3; ++ii) for (int kk = 0; kk < 3; ++kk)
      gL[ii * 3 + kk] = 2.f * (gS[ii * 3] * L[kk] + gS[ii * 3 + 1] * L[3 + kk] + gS[ii * 3 + 2] * L[6 + kk]);
    float gR[9];
    for (int kk = 0; kk < 3; ++kk) {
      g_sc[kk] = prm.scale_modifier * (gL[kk] * R[kk] + gL[3 + kk] * R[3 + kk] + gL[6 + kk] * R[6 + kk]);
      for (int ii = 0; ii < 3; ++ii) gR[ii * 3 + kk] = gL[ii * 3 + kk] * sv[kk];
    }
    g_q[0] = 2.f * (-z * gR[1] + y * gR[2] + z * gR[3] - x * gR[5] - y * gR[6] + x * gR[7]);
    g_q[1] = 2.f * (y * gR[1] + z * gR[2] + y * gR[3] - 2.f * x * gR[4] - r * gR[5] + z * gR[6] + r * gR[7] - 2.f * x * gR[8]);
    g_q[2] = 2.f * (-2.f * y * gR[0] + x * gR[1] + r * gR[2] + x * gR[3] + z * gR[5] - r * gR[6] + z * gR[7] - 2.f * y * gR[8]);
    g_q[3] = 2.f * (-2.f * z * gR[0] - r * gR[1] + x * gR[2] + r * gR[3] - 2.f * z * gR[4] + y * gR[5] + x * gR[6] + y * gR[7]);
  }

  if (a.g_means2D) { a.g_means2D[gid * 2] = gm2x; a.g_means2D[gid * 2 + 1] = gm2y; }
  if (a.activated) {
    const size_t b = (size_t)gid;
    for (int c = 0; c < 3; ++c) { a.g_xyz[b * 3 + c] = g_p[c]; a.g_dc[b * 3 + c] = g_sh[c]; a.g_scaling[b * 3 + c] = g_sc[c]; }
    for (int c = 0; c < 4; ++c) a.g_rotation[b * 4 + c] = g_q[c];
    a.g_opacity[b] = g_opac;
    return;
  }
  // ---- activation backward (GaussianModel.get_*_with_delta)
  if (visible) {
    for (int c = 0; c < 3; ++c) gd[c] = g_p[c];
    for (int c = 0; c < 3; ++c) {
      const float u = sp_raw[c];
      const float sgm = gvf_sigmoidf(u);
      const float spl = prm.softplus ? gvf_softplusf(u) : gvf_expf(u);
      const float dact = prm.softplus ? ((u > 20.0f) ? 1.0f : sgm) : spl;    // d softplus = sigmoid ; d exp = exp
      gd[3 + c] = g_sc[c] * spl / sc[c] * dact;                               // scale = sqrt(act^2 + k^2)
    }
    float dot = 0.f;
    for (int c = 0; c < 4; ++c) dot += q[c] * g_q[c];
    for (int c = 0; c < 4; ++c) gd[6 + c] = (g_q[c] - q[c] * dot) / vn;      // normalize backward
    for (int c = 0; c < 3; ++c) gd[10 + c] = g_sh[c];
    gd[13] = g_opac * opac * (1.0f - opac);
    for (int c = 0; c < 3; ++c) atomicAdd(a.g_xyz + (size_t)i * 3 + c, gd[c] * prm.aabb[3 + c]);
    for (int c = 0; c < 3; ++c) atomicAdd(a.g_scaling + (size_t)i * 3 + c, gd[3 + c]);
    for (int c = 0; c < 4; ++c) atomicAdd(a.g_rotation + (size_t)i * 4 + c, gd[6 + c]);
    for (int c = 0; c < 3; ++c) atomicAdd(a.g_dc + (size_t)i * 3 + c, gd[10 + c]);
    atomicAdd(a.g_opacity + i, gd[13]);
  }
  if (a.g_delta) {
    float* o = a.g_delta + (size_t)gid * 14;
#pragma unroll
    for (int k = 0; k < 14; ++k) o[k] = gd[k];
  }
}

}  // namespace gvf

using namespace gvf;

extern "C" GVF_API int gvf_raster_backward(const gvf_raster_params* prm, int F, int P, int activated,
                                           const float* xyz, const float* dc, const float* scaling,
                                           const float* rotation, const float* opacity, const float* delta,
                                           const float* cams, const float* subpixel_offset,
                                           const float* dL_drgba, void* workspace, size_t workspace_bytes,
                                           int64_t cap, float* g_xyz, float* g_dc, float* g_scaling,
                                           float* g_rotation, float* g_opacity, float* g_delta,
                                           float* g_means2D, void* stream) {
  if (!prm || !xyz || !dc || !scaling || !rotation || !opacity || !cams || !dL_drgba || !workspace ||
      !g_xyz || !g_dc || !g_scaling || !g_rotation || !g_opacity)
    return GVF_ERR_INVALID;
  if (F <= 0 || P <= 0 || cap <= 0 || (activated && delta)) return GVF_ERR_INVALID;
  const RasterLayout L = raster_layout(F, P, prm->H, prm->W, cap);
  if (workspace_bytes < L.total) return GVF_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  char* w = (char*)workspace;
  const int gx = (prm->W + GVF_TILE - 1) / GVF_TILE, gy = (prm->H + GVF_TILE - 1) / GVF_TILE;
  float* dsplat = (float*)(w + L.off[GVF_RB_DSPLAT]);
  if (cudaMemsetAsync(dsplat, 0, (size_t)F * P * kDs * sizeof(float), st) != cudaSuccess) return GVF_ERR_CUDA;
  if (!activated) {
    cudaMemsetAsync(g_xyz, 0, (size_t)P * 3 * 4, st);
    cudaMemsetAsync(g_dc, 0, (size_t)P * 3 * 4, st);
    cudaMemsetAsync(g_scaling, 0, (size_t)P * 3 * 4, st);
    cudaMemsetAsync(g_rotation, 0, (size_t)P * 4 * 4, st);
    if (cudaMemsetAsync(g_opacity, 0, (size_t)P * 4, st) != cudaSuccess) return GVF_ERR_CUDA;
  }
  BlendBwdArgs b;
  b.F = F; b.P = P; b.H = prm->H; b.W = prm->W; b.gx = gx; b.gy = gy;
  b.bg0 = prm->bg[0]; b.bg1 = prm->bg[1]; b.bg2 = prm->bg[2];
  b.splat = (const float4*)(w + L.off[GVF_RB_SPLAT]);
  b.tile_start = (const uint32_t*)(w + L.off[GVF_RB_TILE_START]);
  b.point_list = (const uint32_t*)(w + L.off[GVF_RB_POINT_LIST]);
  b.cap = cap;
  b.subpixel_offset = (const float2*)subpixel_offset;
  b.final_T = (const float*)(w + L.off[GVF_RB_FINAL_T]);
  b.n_contrib = (const uint32_t*)(w + L.off[GVF_RB_N_CONTRIB]);
  b.dL_drgba = dL_drgba;
  b.dsplat = dsplat;
  static int bwd_gen = -1;
  if (bwd_gen < 0) {
    const char* ev = getenv("GVF_RASTER_BWD");
    bwd_gen = (ev && ev[0] == 'v' && ev[1] == '1') ? 1 : 2;
  }
  if (bwd_gen == 1) blend_backward_kernel<<<(unsigned)((size_t)F * gx * gy), GVF_TILE_PIX, 0, st>>>(b);
  else blend_backward2_kernel<<<(unsigned)((size_t)F * gx * gy), GVF_TILE_PIX, 0, st>>>(b);
  if (cudaGetLastError() != cudaSuccess) return GVF_ERR_CUDA;
  PreBwdArgs p;
  p.prm = *prm; p.F = F; p.P = P; p.activated = activated;
  p.xyz = xyz; p.dc = dc; p.scaling = scaling; p.rotation = rotation; p.opacity = opacity;
  p.delta = delta; p.cams = cams; p.splat = b.splat; p.dsplat = dsplat;
  p.g_xyz = g_xyz; p.g_dc = g_dc; p.g_scaling = g_scaling; p.g_rotation = g_rotation; p.g_opacity = g_opacity;
  p.g_delta = g_delta; p.g_means2D = g_means2D;
  const long long n = (long long)F * P;
  preprocess_backward_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p);
  return cudaGetLastError() == cudaSuccess ? GVF_OK : GVF_ERR_CUDA;
}
