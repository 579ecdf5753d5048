// sparse_vae.cu -- the voxel-side operators around the static (canonical Gaussian) VAE:
//
//  (1) gvf_to_representation: SparseVAE.to_representation (reference
//      model/sparse_voxel_diffusion/sparse_vae.py:114-180) -- the decoder's [Nvox, 14 G] feature rows
//      become the raw GaussianModel tensors (_xyz, _features_dc, _scaling, _rotation, _opacity) of
//      P = Nvox * G Gaussians in one launch (the reference runs ~25 slicing / elementwise launches per
//      batch entry and builds one python object per entry).
//  (2) submanifold sparse 3-D convolution (reference sparse/conv/conv_spconv.py:6-15 ->
//      spconv.SubMConv3d, used by trellis/models/structured_latent_flow.py:34-35 and
//      structured_latent_vae/decoder_mesh.py:43-52): neighbour map through a dense voxel -> row
//      index grid (180 GB of HBM make the [B, D, D, D] int32 grid the cheapest exact "hash"), an
//      im2col gather into the fp16 [N, k^3 Cin] operand of the tcgen05 GEMM (gvf_gemm_f16, which
//      applies bias / residual epilogues), and nothing else: out[i] = sum_k W[:, k, :] x[nbr(i, k)].
//
// HBM-bound byte movers: coalesced 16 B accesses, one pass over the data, no tensor cores here.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gvf_b200.h"

namespace gvf {

// one thread per (voxel, gaussian); a warp covers 32 / G voxels whose 14 G-float rows it reads whole
__global__ void __launch_bounds__(256) to_representation_kernel(
    const float* __restrict__ feats, int ldf, const int* __restrict__ coords, int nvox, int G,
    const float* __restrict__ perturb, float lr_xyz, float lr_dc, float lr_scaling, float lr_rotation,
    float lr_opacity, float resolution, int reg_mode, float soft_scale, float* __restrict__ xyz,
    float* __restrict__ dc, float* __restrict__ scaling, float* __restrict__ rotation,
    float* __restrict__ opacity) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= (long long)nvox * G) return;
  const int v = (int)(p / G), g = (int)(p % G);
  const float* row = feats + (size_t)v * ldf;
  const int* c = coords + (size_t)v * 4;
  // layout of a row (sparse_vae.py:211-227): _xyz (G,3) | _features_dc (G,1,3) | _scaling (G,3) |
  // _rotation (G,4) | _opacity (G,1)
  const float* fx = row + 3 * g;
  const float* fd = row + 3 * G + 3 * g;
  const float* fs = row + 6 * G + 3 * g;
  const float* fr = row + 9 * G + 4 * g;
  const float* fo = row + 13 * G + g;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    // xyz = (coords + 0.5) / resolution (:143); offset = feats * lr (+ perturbation) (:146-149)
    const float centre = __fdiv_rn(__fadd_rn((float)c[1 + a], 0.5f), resolution);
    float off = __fmul_rn(fx[a], lr_xyz);
    if (perturb) off = __fadd_rn(off, perturb[3 * g + a]);
    if (reg_mode == 1) off = __fdiv_rn(tanhf(off), resolution);                       // invoxel (:150-151)
    else if (reg_mode == 2)                                                           // soft_invoxel (:152-153)
      off = __fmul_rn(__fmul_rn(__fdiv_rn(tanhf(off), resolution), 0.5f), soft_scale);
    xyz[p * 3 + a] = __fadd_rn(centre, off);
    dc[p * 3 + a] = __fmul_rn(fd[a], lr_dc);
    scaling[p * 3 + a] = __fmul_rn(fs[a], lr_scaling);
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) rotation[p * 4 + a] = __fmul_rn(fr[a], lr_rotation);
  opacity[p] = __fmul_rn(fo[0], lr_opacity);
}

// backward of the above: gradients of the raw GaussianModel tensors -> gradient of the feature rows [nvox, ldf]
// (columns >= 14 G are zeroed).  d offset / d feat = lr_xyz (1 - tanh^2) * {1, 1/res, 0.5 soft_scale / res}.
__global__ void __launch_bounds__(256) to_representation_bwd_kernel(
    const float* __restrict__ feats, int ldf, int nvox, int G, const float* __restrict__ perturb, float lr_xyz,
    float lr_dc, float lr_scaling, float lr_rotation, float lr_opacity, float resolution, int reg_mode, float soft_scale,
    const float* __restrict__ g_xyz, const float* __restrict__ g_dc, const float* __restrict__ g_scaling,
    const float* __restrict__ g_rotation, const float* __restrict__ g_opacity, float* __restrict__ g_feats) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= (long long)nvox * G) return;
  const int v = (int)(p / G), g = (int)(p % G);
  const float* row = feats + (size_t)v * ldf;
  float* out = g_feats + (size_t)v * ldf;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float off = row[3 * g + a] * lr_xyz;
    if (perturb) off += perturb[3 * g + a];
    float d = lr_xyz;
    if (reg_mode != 0) {
      const float t = tanhf(off);
      d *= (1.0f - t * t) / resolution * (reg_mode == 2 ? 0.5f * soft_scale : 1.0f);
    }
    out[3 * g + a] = g_xyz ? g_xyz[p * 3 + a] * d : 0.f;
    out[3 * G + 3 * g + a] = g_dc ? g_dc[p * 3 + a] * lr_dc : 0.f;
    out[6 * G + 3 * g + a] = g_scaling ? g_scaling[p * 3 + a] * lr_scaling : 0.f;
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) out[9 * G + 4 * g + a] = g_rotation ? g_rotation[p * 4 + a] * lr_rotation : 0.f;
  out[13 * G + g] = g_opacity ? g_opacity[p] * lr_opacity : 0.f;
  if (g == 0)
    for (int c = 14 * G; c < ldf; ++c) out[c] = 0.f;
}

// ------------------------------------------------------------------------------------------------
// submanifold convolution: neighbour map
__global__ void __launch_bounds__(256) voxel_grid_fill_kernel(const int* __restrict__ coords, int n, int B, int D,
                                                              int* __restrict__ grid, int* __restrict__ err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int4 c = reinterpret_cast<const int4*>(coords)[i];
  if ((unsigned)c.x >= (unsigned)B || (unsigned)c.y >= (unsigned)D || (unsigned)c.z >= (unsigned)D ||
      (unsigned)c.w >= (unsigned)D) {
    atomicOr(err, 1);
    return;
  }
  const size_t cell = (((size_t)c.x * D + c.y) * D + c.z) * D + c.w;
  // duplicate coordinates: the highest row index wins deterministically (spconv requires unique
  // coordinates; the flag lets the host mirror raise)
  const int prev = atomicMax(&grid[cell], i);
  if (prev >= 0) atomicOr(err, 2);
}

// nbr[i, k] = row index of the voxel at coords[i] + dilation * (k_xyz - ks/2), or -1; k = (kx * ks + ky) * ks + kz
__global__ void __launch_bounds__(256) neighbor_map_kernel(const int* __restrict__ coords, int n, int B, int D, int ks,
                                                           int dilation, const int* __restrict__ grid,
                                                           int* __restrict__ nbr) {
  const int K3 = ks * ks * ks;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * K3) return;
  const int i = (int)(t / K3), k = (int)(t % K3);
  const int4 c = reinterpret_cast<const int4*>(coords)[i];
  const int half = ks / 2;
  const int x = c.y + dilation * (k / (ks * ks) - half);
  const int y = c.z + dilation * ((k / ks) % ks - half);
  const int z = c.w + dilation * (k % ks - half);
  int r = -1;
  if ((unsigned)c.x < (unsigned)B && (unsigned)x < (unsigned)D && (unsigned)y < (unsigned)D && (unsigned)z < (unsigned)D)
    r = grid[(((size_t)c.x * D + x) * D + y) * D + z];
  nbr[t] = r;
}

// im2col: out[i, k * Cin + c] = x[nbr[i, k], c] (0 where nbr < 0).  One thread moves 8 channels (16 B store);
// consecutive threads walk the channels of one (i, k) pair, so loads and stores are whole 128 B lines.
template <typename TIn>
__global__ void __launch_bounds__(256) sparse_im2col_kernel(const TIn* __restrict__ x, int ldx, const int* __restrict__ nbr,
                                                            long long pairs, int cin8, __half* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= pairs * cin8) return;
  const long long pr = t / cin8;
  const int c8 = (int)(t % cin8);
  const int src = nbr[pr];
  uint4 o = make_uint4(0u, 0u, 0u, 0u);
  if (src >= 0) {
    if constexpr (sizeof(TIn) == 2) {
      o = *reinterpret_cast<const uint4*>(x + (size_t)src * ldx + c8 * 8);
    } else {
      const float4 a = *reinterpret_cast<const float4*>(x + (size_t)src * ldx + c8 * 8);
      const float4 b = *reinterpret_cast<const float4*>(x + (size_t)src * ldx + c8 * 8 + 4);
      __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w);
      __half2 h2 = __floats2half2_rn(b.x, b.y), h3 = __floats2half2_rn(b.z, b.w);
      o.x = *reinterpret_cast<uint32_t*>(&h0);
      o.y = *reinterpret_cast<uint32_t*>(&h1);
      o.z = *reinterpret_cast<uint32_t*>(&h2);
      o.w = *reinterpret_cast<uint32_t*>(&h3);
    }
  }
  reinterpret_cast<uint4*>(out)[t] = o;
}

// ---------------------------------------------------------------------------------------
// SparseDownsample / SparseUpsample of the TRELLIS stage (trellis/modules/sparse/spatial.py:13-80).
// pool: out[p] = (sum of the rows of cell p) / (count + 1) -- the reference's torch.scatter_reduce(zeros, ..., 'mean') keeps
// its default include_self = True, so the zero initial value counts as one more element.  The children of a cell are
// listed in order[offsets[p] .. offsets[p + 1]) (ascending row order: the sum is deterministic, fp32, one fp16 rounding).
// One thread per (cell, 8 channels).
__global__ void __launch_bounds__(256) sparse_pool_mean_kernel(const __half* __restrict__ x, int ldx, const int* __restrict__ order,
                                                               const int* __restrict__ offsets, long long cells, int c8,
                                                               __half* __restrict__ out, int ldo) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= cells * c8) return;
  const long long p = gid / c8;
  const int c = (int)(gid - p * c8) * 8;
  const int s = offsets[p], e = offsets[p + 1];
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int i = s; i < e; ++i) {
    const uint4 t = __ldg(reinterpret_cast<const uint4*>(x + (size_t)order[i] * ldx + c));
    const __half2* h2 = reinterpret_cast<const __half2*>(&t);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float2 f = __half22float2(h2[u]);
      acc[2 * u] += f.x;
      acc[2 * u + 1] += f.y;
    }
  }
  const float inv = 1.0f / (float)(e - s + 1);
  uint4 pk;
  __half2* o2 = reinterpret_cast<__half2*>(&pk);
#pragma unroll
  for (int u = 0; u < 4; ++u) o2[u] = __floats2half2_rn(acc[2 * u] * inv, acc[2 * u + 1] * inv);
  *reinterpret_cast<uint4*>(out + (size_t)p * ldo + c) = pk;
}

// out[i] = [ a[idx ? idx[i] : i, 0:Ca] | b[i, 0:Cb] ]: the nearest-neighbour upsample (a gather through the cached cell
// index) and / or the skip concatenation of structured_latent_flow.py:253-256 in one pass.  One thread per 16 B piece.
__global__ void __launch_bounds__(256) gather_concat_kernel(const __half* __restrict__ a, int lda, int ca8, const int* __restrict__ idx,
                                                            const __half* __restrict__ b, int ldb, int cb8, long long rows,
                                                            __half* __restrict__ out, int ldo) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int w8 = ca8 + cb8;
  if (gid >= rows * w8) return;
  const long long r = gid / w8;
  const int c = (int)(gid - r * w8);
  uint4 v;
  if (c < ca8) v = __ldg(reinterpret_cast<const uint4*>(a + (size_t)(idx ? idx[r] : r) * lda + c * 8));
  else v = __ldg(reinterpret_cast<const uint4*>(b + (size_t)r * ldb + (c - ca8) * 8));
  *reinterpret_cast<uint4*>(out + (size_t)r * ldo + c * 8) = v;
}

// Submanifold convolution of a nearest-neighbour UPSAMPLED tensor without materialising it (the first output block of the
// structured-latent flow model, structured_latent_flow.py:166-172: 2 x 1024 channels at every fine voxel).  Every fine row
// is a copy of its coarse cell's row, so the per-tap products are computed ONCE per coarse row by a plain GEMM,
// P[c, k * Cout + o] = sum_i W[o, k, i] a[c, i] (fp32), and the convolution is a gather-sum over the 27 taps:
//   out[n, o] = fp16( bias[o] + sum_k P[idx[nbr[n, k]], k * Cout + o] )       (nbr < 0: absent neighbour)
// 5.4 x fewer flops than the convolution at the fine level (19792 vs 3656 rows) and no [N, 2048] operand.  idx == NULL:
// P is indexed by the neighbour row itself (the same formulation for a tensor that is not upsampled).
// One thread per (row, 4 channels): a warp reads 512 contiguous bytes of a P row per tap.
__global__ void __launch_bounds__(256) sparse_tap_gather_sum_kernel(const float* __restrict__ P, long long ldp, const int* __restrict__ nbr,
                                                                    const int* __restrict__ idx, long long rows, int K3, int co4,
                                                                    const float* __restrict__ bias, __half* __restrict__ out, int ldo) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= rows * co4) return;
  const long long r = gid / co4;
  const int c = (int)(gid - r * co4) * 4;
  const int cout = co4 * 4;
  float4 acc = bias ? __ldg(reinterpret_cast<const float4*>(bias + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  const int* nr = nbr + r * K3;
  for (int k = 0; k < K3; ++k) {
    int j = __ldg(nr + k);
    if (j < 0) continue;
    if (idx) j = __ldg(idx + j);
    const float4 v = __ldg(reinterpret_cast<const float4*>(P + (long long)j * ldp + (long long)k * cout + c));
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  __half2 lo = __floats2half2_rn(acc.x, acc.y), hi = __floats2half2_rn(acc.z, acc.w);
  uint2 pk;
  pk.x = *reinterpret_cast<const uint32_t*>(&lo);
  pk.y = *reinterpret_cast<const uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(out + r * ldo + c) = pk;
}

}  // namespace gvf

using namespace gvf;
#define ST(s) ((cudaStream_t)(s))
#define RET() return cudaGetLastError() == cudaSuccess ? GVF_OK : GVF_ERR_CUDA

extern "C" {

GVF_API int gvf_to_representation(const float* feats, int ldf, const int* coords, int nvox, int G,
                                  const float* perturbation, const float* lr, float resolution, int reg_mode,
                                  float voxel_size, float* xyz, float* features_dc, float* scaling,
                                  float* rotation, float* opacity, void* stream) {
  if (!feats || !coords || !lr || !xyz || !features_dc || !scaling || !rotation || !opacity) return GVF_ERR_INVALID;
  if (nvox <= 0 || G <= 0 || ldf < 14 * G || resolution <= 0.0f || reg_mode < 0 || reg_mode > 2) return GVF_ERR_INVALID;
  const long long n = (long long)nvox * G;
  to_representation_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>(
      feats, ldf, coords, nvox, G, perturbation, lr[0], lr[1], lr[2], lr[3], lr[4], resolution, reg_mode, voxel_size,
      xyz, features_dc, scaling, rotation, opacity);
  RET();
}

GVF_API int gvf_to_representation_bwd(const float* feats, int ldf, int nvox, int G, const float* perturbation,
                                      const float* lr /* host */, float resolution, int reg_mode, float voxel_size,
                                      const float* g_xyz, const float* g_dc, const float* g_scaling, const float* g_rotation,
                                      const float* g_opacity, float* g_feats, void* stream) {
  if (!feats || !lr || !g_feats || nvox <= 0 || G <= 0 || ldf < 14 * G || resolution <= 0.0f || reg_mode < 0 || reg_mode > 2)
    return GVF_ERR_INVALID;
  const long long n = (long long)nvox * G;
  to_representation_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>(
      feats, ldf, nvox, G, perturbation, lr[0], lr[1], lr[2], lr[3], lr[4], resolution, reg_mode, voxel_size, g_xyz, g_dc,
      g_scaling, g_rotation, g_opacity, g_feats);
  RET();
}

GVF_API size_t gvf_sparse_conv_workspace_bytes(int B, int D) {
  if (B <= 0 || D <= 0) return 0;
  return ((size_t)B * D * D * D + 4) * sizeof(int);       // grid + status word (+ padding)
}

GVF_API int gvf_sparse_neighbor_map(const int* coords, int N, int B, int D, int ksize, int dilation, void* workspace,
                                    size_t workspace_bytes, int* nbr, int* status, void* stream) {
  if (!coords || !workspace || !nbr || N <= 0 || B <= 0 || D <= 0 || dilation <= 0) return GVF_ERR_INVALID;
  if (ksize != 1 && ksize != 3 && ksize != 5) return GVF_ERR_UNSUPPORTED;
  if ((uintptr_t)coords & 15) return GVF_ERR_INVALID;
  if (workspace_bytes < gvf_sparse_conv_workspace_bytes(B, D)) return GVF_ERR_WORKSPACE;
  const size_t cells = (size_t)B * D * D * D;
  int* grid = (int*)workspace;
  int* err = grid + cells;
  if (cudaMemsetAsync(grid, 0xFF, cells * sizeof(int), ST(stream)) != cudaSuccess) return GVF_ERR_CUDA;
  if (cudaMemsetAsync(err, 0, sizeof(int), ST(stream)) != cudaSuccess) return GVF_ERR_CUDA;
  voxel_grid_fill_kernel<<<(N + 255) / 256, 256, 0, ST(stream)>>>(coords, N, B, D, grid, err);
  if (cudaGetLastError() != cudaSuccess) return GVF_ERR_CUDA;
  const long long t = (long long)N * ksize * ksize * ksize;
  neighbor_map_kernel<<<(unsigned)((t + 255) / 256), 256, 0, ST(stream)>>>(coords, N, B, D, ksize, dilation, grid, nbr);
  if (cudaGetLastError() != cudaSuccess) return GVF_ERR_CUDA;
  if (status && cudaMemcpyAsync(status, err, sizeof(int), cudaMemcpyDeviceToDevice, ST(stream)) != cudaSuccess)
    return GVF_ERR_CUDA;
  return GVF_OK;
}

GVF_API int gvf_sparse_im2col_f16(const void* x, int x_is_f16, int ldx, const int* nbr, int N, int K3, int Cin,
                                  void* out, void* stream) {
  if (!x || !nbr || !out || N <= 0 || K3 <= 0 || Cin <= 0) return GVF_ERR_INVALID;
  if ((Cin % 8) || (ldx % 8) || (((uintptr_t)x | (uintptr_t)out) & 15)) return GVF_ERR_INVALID;
  const long long pairs = (long long)N * K3;
  const int cin8 = Cin / 8;
  const long long t = pairs * cin8;
  const unsigned blocks = (unsigned)((t + 255) / 256);
  if (x_is_f16)
    sparse_im2col_kernel<__half><<<blocks, 256, 0, ST(stream)>>>((const __half*)x, ldx, nbr, pairs, cin8, (__half*)out);
  else
    sparse_im2col_kernel<float><<<blocks, 256, 0, ST(stream)>>>((const float*)x, ldx, nbr, pairs, cin8, (__half*)out);
  RET();
}

GVF_API int gvf_sparse_pool_mean_f16(const void* x, int ldx, const int* order, const int* offsets, int cells, int C,
                                     void* out, int ldo, void* stream) {
  if (!x || !order || !offsets || !out || cells <= 0 || C <= 0) return GVF_ERR_INVALID;
  if ((C % 8) || (ldx % 8) || (ldo % 8) || (((uintptr_t)x | (uintptr_t)out) & 15)) return GVF_ERR_INVALID;
  const long long t = (long long)cells * (C / 8);
  sparse_pool_mean_kernel<<<(unsigned)((t + 255) / 256), 256, 0, ST(stream)>>>((const __half*)x, ldx, order, offsets, cells,
                                                                            C / 8, (__half*)out, ldo);
  RET();
}

GVF_API int gvf_gather_concat_f16(const void* a, int lda, int Ca, const int* idx, const void* b, int ldb, int Cb, int rows,
                                  void* out, int ldo, void* stream) {
  if (!out || rows <= 0 || Ca < 0 || Cb < 0 || Ca + Cb <= 0 || (Ca > 0 && !a) || (Cb > 0 && !b)) return GVF_ERR_INVALID;
  if ((Ca % 8) || (Cb % 8) || (lda % 8) || (ldb % 8) || (ldo % 8) || ldo < Ca + Cb) return GVF_ERR_INVALID;
  if (((uintptr_t)a | (uintptr_t)b | (uintptr_t)out) & 15) return GVF_ERR_INVALID;
  const long long t = (long long)rows * ((Ca + Cb) / 8);
  gather_concat_kernel<<<(unsigned)((t + 255) / 256), 256, 0, ST(stream)>>>((const __half*)a, lda, Ca / 8, idx, (const __half*)b,
                                                                         ldb, Cb / 8, rows, (__half*)out, ldo);
  RET();
}

GVF_API int gvf_sparse_tap_gather_sum_f16(const float* P, long long ldp, const int* nbr, const int* idx, int N, int K3, int Cout,
                                          const float* bias, void* out, int ldo, void* stream) {
  if (!P || !nbr || !out || N <= 0 || K3 <= 0 || Cout <= 0) return GVF_ERR_INVALID;
  if ((Cout % 4) || (ldp % 4) || (ldo % 4) || ldp < (long long)K3 * Cout) return GVF_ERR_INVALID;
  if ((((uintptr_t)P | (uintptr_t)bias) & 15) || ((uintptr_t)out & 7)) return GVF_ERR_INVALID;
  const long long t = (long long)N * (Cout / 4);
  sparse_tap_gather_sum_kernel<<<(unsigned)((t + 255) / 256), 256, 0, ST(stream)>>>(P, ldp, nbr, idx, N, K3, Cout / 4, bias,
                                                                                 (__half*)out, ldo);
  RET();
}

}  // extern "C"
