// vox2seq.cu -- 3-D voxel coordinate <-> 30-bit space-filling-curve code (Morton / Hilbert).
//
// Replaces the reference's first-party extension model/sparse_voxel_diffusion/vox2seq/src/
// {z_order,hilbert}.cu + api.cu (four kernels over separate x/y/z arrays, used by the serialized
// sparse attention, sparse/attention/serialized_attn.py:67-74).  Here one kernel per direction reads
// the [N,3] int32 coordinates as they are stored (AoS, the axis permutation is a kernel argument, so
// the host never materialises permuted / split copies) and writes the code.  Integer ALU work,
// HBM-bound: 12 B in + 4 B out per voxel.
//   Morton: bit interleave by magic multiplies.  Hilbert: Skilling's transpose algorithm
//   ("Programming the Hilbert curve", AIP Conf. Proc. 707, 2004), 10 bits per axis.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gvf_b200.h"

namespace gvf {

__device__ __forceinline__ uint32_t spread3(uint32_t v) {     // 10 bits -> every third bit
  v = (v * 0x00010001u) & 0xFF0000FFu;
  v = (v * 0x00000101u) & 0x0F00F00Fu;
  v = (v * 0x00000011u) & 0xC30C30C3u;
  v = (v * 0x00000005u) & 0x49249249u;
  return v;
}
__device__ __forceinline__ uint32_t gather3(uint32_t v) {     // inverse of spread3
  v &= 0x49249249u;
  v = (v ^ (v >> 2)) & 0x030C30C3u;
  v = (v ^ (v >> 4)) & 0x0300F00Fu;
  v = (v ^ (v >> 8)) & 0x030000FFu;
  v = (v ^ (v >> 16)) & 0x000003FFu;
  return v;
}

__global__ void __launch_bounds__(256) vox_encode_kernel(const int32_t* __restrict__ coords, long long N, int p0,
                                                         int p1, int p2, int hilbert, int32_t* __restrict__ codes) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  uint32_t X[3] = {(uint32_t)coords[i * 3 + p0], (uint32_t)coords[i * 3 + p1], (uint32_t)coords[i * 3 + p2]};
  if (hilbert) {
    // axes -> transpose (Skilling): undo excess work from the top bit down, then Gray encode
    for (uint32_t Q = 1u << 9; Q > 1; Q >>= 1) {
      const uint32_t P = Q - 1;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        if (X[d] & Q) X[0] ^= P;
        else { const uint32_t t = (X[0] ^ X[d]) & P; X[0] ^= t; X[d] ^= t; }
      }
    }
    X[1] ^= X[0];
    X[2] ^= X[1];
    uint32_t t = 0;
    for (uint32_t Q = 1u << 9; Q > 1; Q >>= 1)
      if (X[2] & Q) t ^= Q - 1;
    X[0] ^= t; X[1] ^= t; X[2] ^= t;
  }
  codes[i] = (int32_t)(spread3(X[0]) * 4u + spread3(X[1]) * 2u + spread3(X[2]));
}

__global__ void __launch_bounds__(256) vox_decode_kernel(const int32_t* __restrict__ codes, long long N, int p0,
                                                         int p1, int p2, int hilbert, int32_t* __restrict__ coords) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const uint32_t c = (uint32_t)codes[i];
  uint32_t X[3] = {gather3(c >> 2), gather3(c >> 1), gather3(c)};
  if (hilbert) {
    uint32_t t = X[2] >> 1;
    X[2] ^= X[1];
    X[1] ^= X[0];
    X[0] ^= t;
    for (uint32_t Q = 2; Q != (2u << 9); Q <<= 1) {
      const uint32_t P = Q - 1;
#pragma unroll
      for (int d = 2; d >= 0; --d) {
        if (X[d] & Q) X[0] ^= P;
        else { t = (X[0] ^ X[d]) & P; X[0] ^= t; X[d] ^= t; }
      }
    }
  }
  // X holds the permuted axes (coords[:, permute]); scatter them back to their columns
  coords[i * 3 + p0] = (int32_t)X[0];
  coords[i * 3 + p1] = (int32_t)X[1];
  coords[i * 3 + p2] = (int32_t)X[2];
}

}  // namespace gvf

static bool perm_ok(const int* p) {
  return p && p[0] >= 0 && p[0] < 3 && p[1] >= 0 && p[1] < 3 && p[2] >= 0 && p[2] < 3 && p[0] != p[1] &&
         p[0] != p[2] && p[1] != p[2];
}

extern "C" GVF_API int gvf_vox2seq_encode(const int32_t* coords, long long N, const int* permute, int hilbert,
                                          int32_t* codes, void* stream) {
  if (N == 0 && perm_ok(permute)) return GVF_OK;
  if (!coords || !codes || N < 0 || !perm_ok(permute)) return GVF_ERR_INVALID;
  gvf::vox_encode_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      coords, N, permute[0], permute[1], permute[2], hilbert, codes);
  return cudaGetLastError() == cudaSuccess ? GVF_OK : GVF_ERR_CUDA;
}

extern "C" GVF_API int gvf_vox2seq_decode(const int32_t* codes, long long N, const int* permute, int hilbert,
                                          int32_t* coords, void* stream) {
  if (N == 0 && perm_ok(permute)) return GVF_OK;
  if (!coords || !codes || N < 0 || !perm_ok(permute)) return GVF_ERR_INVALID;
  gvf::vox_decode_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      codes, N, permute[0], permute[1], permute[2], hilbert, coords);
  return cudaGetLastError() == cudaSuccess ? GVF_OK : GVF_ERR_CUDA;
}
