// output_stage.cu -- the image post-processing of the reference's visualisation loop (utils/inference_utils.py:276-297):
// every rendered 512^2 frame, already clamped and converted to uint8 (gvf_rgba_to_u8), is resized to
// int(512 * scale_factor) with PIL's LANCZOS filter and centre-padded (white) or centre-cropped back to 512^2.  The
// reference does this per frame on the host (`.cpu()` + PIL, 24 x 128 frames per object); here the frames stay on the
// device: two separable passes with PIL's own arithmetic -- 8-bit fixed-point coefficients (22 fractional bits, computed
// on the host in double exactly like Pillow's precompute_coeffs / normalize_coeffs_8bpc), horizontal pass first, an 8-bit
// intermediate image, rounding by adding 2^21 and clamping to [0, 255] -- so the result equals Image.resize(..., LANCZOS)
// byte for byte; then one pad / crop kernel.  HBM-class: 0.8 MB read + written per frame and pass.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gvf_b200.h"

namespace gvf {

constexpr int kPrecisionBits = 32 - 8 - 2;

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= kPrecisionBits;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// out[f, y, xo, c] = clip8(2^21 + sum_x in[f, y, xmin + x, c] * k[xo][x])        (HWC uint8, C = 3)
__global__ void __launch_bounds__(256) resample_h_kernel(const uint8_t* __restrict__ in, int F, int H, int Win, int Wout,
                                                         const int* __restrict__ bounds, const int* __restrict__ kk,
                                                         int ksize, uint8_t* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)F * H * Wout;
  if (i >= n) return;
  const int xo = (int)(i % Wout);
  const long long row = i / Wout;                      // f * H + y
  const int xmin = bounds[2 * xo], xmax = bounds[2 * xo + 1];
  const int* k = kk + (size_t)xo * ksize;
  const uint8_t* src = in + (row * Win + xmin) * 3;
  int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  for (int x = 0; x < xmax; ++x) {
    const int w = __ldg(k + x);
    s0 += src[3 * x] * w; s1 += src[3 * x + 1] * w; s2 += src[3 * x + 2] * w;
  }
  uint8_t* dst = out + i * 3;
  dst[0] = clip8(s0); dst[1] = clip8(s1); dst[2] = clip8(s2);
}

// out[f, yo, x, c] = clip8(2^21 + sum_y in[f, ymin + y, x, c] * k[yo][y])
__global__ void __launch_bounds__(256) resample_v_kernel(const uint8_t* __restrict__ in, int F, int Hin, int W, int Hout,
                                                         const int* __restrict__ bounds, const int* __restrict__ kk,
                                                         int ksize, uint8_t* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)F * Hout * W;
  if (i >= n) return;
  const int x = (int)(i % W);
  const int yo = (int)((i / W) % Hout);
  const long long f = i / ((long long)W * Hout);
  const int ymin = bounds[2 * yo], ymax = bounds[2 * yo + 1];
  const int* k = kk + (size_t)yo * ksize;
  const uint8_t* src = in + ((f * Hin + ymin) * W + x) * 3;
  int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  for (int y = 0; y < ymax; ++y) {
    const int w = __ldg(k + y);
    const uint8_t* p = src + (size_t)y * W * 3;
    s0 += p[0] * w; s1 += p[1] * w; s2 += p[2] * w;
  }
  uint8_t* dst = out + i * 3;
  dst[0] = clip8(s0); dst[1] = clip8(s1); dst[2] = clip8(s2);
}

// centre pad with `fill` (smaller input: pasted at ((S - W) / 2, (S - H) / 2)) or centre crop (larger input) to S x S
__global__ void __launch_bounds__(256) pad_crop_kernel(const uint8_t* __restrict__ in, int F, int Hin, int Win, int S,
                                                       int fill, uint8_t* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)F * S * S;
  if (i >= n) return;
  const int x = (int)(i % S), y = (int)((i / S) % S);
  const long long f = i / ((long long)S * S);
  int sx, sy;
  if (Hin < S || Win < S) {                            // reference :286-292 (pad_w / pad_h = max(0, (512 - size) // 2))
    const int pw = (S - Win) / 2 > 0 ? (S - Win) / 2 : 0, ph = (S - Hin) / 2 > 0 ? (S - Hin) / 2 : 0;
    sx = x - pw; sy = y - ph;
  } else {                                             // :293-297
    sx = x + (Win - S) / 2; sy = y + (Hin - S) / 2;
  }
  uint8_t* dst = out + i * 3;
  if (sx >= 0 && sx < Win && sy >= 0 && sy < Hin) {
    const uint8_t* p = in + ((f * Hin + sy) * Win + sx) * 3;
    dst[0] = p[0]; dst[1] = p[1]; dst[2] = p[2];
  } else {
    dst[0] = dst[1] = dst[2] = (uint8_t)fill;
  }
}

}  // namespace gvf

using namespace gvf;

extern "C" {

GVF_API int gvf_resample_u8(const uint8_t* in, int F, int Hin, int Win, int Hout, int Wout, const int* bounds_h,
                            const int* coef_h, int ksize_h, const int* bounds_v, const int* coef_v, int ksize_v,
                            uint8_t* tmp, uint8_t* out, void* stream) {
  if (!in || !out || !tmp || !bounds_h || !coef_h || !bounds_v || !coef_v) return GVF_ERR_INVALID;
  if (F <= 0 || Hin <= 0 || Win <= 0 || Hout <= 0 || Wout <= 0 || ksize_h <= 0 || ksize_v <= 0) return GVF_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  const long long n1 = (long long)F * Hin * Wout, n2 = (long long)F * Hout * Wout;
  resample_h_kernel<<<(unsigned)((n1 + 255) / 256), 256, 0, st>>>(in, F, Hin, Win, Wout, bounds_h, coef_h, ksize_h, tmp);
  if (cudaGetLastError() != cudaSuccess) return GVF_ERR_CUDA;
  resample_v_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(tmp, F, Hin, Wout, Hout, bounds_v, coef_v, ksize_v, out);
  return cudaGetLastError() == cudaSuccess ? GVF_OK : GVF_ERR_CUDA;
}

GVF_API int gvf_pad_crop_u8(const uint8_t* in, int F, int Hin, int Win, int S, int fill, uint8_t* out, void* stream) {
  if (!in || !out || F <= 0 || Hin <= 0 || Win <= 0 || S <= 0) return GVF_ERR_INVALID;
  const long long n = (long long)F * S * S;
  pad_crop_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, F, Hin, Win, S, fill, out);
  return cudaGetLastError() == cudaSuccess ? GVF_OK : GVF_ERR_CUDA;
}

}  // extern "C"
