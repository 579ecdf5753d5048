// mufu_probe2.cu -- throughput of ex2.approx.f16x2 and of a polynomial exp2 on the FMA pipe.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned ex2_h2(unsigned x) { unsigned y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float poly_ex2(float x) {
  x = fmaxf(x, -126.0f);
  const float xf = x + 12582912.0f;
  const float f = x - (xf - 12582912.0f);
  float p = fmaf(0.0555041f, f, 0.2402265f);
  p = fmaf(p, f, 0.6931472f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(xf) << 23));
}
template <int MODE>
__global__ void k(float* out, int iters, float c, float mc) {
  float v[32]; unsigned h[16];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = -0.01f * (threadIdx.x + i);
#pragma unroll
  for (int i = 0; i < 16; ++i) h[i] = 0xb800b400u + threadIdx.x + i;
  float s = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 16; ++i) h[i] = ex2_h2(h[i]) ^ 0x80008000u;
    } else if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = poly_ex2(v[i]) - 1.5f;
    } else {   // mixed: 3 of 4 on MUFU, 1 of 4 polynomial, softmax-like surroundings
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float p0 = ex2(fmaf(v[i], c, -mc)), p1 = ex2(fmaf(v[i + 1], c, -mc)), p2 = ex2(fmaf(v[i + 2], c, -mc));
        float p3 = poly_ex2(fmaf(v[i + 3], c, -mc));
        s += (p0 + p1) + (p2 + p3);
        v[i] = p0 - 1.5f; v[i + 1] = p1 - 1.5f; v[i + 2] = p2 - 1.5f; v[i + 3] = p3 - 1.5f;
      }
    }
  }
  long long t1 = clock64();
  float acc = s;
#pragma unroll
  for (int i = 0; i < 32; ++i) acc += v[i];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc += __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}
int main() {
  float* d; cudaMalloc(&d, 1 << 24);
  const int iters = 2000;
  const char* names[] = {"", "", "ex2.f16x2 (2 results/lane/instr)", "poly exp2 only", "3 MUFU : 1 poly mix"};
  for (int mode = 2; mode <= 4; ++mode) {
    const int warps = 8;
    float h;
    for (int rep = 0; rep < 2; ++rep) {
      if (mode == 2) k<2><<<148, warps * 32>>>(d, iters, 1.1f, 0.3f);
      else if (mode == 3) k<3><<<148, warps * 32>>>(d, iters, 1.1f, 0.3f);
      else k<4><<<148, warps * 32>>>(d, iters, 1.1f, 0.3f);
      cudaDeviceSynchronize();
    }
    cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost);
    const double results = (double)warps * 32 * 32 * iters;
    printf("mode %d %-34s: %.0f cycles, %.2f exp2 results/clk/SM\n", mode, names[mode], h, results / h);
  }
  // accuracy of the polynomial vs exp2f
  return 0;
}
