// mufu_probe.cu -- measures the exp2 (MUFU.EX2) issue rate per SM and the rate of the softmax inner
// loop mix (FFMA + EX2 + FADD + FMNMX + F2FP), to put a measured ceiling under the attention roofline.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE>
__global__ void k(float* out, int iters, float c, float mc) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = -0.01f * (threadIdx.x + i);
  float s = 0.f, mx = -1e30f;
  unsigned pk = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      if (MODE == 0) { v[i] = ex2(v[i]); v[i + 1] = ex2(v[i + 1]); }
      else {
        mx = fmaxf(mx, v[i]); mx = fmaxf(mx, v[i + 1]);
        float p0 = ex2(fmaf(v[i], c, -mc)), p1 = ex2(fmaf(v[i + 1], c, -mc));
        s += p0 + p1;
        __half2 h = __floats2half2_rn(p0, p1);
        pk ^= *reinterpret_cast<unsigned*>(&h);
        v[i] = p0 - 1.5f; v[i + 1] = p1 - 1.5f;
      }
    }
  }
  long long t1 = clock64();
  float acc = s + mx + __uint_as_float(pk);
#pragma unroll
  for (int i = 0; i < 32; ++i) acc += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}
int main() {
  float* d; cudaMalloc(&d, 1 << 24);
  const int iters = 2000;
  for (int mode = 0; mode < 2; ++mode)
    for (int warps = 4; warps <= 16; warps *= 2) {
      float h;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, warps * 32>>>(d, iters, 1.1f, 0.3f); else k<1><<<148, warps * 32>>>(d, iters, 1.1f, 0.3f);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost);
      double ex_per_clk = (double)warps * 32 * 32 * iters / h;   // exp2 results per clock per SM
      printf("mode %d (%s) warps/SM %2d: %.0f cycles, %.2f exp2/clk/SM, %.2f cycles per warp-wide EX2 per SMSP\n", mode,
             mode ? "softmax mix" : "EX2 only", warps, h, ex_per_clk, 32.0 * 4 / ex_per_clk);
    }
  return 0;
}
