// mufu_probe3.cu -- which instructions of the softmax inner loop share the MUFU (XU) pipe?
// clocks per exponential for 1 and 2 warps per SM sub-partition, for several instruction mixes.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {
  asm volatile("{\n.reg .b64 ra, rb, rc, rd;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\nmov.b64 rc, {%6, %7};\n"
      "fma.rn.f32x2 rd, ra, rb, rc;\nmov.b64 {%0, %1}, rd;\n}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
__device__ __forceinline__ void fadd2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm volatile("{\n.reg .b64 ra, rb, rd;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\n"
      "add.rn.f32x2 rd, ra, rb;\nmov.b64 {%0, %1}, rd;\n}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ unsigned pack_cvt(float a, float b) {
  unsigned r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r;
}
__device__ __forceinline__ unsigned pack_int(float a, float b) {   // exponent-biased fp32 -> fp16 bits, round half up
  const unsigned ua = __float_as_uint(a) + 0x1000u, ub = __float_as_uint(b) + 0x1000u;
  return __byte_perm(ua >> 13, ub << 3, 0x7610);
}
// MODE 0: ex2 only   1: ex2 + cvt pack   2: ex2 + integer pack   3: ffma2 + ex2 + fadd2 + cvt pack
//      4: ffma2 + ex2 + fadd2 + integer pack   5: ffma2 + ex2 + fadd2 (no pack)
template <int MODE>
__global__ void k(float* out, int iters, float c, float mc) {
  float v[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) v[i] = -0.01f * (threadIdx.x + i);
  float s0 = 0.f, s1 = 0.f;
  unsigned acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 64; i += 2) {
      float x0 = v[i], x1 = v[i + 1];
      if (MODE >= 3) ffma2(x0, x1, x0, x1, c, c, mc, mc);
      const float p0 = ex2(x0), p1 = ex2(x1);
      if (MODE >= 3) fadd2(s0, s1, s0, s1, p0, p1);
      if (MODE == 1 || MODE == 3) acc ^= pack_cvt(p0, p1);
      else if (MODE == 2 || MODE == 4) acc ^= pack_int(p0, p1);
      v[i] = p0 - 1.5f; v[i + 1] = p1 - 1.5f;
    }
  }
  long long t1 = clock64();
  float r = s0 + s1 + __uint_as_float(acc & 0x3fffffffu);
#pragma unroll
  for (int i = 0; i < 64; ++i) r += v[i];
  out[1 + blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}
template <int MODE> static void run(float* d, const char* name) {
  const int iters = 500;
  for (int wps = 1; wps <= 2; wps *= 2) {
    float h = 0;
    for (int rep = 0; rep < 2; ++rep) {
      k<MODE><<<148, wps * 128>>>(d, iters, 1.1f, -0.3f);
      cudaDeviceSynchronize();
      cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost);
    }
    printf("%-44s warps/SMSP %d: %6.2f clk per exponential per warp, %6.2f clk per MUFU instr on the pipe\n", name, wps,
           h / (iters * 64.0), h / (iters * 64.0) / wps);
  }
}
int main() {
  float* d; cudaMalloc(&d, 1 << 24);
  run<0>(d, "ex2 only");
  run<1>(d, "ex2 + cvt.rn.f16x2.f32");
  run<2>(d, "ex2 + integer pack");
  run<5>(d, "ffma2 + ex2 + fadd2");
  run<3>(d, "ffma2 + ex2 + fadd2 + cvt pack");
  run<4>(d, "ffma2 + ex2 + fadd2 + integer pack");
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
