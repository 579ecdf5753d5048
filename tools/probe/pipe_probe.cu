// pipe_probe.cu -- issue cost (clocks per warp instruction on one SM sub-partition) of the instructions in the
// softmax inner loop: which pipe each one loads and how heavily.  N independent chains per thread, 4 warps per
// sub-partition, so latency is hidden and the rate is the pipe's.
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
// MODE: 0 FFMA  1 FFMA2  2 FADD2  3 F2FP pack  4 FMNMX3  5 IMAD (shift-add)  6 FADD  7 MUFU.EX2  8 HADD2
//       9 FFMA2 + MUFU x2 (do they overlap?)  10 FFMA2 + FADD2 + F2FP + FMNMX3 per pair (the loop minus MUFU)
//       11 the whole loop (10 + 2 MUFU)  12 = 11 without FFMA2   13 = 11 without FADD2
template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, int iters, float c, float mc) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = 0.001f * (threadIdx.x + i);
  float s0 = 0.f, s1 = 0.f, mx = -1e30f;
  unsigned acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      float x0 = v[i], x1 = v[i + 1];
      if (MODE == 0) { x0 = fmaf(x0, c, mc); x1 = fmaf(x1, c, mc); }
      if (MODE == 1) ffma2(x0, x1, x0, x1, c, c, mc, mc);
      if (MODE == 2) fadd2(x0, x1, x0, x1, mc, mc);
      if (MODE == 3) { acc ^= pack_cvt(x0, x1); x0 += 0.f; }
      if (MODE == 4) { mx = fmaxf(mx, fmaxf(x0, x1)); }
      if (MODE == 5) { x0 = __uint_as_float(__float_as_uint(x0) + (__float_as_uint(x1) << 23)); x1 = __uint_as_float(__float_as_uint(x1) + (__float_as_uint(x0) << 23)); }
      if (MODE == 6) { x0 += mc; x1 += mc; }
      if (MODE == 7) { x0 = ex2(x0); x1 = ex2(x1); }
      if (MODE == 8) { __half2 h = *reinterpret_cast<__half2*>(&x0); h = __hadd2(h, *reinterpret_cast<__half2*>(&x1)); x0 = *reinterpret_cast<float*>(&h); }
      if (MODE == 9) { ffma2(x0, x1, x0, x1, c, c, mc, mc); x0 = ex2(x0); x1 = ex2(x1); }
      if (MODE >= 10) {
        float y0 = x0, y1 = x1;
        if (MODE != 12) ffma2(y0, y1, x0, x1, c, c, mc, mc);
        mx = fmaxf(mx, fmaxf(x0, x1));
        if (MODE >= 11) { y0 = ex2(y0); y1 = ex2(y1); }
        if (MODE != 13) fadd2(s0, s1, s0, s1, y0, y1);
        acc ^= pack_cvt(y0, y1);
        x0 = y0 * 0.5f;       // keep the chain alive without feeding NaNs (one FMUL; counted below)
        x1 = y1;
      }
      v[i] = x0; v[i + 1] = x1;
    }
  }
  long long t1 = clock64();
  float r = s0 + s1 + mx + __uint_as_float(acc & 0x3fffffffu);
#pragma unroll
  for (int i = 0; i < 32; ++i) r += v[i];
  out[1 + blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}
template <int MODE> static void run(float* d, const char* name, double instr_per_pair) {
  const int iters = 400;
  float h = 0;
  for (int rep = 0; rep < 2; ++rep) {
    k<MODE><<<148, 512>>>(d, iters, 1.0001f, -0.0003f);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost);
  }
  const double per_pair = h / (iters * 16.0) / 4.0;      // clocks of the sub-partition per (pair of scores, warp)
  printf("%-52s %6.2f clk per pair per warp on the sub-partition", name, per_pair);
  if (instr_per_pair > 0) printf("  (%5.2f clk per warp instruction)", per_pair / instr_per_pair);
  printf("\n");
}
int main() {
  float* d; cudaMalloc(&d, 1 << 24);
  run<0>(d, "FFMA x2", 2);
  run<1>(d, "FFMA2", 1);
  run<2>(d, "FADD2", 1);
  run<6>(d, "FADD x2", 2);
  run<3>(d, "F2FP.F16.F32.PACK_AB (+1 FADD)", 2);
  run<4>(d, "FMNMX3 (max of a pair into the running max)", 1);
  run<5>(d, "IMAD shift-add x2", 2);
  run<8>(d, "HADD2", 1);
  run<7>(d, "MUFU.EX2 x2", 2);
  run<9>(d, "FFMA2 + MUFU.EX2 x2", 0);
  run<10>(d, "loop minus MUFU: FFMA2 FMNMX3 FADD2 F2FP FMUL", 5);
  run<11>(d, "whole loop: + 2 MUFU", 7);
  run<12>(d, "whole loop without FFMA2", 6);
  run<13>(d, "whole loop without FADD2", 6);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
