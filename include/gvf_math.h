/* gvf_math.h -- reproducible single-precision transcendentals.
 *
 * The rasteriser's *integer* outputs (radii, tile rectangles, depth-sort keys, per-tile
 * lists) must be bit-identical between the sm_100a kernels and the CPU oracle
 * (BASELINE.json north_star: "bit-exact on tile/sort indices").  libm (glibc) and the
 * CUDA math library round expf/logf differently, so every transcendental that feeds an
 * index is computed with the routines below instead: they use only IEEE-754 +,-,*,/,
 * sqrtf, rintf and integer bit moves, all of which are correctly rounded on both sides
 * as long as FMA contraction is off (nvcc -fmad=false for the preprocess TU,
 * gcc -ffp-contract=off for the oracle).
 *
 * Accuracy (checked in tests/test_gvf_math.py against float64): expf <= 2 ulp on
 * [-87, 88]; logf <= 2 ulp; log1pf <= 4 ulp; softplus/sigmoid follow torch's formulas
 * (softplus threshold 20, beta 1).
 *
 * C99 / CUDA C++; header-only.
 */
#ifndef GVF_MATH_H_
#define GVF_MATH_H_

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define GVF_HD __host__ __device__ __forceinline__
#else
#define GVF_HD static inline
#endif

GVF_HD uint32_t gvf_f2u(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
#endif
}

GVF_HD float gvf_u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}

/* e^x, Cody-Waite reduction + degree-5 polynomial (Cephes coefficients). */
GVF_HD float gvf_expf(float x) {
  if (x > 88.0f) x = 88.0f;
  if (x < -87.0f) return 0.0f;
  float n = rintf(x * 1.44269504088896341f);
  float r = x - n * 0.693359375f;
  r = r - n * -2.12194440e-4f;
  float p = 1.9875691500e-4f;
  p = p * r + 1.3981999507e-3f;
  p = p * r + 8.3334519073e-3f;
  p = p * r + 4.1665795894e-2f;
  p = p * r + 1.6666665459e-1f;
  p = p * r + 5.0000001201e-1f;
  p = p * (r * r) + r;
  p = p + 1.0f;
  int32_t e = (int32_t)n;
  /* 2^e by exponent-field construction; e in [-126, 127] after the clamps above */
  return p * gvf_u2f((uint32_t)(e + 127) << 23);
}

/* natural log of a positive normal float (Cephes logf). */
GVF_HD float gvf_logf(float x) {
  uint32_t u = gvf_f2u(x);
  int32_t e = (int32_t)((u >> 23) & 0xff) - 126;           /* x = m * 2^e, m in [0.5, 1) */
  float m = gvf_u2f((u & 0x007fffffu) | 0x3f000000u);
  if (m < 0.707106781186547524f) {
    e = e - 1;
    m = m + m - 1.0f;
  } else {
    m = m - 1.0f;
  }
  float z = m * m;
  float p = 7.0376836292e-2f;
  p = p * m + -1.1514610310e-1f;
  p = p * m + 1.1676998740e-1f;
  p = p * m + -1.2420140846e-1f;
  p = p * m + 1.4249322787e-1f;
  p = p * m + -1.6668057665e-1f;
  p = p * m + 2.0000714765e-1f;
  p = p * m + -2.4999993993e-1f;
  p = p * m + 3.3333331174e-1f;
  float y = m * z * p;
  float fe = (float)e;
  y = y + fe * -2.12194440e-4f;
  y = y - 0.5f * z;
  float r = m + y;
  r = r + fe * 0.693359375f;
  return r;
}

/* log(1+y) for y >= 0 (Kahan's correction of log(1+y)). */
GVF_HD float gvf_log1pf(float y) {
  float u = 1.0f + y;
  if (u == 1.0f) return y;
  return gvf_logf(u) * (y / (u - 1.0f));
}

/* torch.nn.functional.softplus(x) with beta=1, threshold=20 */
GVF_HD float gvf_softplusf(float x) {
  if (x > 20.0f) return x;
  return gvf_log1pf(gvf_expf(x));
}

/* torch.sigmoid */
GVF_HD float gvf_sigmoidf(float x) {
  return 1.0f / (1.0f + gvf_expf(-x));
}

#endif /* GVF_MATH_H_ */
