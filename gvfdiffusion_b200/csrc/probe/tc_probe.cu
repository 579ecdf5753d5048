// tc_probe.cu -- stand-alone known-answer probe of the tcgen05 / TMA building blocks the
// attention kernels rely on (descriptor encodings, swizzle modes, MN-major B, A-from-TMEM).
// Not part of libgvf_b200.so.  One (test, variant) per process so that a faulting variant
// cannot poison the others:   tc_probe <test> <variant>   -> prints "max_abs_err".
//
//   test 0: D[128x128] = A[128x32]  * B[128x32]^T   K-major/K-major, SWIZZLE_64B   (QK^T, d=32)
//   test 1: D[128x128] = A[128x64]  * B[128x64]^T   K-major/K-major, SWIZZLE_128B  (QK^T, d=64)
//   test 2: D[128x32]  = P[128x128] * V[128x32]     A K-major SW128 (2 atoms), B MN-major SW64 (PV, d=32)
//   test 3: D[128x64]  = P[128x128] * V[128x64]     A K-major SW128, B MN-major SW128       (PV, d=64)
//   test 4: as test 2 but P comes from TMEM (tcgen05.st by the owning row threads)   (TS mode)
//   test 5: as test 3 but P from TMEM
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>

#include "../tc_common.cuh"

using namespace gvf::tc;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                CUtensorMapFloatOOBfill);

static EncodeTiled get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn) { printf("no cuTensorMapEncodeTiled\n"); exit(2); }
  return (EncodeTiled)fn;
}

static CUtensorMap make_map_2d(const __half* g, int rows, int cols, int box_rows, int box_cols,
                               CUtensorMapSwizzle sw) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(__half)};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)g, dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d\n", (int)r); exit(2); }
  return m;
}

struct Cfg {
  int N;            // MMA N
  int K;            // total K
  int a_from_tmem;  // TS mode
  int a_row_bytes;  // bytes per A smem row inside one atom (64 or 128)
  int a_atoms;      // number of K atoms of A (each a separate [128 x a_row_bytes] tile)
  int b_mn_major;
  int b_row_bytes;  // bytes per B smem row (64 / 128)
  int b_rows;       // rows of the B tile in smem
  uint64_t a_swz, b_swz;
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
  uint32_t b_kstep_bytes;  // descriptor start-address advance of B per K=16 step
};

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap mapA,
                                                    const __grid_constant__ CUtensorMap mapB,
                                                    const __half* __restrict__ Pglobal,
                                                    float* __restrict__ out, Cfg c) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_s;
  uint8_t* sbase = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);  // swizzle atoms need 1 KB alignment
  uint8_t* sA = sbase;                 // up to 2 atoms x 16 KB
  uint8_t* sB = sbase + 32768;         // up to 16 KB
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tmem_D = tmem;          // columns [0, N)
  const uint32_t tmem_P = tmem + 128;    // columns [128, 128 + K/2) for TS mode

  if (tid == 0) {
    uint32_t bytes = (uint32_t)c.b_rows * c.b_row_bytes;
    if (!c.a_from_tmem) bytes += (uint32_t)c.a_atoms * 128 * c.a_row_bytes;
    mbar_arrive_expect_tx(&bar_load, bytes);
    if (!c.a_from_tmem)
      for (int at = 0; at < c.a_atoms; ++at)
        tma_load_2d(sA + at * 128 * c.a_row_bytes, &mapA, &bar_load, at * (c.a_row_bytes / 2), 0);
    tma_load_2d(sB, &mapB, &bar_load, 0, 0);
  }
  if (c.a_from_tmem) {
    // each thread owns row (32*warp + lane): K fp16 = K/2 packed words
    const int row = warp * 32 + lane;
    const uint32_t* prow = reinterpret_cast<const uint32_t*>(Pglobal + (size_t)row * c.K);
    for (int w0 = 0; w0 < c.K / 2; w0 += 16) {
      uint32_t r[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) r[j] = prow[w0 + j];
      tmem_st_x16(tmem_P + ((uint32_t)(warp * 32) << 16) + w0, r);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (tid == 0) {
    mbar_wait(&bar_load, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_f16(128, c.N, 0, c.b_mn_major);
    const int ksteps = c.K / 16;
    for (int k = 0; k < ksteps; ++k) {
      const uint64_t bdesc = make_smem_desc(smem_u32(sB) + k * c.b_kstep_bytes, c.b_lbo, c.b_sbo, c.b_swz);
      if (c.a_from_tmem) {
        mma_ts(tmem_D, tmem_P + k * 8, bdesc, idesc, k > 0);
      } else {
        const int per_atom = c.a_row_bytes / 32;  // K=16 steps per atom
        const int at = k / per_atom, kk = k % per_atom;
        const uint64_t adesc = make_smem_desc(smem_u32(sA) + at * 128 * c.a_row_bytes + kk * 32,
                                              c.a_lbo, c.a_sbo, c.a_swz);
        mma_ss(tmem_D, adesc, bdesc, idesc, k > 0);
      }
    }
    tc_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  for (int n0 = 0; n0 < c.N; n0 += 32) {
    uint32_t r[32];
    tmem_ld_x32(tmem_D + ((uint32_t)(warp * 32) << 16) + n0, r);
    tmem_ld_wait();
    const int row = warp * 32 + lane;
#pragma unroll
    for (int j = 0; j < 32; ++j) out[(size_t)row * c.N + n0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

int main(int argc, char** argv) {
  const int test = argc > 1 ? atoi(argv[1]) : 0;
  const int variant = argc > 2 ? atoi(argv[2]) : 0;
  Cfg c{};
  int a_rows = 128, a_cols, b_rows, b_cols;
  const bool pv = test >= 2;
  const int d = (test == 1 || test == 3 || test == 5) ? 64 : 32;
  if (!pv) {
    c.N = 128; c.K = d; c.a_from_tmem = 0;
    c.a_row_bytes = d * 2; c.a_atoms = 1; c.b_mn_major = 0; c.b_row_bytes = d * 2; c.b_rows = 128;
    c.a_swz = c.b_swz = (d == 32) ? SWZ_64B : SWZ_128B;
    c.a_sbo = c.b_sbo = 8 * d * 2; c.a_lbo = c.b_lbo = 16;
    c.b_kstep_bytes = 32;
    a_cols = d; b_rows = 128; b_cols = d;
  } else {
    c.N = d; c.K = 128; c.a_from_tmem = (test >= 4);
    c.a_row_bytes = 128; c.a_atoms = 2; c.a_swz = SWZ_128B; c.a_sbo = 1024; c.a_lbo = 16;
    c.b_mn_major = 1; c.b_row_bytes = d * 2; c.b_rows = 128;
    c.b_swz = (d == 32) ? SWZ_64B : SWZ_128B;
    c.b_sbo = 8 * d * 2; c.b_lbo = 8 * d * 2;
    c.b_kstep_bytes = 16 * d * 2;
    a_cols = 128; b_rows = 128; b_cols = d;
  }
  // descriptor variants for the less certain encodings
  if (variant == 1) { c.b_lbo = 16; }
  if (variant == 2) { c.a_lbo = 0; c.b_lbo = 0; }
  if (variant == 3) { c.b_sbo = 16 * d * 2; }
  if (variant == 4) { c.b_lbo = 128 * d * 2; }

  std::vector<__half> hA((size_t)a_rows * a_cols), hB((size_t)b_rows * b_cols);
  srand(1234 + test);
  auto rnd = []() { return (float)((rand() % 2001) - 1000) / 1000.0f; };
  for (auto& v : hA) v = __float2half(rnd());
  for (auto& v : hB) v = __float2half(rnd());
  std::vector<float> ref((size_t)128 * c.N, 0.f);
  for (int i = 0; i < 128; ++i)
    for (int j = 0; j < c.N; ++j) {
      double s = 0;
      for (int k = 0; k < c.K; ++k) {
        const float a = __half2float(hA[(size_t)i * a_cols + k]);
        const float b = pv ? __half2float(hB[(size_t)k * b_cols + j]) : __half2float(hB[(size_t)j * b_cols + k]);
        s += (double)a * b;
      }
      ref[(size_t)i * c.N + j] = (float)s;
    }
  __half *dA, *dB;
  float* dO;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dO, ref.size() * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dO, 0, ref.size() * 4));
  const CUtensorMapSwizzle swA = (c.a_swz == SWZ_64B) ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  const CUtensorMapSwizzle swB = (c.b_swz == SWZ_64B) ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  CUtensorMap mA = make_map_2d(dA, a_rows, a_cols, 128, c.a_row_bytes / 2, swA);
  CUtensorMap mB = make_map_2d(dB, b_rows, b_cols, c.b_rows, c.b_row_bytes / 2, swB);
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152 + 1024));
  probe_kernel<<<1, 128, 49152 + 1024>>>(mA, mB, dA, dO, c);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> out(ref.size());
  CK(cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (size_t i = 0; i < ref.size(); ++i) {
    maxerr = fmax(maxerr, fabs((double)out[i] - ref[i]));
    maxref = fmax(maxref, fabs((double)ref[i]));
  }
  printf("test %d variant %d: max_abs_err %.6f (max |ref| %.3f) %s\n", test, variant, maxerr, maxref,
         maxerr < 2e-2 ? "PASS" : "FAIL");
  return maxerr < 2e-2 ? 0 : 1;
}
