// gemm.cu -- tcgen05 / TMA GEMM with fused epilogues:  D[M,N] = A[M,K] * W[N,K]^T (+ ...)
//
// Replaces every nn.Linear on the DiT / motion-VAE path that the reference runs as a cuBLAS
// fp16 GEMM followed by separate bias / GELU / gate / residual elementwise kernels
// (reference model/dit.py:128-138,240-277; model/attention/modules.py:98-146;
// model/autoencoder.py:90-163).  fp16 operands, fp32 accumulation in tensor memory.
//
// Structure (one 128 x BN output tile per CTA, 192 threads):
//   warp 0    TMA producer: A tile [128 x 64] and W tile [BN x 64] per k-block, SWIZZLE_128B,
//             kStages-deep mbarrier ring
//   warp 1    single-thread tcgen05.mma issuer (kind::f16, M=128, N=BN, K=16 x4 per k-block),
//             tcgen05.commit releases the smem stage / publishes the accumulator
//   warps 2-5 epilogue: tcgen05.ld (thread = one row, 32 columns at a time) -> fused
//             bias / GELU-tanh / gate / residual -> global
// Two CTAs fit per SM (3 x 32 KB stages, 128 TMEM columns each), so one CTA's epilogue
// overlaps the other's main loop.
#include "../../include/gvf_b200.h"
#include "tc_common.cuh"
#include "tma_host.h"

namespace gvf {
using namespace tc;

constexpr int kBM = 128, kBK = 64;

struct GemmEpi {
  int mode;
  const float* bias;        // [N] or null
  void* out;                // fp16 [M,N] (modes 0,1,3) or fp32 [M,N] (modes 2,4,5)
  const __half* gate;       // mode 2: [batches, gate_stride] fp16-valued gates or null
  int gate_stride;          // elements between batches in gate
  int rows_per_batch;       // mode 2: row -> batch index
  int ldo;                  // leading dimension of out (elements)
};

__device__ __forceinline__ float gelu_tanh(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  const float inner = k0 * (x + k1 * x * x * x);
  return 0.5f * x * (1.0f + tanhf(inner));
}
__device__ __forceinline__ float r16(float x) { return __half2float(__float2half_rn(x)); }

template <int BN, int STAGES>
__global__ void __launch_bounds__(192, 2)
gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW,
            int M, int N, int K, GemmEpi ep) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], accum_bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int A_BYTES = kBM * kBK * 2, W_BYTES = BN * kBK * 2, STAGE_BYTES = A_BYTES + W_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_n = blockIdx.x, tile_m = blockIdx.y;
  const int kblocks = (K + kBK - 1) / kBK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&accum_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapW);
  }
  if (warp == 2) {
    tmem_alloc(&tmem_base_s, BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < kblocks; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
        uint8_t* st = smem + s * STAGE_BYTES;
        tma_load_2d(st, &mapA, &full_bar[s], kb * kBK, tile_m * kBM);
        tma_load_2d(st + A_BYTES, &mapW, &full_bar[s], kb * kBK, tile_n * BN);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(kBM, BN, 0, 0);
      for (int kb = 0; kb < kblocks; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a0 = smem_u32(smem + s * STAGE_BYTES), w0 = a0 + A_BYTES;
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) {
          const uint64_t ad = make_smem_desc(a0 + k * 32, 16, 1024, SWZ_128B);
          const uint64_t wd = make_smem_desc(w0 + k * 32, 16, 1024, SWZ_128B);
          mma_ss(tmem, ad, wd, idesc, (kb | k) != 0);
        }
        tc_commit(&empty_bar[s]);
      }
      tc_commit(&accum_bar);
    }
  } else {
    // epilogue: warp (2..5) owns TMEM lanes 32*(warp%4) .. +31
    const int q = warp & 3;
    const int row = tile_m * kBM + q * 32 + lane;
    mbar_wait(&accum_bar, 0);
    tc_fence_after();
    const bool row_ok = row < M;
    const int b = (ep.mode == 2 && ep.gate) ? row / ep.rows_per_batch : 0;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      __syncwarp();   // tcgen05.ld is .sync.aligned: reconverge after the divergent stores
      tmem_ld_x32(tmem + ((uint32_t)(q * 32) << 16) + c0, r);
      tmem_ld_wait();
      const int col0 = tile_n * BN + c0;
      if (!row_ok || col0 >= N) continue;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] = __uint_as_float(r[j]);
        if (ep.bias) v[j] += (col0 + j < N) ? __ldg(ep.bias + col0 + j) : 0.f;
      }
      const int ncol = min(32, N - col0);   // N % 8 == 0
      if (ep.mode == 0 || ep.mode == 1) {
        __half* o = reinterpret_cast<__half*>(ep.out) + (size_t)row * ep.ldo + col0;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          if (j < ncol) {
            __align__(16) __half h[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              float x = v[j + t];
              if (ep.mode == 1) x = gelu_tanh(r16(x));
              h[t] = __float2half_rn(x);
            }
            *reinterpret_cast<uint4*>(o + j) = *reinterpret_cast<uint4*>(h);
          }
        }
      } else if (ep.mode == 2) {
        float* o = reinterpret_cast<float*>(ep.out) + (size_t)row * ep.ldo + col0;
        const __half* g = ep.gate ? ep.gate + (size_t)b * ep.gate_stride + col0 : nullptr;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          if (j < ncol) {
            float4 x = *reinterpret_cast<float4*>(o + j);
            float h0 = r16(v[j]), h1 = r16(v[j + 1]), h2 = r16(v[j + 2]), h3 = r16(v[j + 3]);
            if (g) {
              h0 = r16(h0 * __half2float(g[j]));
              h1 = r16(h1 * __half2float(g[j + 1]));
              h2 = r16(h2 * __half2float(g[j + 2]));
              h3 = r16(h3 * __half2float(g[j + 3]));
            }
            x.x += h0; x.y += h1; x.z += h2; x.w += h3;
            *reinterpret_cast<float4*>(o + j) = x;
          }
        }
      } else if (ep.mode == 3) {
        __half* o = reinterpret_cast<__half*>(ep.out) + (size_t)row * ep.ldo + col0;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          if (j < ncol) {
            uint4 old = *reinterpret_cast<uint4*>(o + j);
            __half* oh = reinterpret_cast<__half*>(&old);
            __align__(16) __half h[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) h[t] = __float2half_rn(r16(v[j + t]) + __half2float(oh[t]));
            *reinterpret_cast<uint4*>(o + j) = *reinterpret_cast<uint4*>(h);
          }
        }
      } else if (ep.mode == 4) {
        float* o = reinterpret_cast<float*>(ep.out) + (size_t)row * ep.ldo + col0;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          if (j < ncol) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      } else {
        // mode 5: compact fp32 rows of ldo (<= N) fp16-rounded values, e.g. Linear(768 -> 14)
        float* o = reinterpret_cast<float*>(ep.out) + (size_t)row * ep.ldo;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + j < ep.ldo) o[col0 + j] = r16(v[j]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, BN);
}

template <int BN, int STAGES>
static int launch_gemm(const CUtensorMap& mA, const CUtensorMap& mW, int M, int N, int K,
                       const GemmEpi& ep, cudaStream_t st) {
  constexpr int SMEM = STAGES * (kBM * kBK * 2 + BN * kBK * 2) + 1024;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(gemm_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             SMEM) != cudaSuccess)
      return GVF_ERR_CUDA;
    configured = true;
  }
  dim3 grid((N + BN - 1) / BN, (M + kBM - 1) / kBM);
  gemm_kernel<BN, STAGES><<<grid, 192, SMEM, st>>>(mA, mW, M, N, K, ep);
  return cudaGetLastError() == cudaSuccess ? GVF_OK : GVF_ERR_CUDA;
}

}  // namespace gvf

using namespace gvf;

extern "C" GVF_API int gvf_gemm_f16(const void* A, int lda, const void* W, int ldw, int M, int N,
                                    int K, int epilogue, const float* bias, void* out, int ldo,
                                    const void* gate, int gate_stride, int rows_per_batch,
                                    void* stream) {
  if (!A || !W || !out || M <= 0 || N <= 0 || K <= 0) return GVF_ERR_INVALID;
  if ((N % 8) || (K % 8) || (lda % 8) || (ldw % 8) || epilogue < 0 || epilogue > 5) return GVF_ERR_INVALID;
  if (((uintptr_t)A | (uintptr_t)W | (uintptr_t)out) & 15) return GVF_ERR_INVALID;
  if (epilogue == 2 && gate && rows_per_batch <= 0) return GVF_ERR_INVALID;
  const bool wide = (N % 128) == 0 && N >= 1024;   // BN = 128 everywhere for now; see DESIGN.md
  (void)wide;
  constexpr int BN = 128;
  CUtensorMap mA, mW;
  const uint64_t dA[2] = {(uint64_t)K, (uint64_t)M}, sA[2] = {1, (uint64_t)lda};
  const uint64_t dW[2] = {(uint64_t)K, (uint64_t)N}, sW[2] = {1, (uint64_t)ldw};
  const uint32_t bA[2] = {kBK, kBM}, bW[2] = {kBK, BN};
  if (!make_tmap_f16(&mA, A, 2, dA, sA, bA, CU_TENSOR_MAP_SWIZZLE_128B)) return GVF_ERR_CUDA;
  if (!make_tmap_f16(&mW, W, 2, dW, sW, bW, CU_TENSOR_MAP_SWIZZLE_128B)) return GVF_ERR_CUDA;
  GemmEpi ep;
  ep.mode = epilogue; ep.bias = bias; ep.out = out; ep.gate = (const __half*)gate;
  ep.gate_stride = gate_stride; ep.rows_per_batch = rows_per_batch > 0 ? rows_per_batch : 1;
  ep.ldo = ldo;
  return launch_gemm<BN, 3>(mA, mW, M, N, K, ep, (cudaStream_t)stream);
}
