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
#include "launch.h"

namespace gvf {
using namespace tc;

int g_pdl_enabled = 0;   // measured on B200: 296.3 ms / object with the attribute, 290.3 ms without (see launch.h)

constexpr int kBM = 128, kBK = 64;

struct GemmEpi {
  int mode;
  const float* bias;        // [N] or null
  void* out;                // fp16 [M,N] (modes 0,1,3,6) or fp32 [M,N] (modes 2,4,5)
  const __half* gate;       // mode 2: [batches, gate_stride] fp16-valued gates or null
  int gate_stride;          // elements between batches in gate
  int rows_per_batch;       // mode 2: row -> batch index
  int ldo;                  // leading dimension of out (elements)
  const float* gamma_q;     // mode 6: per-head RMS-norm gains [H, 32] for columns [0, norm_cols/2)
  const float* gamma_k;     //         and [norm_cols/2, norm_cols)
  int norm_cols;            // mode 6: columns below this are RMS-normalised per 32-wide head
  // GATHER kernels (submanifold sparse convolution): row m of the A operand in k-block kb is row
  // gather_idx[m * gather_k3 + kb / gather_cb] of the tensor behind mapA (negative = absent voxel = zeros), columns
  // [64 (kb % gather_cb), +64)
  const int* gather_idx = nullptr;
  int gather_k3 = 0, gather_cb = 1;
};

__device__ __forceinline__ float gelu_tanh(float x) {
  // 0.5 x (1 + tanh(u)), tanh(u) = 1 - 2 / (exp(2u) + 1): two MUFU ops, ~1e-6 accurate
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  const float u = k0 * (x + k1 * x * x * x);
  const float t = 1.0f - __fdividef(2.0f, __expf(2.0f * u) + 1.0f);
  return 0.5f * x * (1.0f + t);
}
__device__ __forceinline__ float r16(float x) { return __half2float(__float2half_rn(x)); }
// One MUFU instead of two: tanh.approx.f32 (relative error 2^-11, i.e. at most half an fp16 ulp of the
// result, which is rounded to fp16 right after).  The 128 x 256 GELU tile costs 2 x 128 x 256 / 16 = 4096
// MUFU clocks with exp + reciprocal -- as long as its whole main loop at K = 512.
__device__ __forceinline__ float gelu_tanh_fast(float x) {
  const float k0 = 0.7978845608028654f, k0k1 = 0.7978845608028654f * 0.044715f;
  const float u = x * fmaf(k0k1, x * x, k0);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  const float h = 0.5f * x;
  return fmaf(h, t, h);
}

// d/dx of the above with the same tanh.approx: 0.5 (1 + t) + 0.5 x (1 - t^2) k0 (1 + 3 k1 x^2)
__device__ __forceinline__ float gelu_tanh_grad_fast(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  const float x2 = x * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x * fmaf(k0 * k1, x2, k0)));
  return fmaf(0.5f * x * (1.0f - t * t), k0 * fmaf(3.0f * k1, x2, 1.0f), 0.5f * (1.0f + t));
}

constexpr int kEpiCols = 64;                 // columns per epilogue chunk (default)
constexpr int kStgLd = kEpiCols + 1;         // padded row of the per-warp staging tile (floats)

// Epilogue of one 128 x BN accumulator tile, executed by the four epilogue warps (q = warp % 4 owns
// TMEM lanes 32q..32q+31).
template <int BN, int EC = kEpiCols>
__device__ __forceinline__ void epilogue_tile(const GemmEpi& ep, const uint32_t tmem, float* stg,
                                              const float* s_bias, const int q, const int lane,
                                              const int tile_m, const int tile_n, const int M, const int N) {
    // ---------------------------------------------------------------- epilogue (warps 2..5)
    // Phase 1 (thread = row): tcgen05.ld 64 columns, bias + row-wise math (GELU / fp16 rounding /
    // per-head RMS norm), park the values in a per-warp staging tile (the pipeline smem is free
    // once accum_bar fires).  Phase 2 (lane = column pair): walk the 32 rows so that every global
    // access of the warp is one contiguous 128 B (fp16) / 256 B (fp32) row segment.
    const int row0 = tile_m * kBM + q * 32;
    const int mode = ep.mode;
    const int rows_here = min(32, M - row0);
    const int b_first = row0 / ep.rows_per_batch;
    const bool one_batch = rows_here <= 0 || (row0 + rows_here - 1) / ep.rows_per_batch == b_first;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += EC) {
      const int col0 = tile_n * BN + c0;
      float v[EC];
      __syncwarp();
      {
        uint32_t rr[EC];
#pragma unroll
        for (int j0 = 0; j0 < EC; j0 += 32)
          tmem_ld_x32(tmem + ((uint32_t)(q * 32) << 16) + c0 + j0, *reinterpret_cast<uint32_t(*)[32]>(&rr[j0]));
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < EC; ++j) v[j] = __uint_as_float(rr[j]) + s_bias[c0 + j];
      }
      if (col0 >= N || rows_here <= 0) continue;       // warp-uniform
      if (mode == 1) {
#pragma unroll
        for (int j = 0; j < EC; ++j) v[j] = gelu_tanh(r16(v[j]));
      } else if (mode == 6) {
        // MultiHeadRMSNorm on the fp16-rounded Linear output, one head = 32 columns
#pragma unroll
        for (int hh = 0; hh < EC / 32; ++hh) {
          const int hc = col0 + hh * 32;
          if (hc < ep.norm_cols) {
            float ss = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) { v[hh * 32 + j] = r16(v[hh * 32 + j]); ss += v[hh * 32 + j] * v[hh * 32 + j]; }
            const float inv = 5.656854249492381f / fmaxf(sqrtf(ss), 1e-12f);   // sqrt(32) / max(||x||, eps)
            const int half_cols = ep.norm_cols >> 1;
            const float* g = (hc < half_cols) ? ep.gamma_q + hc : ep.gamma_k + (hc - half_cols);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[hh * 32 + j] = v[hh * 32 + j] * inv * __ldg(g + j);
          }
        }
      } else if (mode == 2 || mode == 3 || mode == 5) {
#pragma unroll
        for (int j = 0; j < EC; ++j) v[j] = r16(v[j]);
      }
#pragma unroll
      for (int j = 0; j < EC; ++j) stg[lane * (EC + 1) + j] = v[j];
      __syncwarp();
      // ---- phase 2: lane owns columns (2*lane, 2*lane+1) of the chunk
      // lane -> (column pair lc, row phase ro_): EC == 64 one row per pass, EC == 32 two rows per pass
      constexpr int RS = 64 / EC;
      const int lc = (2 * lane) % EC, ro_ = (2 * lane) / EC;
      const int cc = col0 + lc;
      const bool col_ok = cc < N;                       // N % 8 == 0, so the pair is in or out together
      if (mode == 0 || mode == 1 || mode == 6) {
        __half* o = reinterpret_cast<__half*>(ep.out) + (size_t)row0 * ep.ldo + cc;
        if (col_ok)
#pragma unroll 8
          for (int r = ro_; r < rows_here; r += RS)
            *reinterpret_cast<__half2*>(o + (size_t)r * ep.ldo) =
                __floats2half2_rn(stg[r * (EC + 1) + lc], stg[r * (EC + 1) + lc + 1]);
      } else if (mode == 2) {
        // fp32 residual read-modify-write; loads of 16 rows are issued before their stores so the
        // DRAM latency is paid once per batch, not once per row
        float* o = reinterpret_cast<float*>(ep.out) + (size_t)row0 * ep.ldo + cc;
        if (col_ok) {
          float g0 = 1.f, g1 = 1.f;
          if (ep.gate && one_batch) {
            const __half2 gg = *reinterpret_cast<const __half2*>(ep.gate + (size_t)b_first * ep.gate_stride + cc);
            g0 = __low2float(gg); g1 = __high2float(gg);
          }
          for (int rb = 0; rb < rows_here; rb += 16 * RS) {
            float2 x[16];
#pragma unroll
            for (int k = 0; k < 16; ++k)
              if (rb + k * RS + ro_ < rows_here)
                x[k] = *reinterpret_cast<const float2*>(o + (size_t)(rb + k * RS + ro_) * ep.ldo);
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              const int r = rb + k * RS + ro_;
              if (r < rows_here) {
                if (ep.gate && !one_batch) {
                  const __half2 gg = *reinterpret_cast<const __half2*>(
                      ep.gate + (size_t)((row0 + r) / ep.rows_per_batch) * ep.gate_stride + cc);
                  g0 = __low2float(gg); g1 = __high2float(gg);
                }
                float h0 = stg[r * (EC + 1) + lc], h1 = stg[r * (EC + 1) + lc + 1];
                if (ep.gate) { h0 = r16(h0 * g0); h1 = r16(h1 * g1); }
                x[k].x += h0; x[k].y += h1;
                *reinterpret_cast<float2*>(o + (size_t)r * ep.ldo) = x[k];
              }
            }
          }
        }
      } else if (mode == 3) {
        __half* o = reinterpret_cast<__half*>(ep.out) + (size_t)row0 * ep.ldo + cc;
        if (col_ok)
          for (int rb = 0; rb < rows_here; rb += 8 * RS) {
            __half2 old[8];
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (rb + k * RS + ro_ < rows_here)
                old[k] = *reinterpret_cast<const __half2*>(o + (size_t)(rb + k * RS + ro_) * ep.ldo);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int r = rb + k * RS + ro_;
              if (r < rows_here)
                *reinterpret_cast<__half2*>(o + (size_t)r * ep.ldo) =
                    __floats2half2_rn(stg[r * (EC + 1) + lc] + __low2float(old[k]),
                                      stg[r * (EC + 1) + lc + 1] + __high2float(old[k]));
            }
          }
      } else if (mode == 4) {
        float* o = reinterpret_cast<float*>(ep.out) + (size_t)row0 * ep.ldo + cc;
        if (col_ok)
#pragma unroll 8
          for (int r = ro_; r < rows_here; r += RS)
            *reinterpret_cast<float2*>(o + (size_t)r * ep.ldo) =
                make_float2(stg[r * (EC + 1) + lc], stg[r * (EC + 1) + lc + 1]);
      } else {
        // mode 5: compact fp32 rows of ldo (<= N) columns, e.g. Linear(768 -> 14)
        float* o = reinterpret_cast<float*>(ep.out) + (size_t)row0 * ep.ldo;
        for (int r = ro_; r < rows_here; r += RS)
#pragma unroll
          for (int t = 0; t < 2; ++t)
            if (cc + t < ep.ldo) o[(size_t)r * ep.ldo + cc + t] = stg[r * (EC + 1) + lc + t];
      }
    }
}

template <int BN, int STAGES, int EC = kEpiCols, int MINB = 2>
__global__ void __launch_bounds__(192, MINB)
gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW,
            int M, int N, int K, GemmEpi ep) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], accum_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_bias[BN];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int A_BYTES = kBM * kBK * 2, W_BYTES = BN * kBK * 2, STAGE_BYTES = A_BYTES + W_BYTES;
  static_assert(4 * 32 * (EC + 1) * 4 <= STAGES * STAGE_BYTES, "epilogue staging must fit in the pipeline smem");
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
  if (threadIdx.x >= 64 && threadIdx.x < 64 + BN) {   // bias of this tile's columns -> smem (epilogue warps)
    const int c = tile_n * BN + (threadIdx.x - 64);
    s_bias[threadIdx.x - 64] = (ep.bias && c < N) ? __ldg(ep.bias + c) : 0.f;
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
    const int q = warp & 3;
    mbar_wait(&accum_bar, 0);
    tc_fence_after();
    epilogue_tile<BN, EC>(ep, tmem, reinterpret_cast<float*>(smem) + q * 32 * (EC + 1), s_bias, q, lane, tile_m, tile_n, M, N);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, BN);
}

// ---------------------------------------------------------------------------------------
// Persistent variant: one CTA per SM walks the output tiles (n fastest, so CTAs running together
// share A rows in L2); the accumulator is double buffered in TMEM (2 x BN columns) so the epilogue of
// tile i overlaps the TMA / MMA main loop of tile i+1.
template <int BN, int STAGES>
__global__ void __launch_bounds__(192, 1)
gemm_persistent_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW,
                       int M, int N, int K, GemmEpi ep) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_bias[BN];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int A_BYTES = kBM * kBK * 2, W_BYTES = BN * kBK * 2, STAGE_BYTES = A_BYTES + W_BYTES;
  float* stg_base = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kblocks = (K + kBK - 1) / kBK;
  const int tiles_n = (N + BN - 1) / BN, tiles_m = (M + kBM - 1) / kBM;
  const int num_tiles = tiles_n * tiles_m;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 128); }
    fence_barrier_init();
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapW);
  }
  if (warp == 2) {
    tmem_alloc(&tmem_base_s, 2 * BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int tile_m = tile / tiles_n, tile_n = tile - tile_m * tiles_n;
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
          mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
          uint8_t* st = smem + s * STAGE_BYTES;
          tma_load_2d(st, &mapA, &full_bar[s], kb * kBK, tile_m * kBM);
          tma_load_2d(st + A_BYTES, &mapW, &full_bar[s], kb * kBK, tile_n * BN);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(kBM, BN, 0, 0);
      int it = 0, lt = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
        const int acc = lt & 1;
        mbar_wait(&tempty_bar[acc], ((lt >> 1) & 1) ^ 1);       // epilogue has drained this accumulator
        tc_fence_after();
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&full_bar[s], (it / STAGES) & 1);
          tc_fence_after();
          const uint32_t a0 = smem_u32(smem + s * STAGE_BYTES), w0 = a0 + A_BYTES;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            mma_ss(tmem + acc * BN, make_smem_desc(a0 + k * 32, 16, 1024, SWZ_128B),
                   make_smem_desc(w0 + k * 32, 16, 1024, SWZ_128B), idesc, (kb | k) != 0);
          tc_commit(&empty_bar[s]);
        }
        tc_commit(&tfull_bar[acc]);
      }
    }
  } else {
    const int q = warp & 3;
    const int et = threadIdx.x - 64;                             // 0..127 among the epilogue threads
    int lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      const int tile_m = tile / tiles_n, tile_n = tile - tile_m * tiles_n;
      const int acc = lt & 1;
      asm volatile("bar.sync 1, 128;");                          // previous tile's s_bias readers are done
      for (int c = et; c < BN; c += 128) {
        const int col = tile_n * BN + c;
        s_bias[c] = (ep.bias && col < N) ? __ldg(ep.bias + col) : 0.f;
      }
      asm volatile("bar.sync 1, 128;");
      mbar_wait(&tfull_bar[acc], (lt >> 1) & 1);
      tc_fence_after();
      epilogue_tile<BN>(ep, tmem + acc * BN, stg_base + q * 32 * kStgLd, s_bias, q, lane, tile_m, tile_n, M, N);
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 2 * BN);
}

// ---------------------------------------------------------------------------------------
// Persistent kernel, generation 2: the epilogue no longer bounds the short-K shapes.
// Measured on the generation above (tools/gemm_bench.py): qkv N=1536 K=512 ran 40 us against 8 us of
// tensor time -- four epilogue warps, one per SM sub-partition, walking every accumulator element through a
// scalar shared-memory transpose (~2000 dependent instructions per thread and 128 x 256 tile).  Here
//   * eight epilogue warps (two per TMEM lane quarter, each owning half of the tile's columns),
//   * thread = row keeps its values in registers, packs them and writes 16 B pieces into a SWIZZLE_128B
//     staging tile (32 rows x 128 B per warp, two buffers),
//   * one lane per warp hands the tile to the TMA unit (cp.async.bulk.tensor store, bulk groups), which
//     also clips the M / N edges,
//   * the fp32 / fp16 residual of the read-modify-write epilogues is fetched with 16 B loads issued before
//     the accumulator wait, so its DRAM latency overlaps the main loop of the tile.
// Same rounding points as epilogue_tile() above.  320 threads: warp 0 TMA producer, warp 1 MMA issuer,
// warps 2-9 epilogue; accumulators double buffered in TMEM (2 x BN columns).
// TRANS (NCTA = 1 only): both operands are MN-major -- out[m, n] = sum_r A[r, m] W[r, n] with A [R, M] and W [R, N]
// row-major, i.e. the weight gradient dW = dY^T X straight from the activation tensors (no transposed copies).  A
// stage holds 64 reduction rows: 64-column boxes (128 B inner extent, SWIZZLE_128B) of 8 KB each, two for A and
// BN / 64 for W; the shared-memory descriptors walk them with LBO = 8 KB between the 64-wide column blocks and
// SBO = 1 KB between 8-row groups (canonical MN-major layout), 2 KB per 16-row k-step.
// GATHER (NCTA = 1 only): the A operand is never materialised -- the producer warp fetches the 128 rows of a stage
// with 32 `cp.async.bulk.tensor.2d.tile::gather4` loads (lane l: rows 4 l .. 4 l + 3 of the tile, row indices read
// from the neighbour map, out-of-range = zero-filled), each landing as four 128 B rows exactly where the one-box load
// of the dense kernel would have put them (TMA swizzling is a function of the shared-memory address).
// TRANSB (NCTA = 1 only): only W is MN-major -- out[m, n] = sum_r A[m, r] W[r, n] with A [M, R] row-major (K-major, as in
// the plain kernel) and W [R, N] row-major: the input gradient dX = dY W of a Linear straight from its [out, in] weight, no
// transposed weight copy.  A is staged as one 128 x 64 box, W as BN / 64 boxes of 64 reduction rows like TRANS does.
template <int BN, int STAGES, int MODE, int NCTA, bool TRANS = false, bool GATHER = false, bool TRANSB = false>
__global__ void __launch_bounds__(320, 1)
gemm_ws_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW,
               const __grid_constant__ CUtensorMap mapO, int M, int N, int K, GemmEpi ep, int ksplit) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], tfull_bar[2], tempty_bar[2];
  __shared__ uint64_t xres_bar[8][2];     // mode 2, BN = 128: residual chunks arriving by TMA (one per warp and buffer)
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_bias[2][BN];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // NCTA = 2: a CTA pair computes a 256 x BN tile with tcgen05.mma.cta_group::2 -- this CTA stages its own 128
  // rows of A and BN / 2 rows of W (the operand bytes entering each SM drop by a third at BN = 256), the leader
  // (cluster rank 0) issues the MMAs for both, each CTA runs the epilogue of its own 128 accumulator rows.
  constexpr int WROWS = BN / NCTA;
  constexpr int A_BYTES = kBM * kBK * 2, W_BYTES = WROWS * kBK * 2, STAGE_BYTES = A_BYTES + W_BYTES;
  const uint32_t rank = (NCTA == 2) ? cluster_ctarank() : 0u;
  uint8_t* stg_base = smem + STAGES * STAGE_BYTES;              // 8 warps x 2 buffers x 4 KB, 1 KB aligned
  const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
  const int kblocks = (K + kBK - 1) / kBK;
  const int tiles_n = (N + BN - 1) / BN, tiles_m = (M + NCTA * kBM - 1) / (NCTA * kBM);   // pair tiles for NCTA = 2
  // Split-K (wgrad shapes of the training step: few output tiles, reduction length 12288 ..): a work item is
  // (output tile, k range); the partial tiles are summed in global memory by TMA reduce-add stores (mode 4 only, the
  // host zeroes the output first).  ksplit = 1: one item per tile, plain stores.
  const int out_tiles = tiles_n * tiles_m;
  const int num_tiles = out_tiles * ksplit;           // work items
  const int kb_per = (kblocks + ksplit - 1) / ksplit;
  const int first_tile = blockIdx.x / NCTA, tile_step = gridDim.x / NCTA;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 8 * NCTA); }
    for (int i = 0; i < 16; ++i) mbar_init(&xres_bar[i >> 1][i & 1], 1);
    fence_barrier_init();
    tma_prefetch_desc(&mapO);
    if constexpr (NCTA == 1 && !GATHER) {
      // first stages of the first tile: requested by the thread that has just created the barriers, so their
      // latency overlaps the TMEM allocation and the CTA barrier (pairs must wait for the cluster barrier first)
      pdl_wait();
      if (first_tile < num_tiles) {
        const int ot = first_tile % out_tiles, kb0 = (first_tile / out_tiles) * kb_per;
        const int nkb = min(kblocks, kb0 + kb_per) - kb0;
        const int tile_m = ot / tiles_n, tile_n = ot % tiles_n;
        for (int kb = 0; kb < (nkb < STAGES ? nkb : STAGES); ++kb) {
          mbar_arrive_expect_tx(&full_bar[kb], STAGE_BYTES);
          uint8_t* st = smem + kb * STAGE_BYTES;
          if constexpr (TRANS) {
            for (int b = 0; b < kBM / 64; ++b) tma_load_2d(st + b * 8192, &mapA, &full_bar[kb], tile_m * kBM + b * 64, (kb0 + kb) * kBK);
          } else {
            tma_load_2d(st, &mapA, &full_bar[kb], (kb0 + kb) * kBK, tile_m * kBM);
          }
          if constexpr (TRANS || TRANSB) {
            for (int b = 0; b < BN / 64; ++b) tma_load_2d(st + A_BYTES + b * 8192, &mapW, &full_bar[kb], tile_n * BN + b * 64, (kb0 + kb) * kBK);
          } else {
            tma_load_2d(st + A_BYTES, &mapW, &full_bar[kb], (kb0 + kb) * kBK, tile_n * BN);
          }
        }
      }
    } else {
      tma_prefetch_desc(&mapA);
      tma_prefetch_desc(&mapW);
    }
  }
  if (warp == 2) {
    if constexpr (NCTA == 2) { tmem_alloc_2cta(&tmem_base_s, 2 * BN); tmem_relinquish_2cta(); }
    else { tmem_alloc(&tmem_base_s, 2 * BN); tmem_relinquish(); }
  }
  tc_fence_before();
  if constexpr (NCTA == 2) cluster_sync_all();      // the peer's barriers exist before anything signals them
  else __syncthreads();
  tc_fence_after();
  pdl_wait();                       // everything above is independent of the previous kernel's output
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    // TMA producer: all 32 lanes run the (warp-uniform) loop, one elected lane issues -- see elect_one()
    int s = 0;
    uint32_t ph = 0;                               // parity of full / empty for the current pass over the ring
    int skip = 0;
    if (NCTA == 1 && !GATHER && first_tile < num_tiles) {                    // requested in the prologue
      const int kb0 = (first_tile / out_tiles) * kb_per, nkb = min(kblocks, kb0 + kb_per) - kb0;
      skip = nkb < STAGES ? nkb : STAGES;
    }
    for (int item = first_tile; item < num_tiles; item += tile_step) {
      const int tile = item % out_tiles, kb0 = (item / out_tiles) * kb_per, kb1 = min(kblocks, kb0 + kb_per);
      const int tile_m = (tile / tiles_n) * NCTA + (int)rank, tile_n = tile % tiles_n;
      for (int kb = kb0; kb < kb1; ++kb) {
        if (skip > 0) --skip;
        else if constexpr (GATHER) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* st = smem + s * STAGE_BYTES;
          const int kk = kb / ep.gather_cb, cc = (kb - kk * ep.gather_cb) * kBK;
          const int lane_p = threadIdx.x & 31;
          int ridx[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int r = tile_m * kBM + 4 * lane_p + j;
            ridx[j] = r < M ? __ldg(ep.gather_idx + (size_t)r * ep.gather_k3 + kk) : -1;
          }
          if (elect_one()) {
            mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
            tma_load_2d(st + A_BYTES, &mapW, &full_bar[s], kb * kBK, tile_n * BN);
          }
          __syncwarp();
          tma_gather4_2d(st + lane_p * 512, &mapA, &full_bar[s], cc, ridx[0], ridx[1], ridx[2], ridx[3]);
        } else {
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (elect_one()) {
            uint8_t* st = smem + s * STAGE_BYTES;
            if constexpr (NCTA == 2) {
              // both CTAs' bytes are credited to the leader's full barrier, which expects the sum
              if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * STAGE_BYTES);
              tma_load_2d_2cta(st, &mapA, &full_bar[s], kb * kBK, tile_m * kBM);
              tma_load_2d_2cta(st + A_BYTES, &mapW, &full_bar[s], kb * kBK, tile_n * BN + (int)rank * WROWS);
            } else if constexpr (TRANS || TRANSB) {
              mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
              if constexpr (TRANS) {
                for (int b = 0; b < kBM / 64; ++b) tma_load_2d(st + b * 8192, &mapA, &full_bar[s], tile_m * kBM + b * 64, kb * kBK);
              } else {
                tma_load_2d(st, &mapA, &full_bar[s], kb * kBK, tile_m * kBM);
              }
              for (int b = 0; b < BN / 64; ++b) tma_load_2d(st + A_BYTES + b * 8192, &mapW, &full_bar[s], tile_n * BN + b * 64, kb * kBK);
            } else {
              mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
              tma_load_2d(st, &mapA, &full_bar[s], kb * kBK, tile_m * kBM);
              tma_load_2d(st + A_BYTES, &mapW, &full_bar[s], kb * kBK, tile_n * BN);
            }
          }
          __syncwarp();
        }
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // MMA issuer (leader CTA of a pair): warp-uniform loop; the descriptors' high words are constants, the low
    // words (address >> 4 | LBO field) advance by adds; only tcgen05.mma / commit sit behind elect.sync
    if (rank == 0) {
      constexpr bool TA = TRANS, TB = TRANS || TRANSB;
      const uint32_t idesc = make_idesc_f16(kBM * NCTA, BN, TA ? 1 : 0, TB ? 1 : 0);
      constexpr uint32_t LBO = TA ? 8192u : 16u;             // MN-major: between 64-column blocks; K-major: unused
      constexpr uint32_t LBO_W = TB ? 8192u : 16u;
      constexpr uint32_t KSTEP16 = TA ? 128u : 2u;           // descriptor advance per 16-deep k-step (16 B units)
      constexpr uint32_t KSTEP16_W = TB ? 128u : 2u;
      const uint32_t d_hi = (uint32_t)(make_smem_desc(0, LBO, 1024, SWZ_128B) >> 32);      // SBO / swizzle: same for both
      const uint32_t a_lo0 = (uint32_t)make_smem_desc(smem_u32(smem), LBO, 1024, SWZ_128B);
      // W's low word relative to A's: its address offset plus the difference of the LBO fields (bits 16..29)
      constexpr uint32_t STAGE16 = STAGE_BYTES >> 4, A16 = (A_BYTES >> 4) + (((LBO_W >> 4) - (LBO >> 4)) << 16);
      const uint32_t b_full = smem_u32(&full_bar[0]), b_empty = smem_u32(&empty_bar[0]);
      int s = 0, lt = 0;
      uint32_t ph = 0, s_lo = a_lo0;
      for (int item = first_tile; item < num_tiles; item += tile_step, ++lt) {
        const int kb0 = (item / out_tiles) * kb_per, kb1 = min(kblocks, kb0 + kb_per);
        const int acc = lt & 1;
        mbar_wait(&tempty_bar[acc], ((lt >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait_u32(b_full + 8 * s, ph);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              const uint64_t ad = ((uint64_t)d_hi << 32) | (s_lo + KSTEP16 * k);
              const uint64_t wd = ((uint64_t)d_hi << 32) | (s_lo + A16 + KSTEP16_W * k);
              if constexpr (NCTA == 2) mma_ss_2cta(d_tmem, ad, wd, idesc, ((kb - kb0) | k) != 0);
              else mma_ss(d_tmem, ad, wd, idesc, ((kb - kb0) | k) != 0);
            }
            if constexpr (NCTA == 2) tc_commit_2cta(&empty_bar[s]);
            else asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b_empty + 8 * s) : "memory");
          }
          __syncwarp();
          s_lo += STAGE16;
          if (++s == STAGES) { s = 0; ph ^= 1; s_lo = a_lo0; }
        }
        if (elect_one()) {
          if constexpr (NCTA == 2) tc_commit_2cta(&tfull_bar[acc]); else tc_commit(&tfull_bar[acc]);
        }
        __syncwarp();
      }
    }
  } else {
    const int ew = warp - 2;                       // 0..7
    const int q = warp & 3;                        // TMEM lane quarter this warp may read
    const int half = ew >> 2;                      // which half of the tile's columns
    constexpr int HALF = BN / 2;
    const int et = threadIdx.x - 64;               // 0..255
    constexpr int mode = MODE;
    constexpr bool out16 = (mode == 0 || mode == 1 || mode == 3 || mode == 6 || mode == 7 || mode == 8);
    uint8_t* stg = stg_base + ew * 2 * 4096;
    const uint32_t sw = (uint32_t)(lane & 7);      // SWIZZLE_128B: 16 B piece j of row r sits at j ^ (r & 7)
    int lt = 0, sbuf = 0;
    uint32_t xph = 0;                              // phase bits of this warp's two residual barriers
    for (int item = first_tile; item < num_tiles; item += tile_step, ++lt) {
      const int tile = item % out_tiles;
      const bool first_split = item < out_tiles;     // the bias is added once
      const int tile_m = (tile / tiles_n) * NCTA + (int)rank, tile_n = tile % tiles_n;
      const int acc = lt & 1;
      float* sb = s_bias[acc];
      for (int c = et; c < BN; c += 256) {
        const int col = tile_n * BN + c;
        sb[c] = (ep.bias && col < N && first_split) ? __ldg(ep.bias + col) : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int row0 = tile_m * kBM + q * 32, row = row0 + lane;
      const bool row_ok = row < M;
      const int colh = tile_n * BN + half * HALF;   // first column of this warp's half
      const uint32_t tacc = tmem + acc * BN + half * HALF + ((uint32_t)(q * 32) << 16);
      if constexpr (out16) {
        // ---------------- fp16 outputs: 64-column chunks (128 B rows)
        // mode 7 (GEGLU, model/autoencoder.py:90-93): the W rows of a tile are [BN/2 value rows | BN/2 gate rows]
        // (interleaved once at load time), so value j and gate j of an output column sit BN/2 accumulator columns
        // apart in the same TMEM lane; the tile's output is BN/2 wide and each warp owns 64 of its columns
        constexpr int NCOLS = (mode == 7) ? 64 : HALF;
        const int nlim = (mode == 7) ? (N >> 1) : N;
#pragma unroll 1
        for (int c0 = 0; c0 < NCOLS; c0 += 64) {
          const int col0 = (mode == 7) ? tile_n * (BN / 2) + half * 64 : colh + c0;
          const uint32_t tsrc = (mode == 7) ? tmem + acc * BN + half * 64 + ((uint32_t)(q * 32) << 16) : tacc + c0;
          const int sboff = (mode == 7) ? half * 64 : half * HALF + c0;
          uint4 old[8];
          if (mode == 3) {                          // residual rows first: latency overlaps the accumulator wait
            const __half* orow = reinterpret_cast<const __half*>(ep.out) + (size_t)row * ep.ldo + col0;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              old[j] = (row_ok && col0 + 8 * j < N) ? *reinterpret_cast<const uint4*>(orow + 8 * j)
                                                    : make_uint4(0u, 0u, 0u, 0u);
          }
          if (mode == 8) {                          // GELU backward: the saved pre-activation rows, same overlap
            const __half* hrow = ep.gate + (size_t)row * ep.gate_stride + col0;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              old[j] = (row_ok && col0 + 8 * j < N) ? __ldg(reinterpret_cast<const uint4*>(hrow + 8 * j))
                                                    : make_uint4(0u, 0u, 0u, 0u);
          }
          if (c0 == 0) {
            mbar_wait(&tfull_bar[acc], (lt >> 1) & 1);
            tc_fence_after();
          }
          float v[64];
          {
            uint32_t rr[64];
            tmem_ld_x32(tsrc, *reinterpret_cast<uint32_t(*)[32]>(&rr[0]));
            tmem_ld_x32(tsrc + 32, *reinterpret_cast<uint32_t(*)[32]>(&rr[32]));
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 64; ++j) v[j] = __uint_as_float(rr[j]) + sb[sboff + j];
          }
          if constexpr (mode == 7) {
            // out = fp16(fp16(value) * fp16(gelu_erf(fp16(gate)))): the rounding points of the reference's autocast
            // Linear -> chunk -> F.gelu -> multiply (and of geglu_kernel, which this epilogue replaces)
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
              uint32_t gg[32];
              tmem_ld_x32(tsrc + BN / 2 + 32 * h2, gg);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float x = r16(__uint_as_float(gg[j]) + sb[BN / 2 + sboff + 32 * h2 + j]);
                const float ge = r16(0.5f * x * (1.0f + erff(x * 0.7071067811865476f)));
                v[32 * h2 + j] = r16(v[32 * h2 + j]) * ge;
              }
            }
          }
          if (mode == 1) {
            if (ep.gate) {
              // training forward: the fp16 pre-activation is kept for GELU' (rows are 128 B per thread and chunk)
              __half* hrow = const_cast<__half*>(ep.gate) + (size_t)row * ep.gate_stride + col0;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                uint4 pk;
                uint32_t* pw = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  const __half2 h2 = __floats2half2_rn(v[8 * j + 2 * t], v[8 * j + 2 * t + 1]);
                  pw[t] = *reinterpret_cast<const uint32_t*>(&h2);
                }
                if (row_ok && col0 + 8 * j < N) *reinterpret_cast<uint4*>(hrow + 8 * j) = pk;
              }
            }
#pragma unroll
            for (int j = 0; j < 64; ++j) v[j] = gelu_tanh_fast(r16(v[j]));
          } else if (mode == 8) {
            // out = fp16(fp16(acc) * gelu'(h0)): the dgrad of the MLP's second Linear with the activation's backward
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const __half2* o2 = reinterpret_cast<const __half2*>(&old[j]);
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                v[8 * j + 2 * t] = r16(v[8 * j + 2 * t]) * gelu_tanh_grad_fast(__low2float(o2[t]));
                v[8 * j + 2 * t + 1] = r16(v[8 * j + 2 * t + 1]) * gelu_tanh_grad_fast(__high2float(o2[t]));
              }
            }
          } else if (mode == 6) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const int hc = col0 + hh * 32;
              if (hc < ep.norm_cols) {
                float ss = 0.f;
#pragma unroll
                for (int j = 0; j < 32; ++j) { v[hh * 32 + j] = r16(v[hh * 32 + j]); ss += v[hh * 32 + j] * v[hh * 32 + j]; }
                const float inv = 5.656854249492381f / fmaxf(sqrtf(ss), 1e-12f);
                const int half_cols = ep.norm_cols >> 1;
                const float* gm = (hc < half_cols) ? ep.gamma_q + hc : ep.gamma_k + (hc - half_cols);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[hh * 32 + j] = v[hh * 32 + j] * inv * __ldg(gm + j);
              }
            }
          } else if (mode == 3) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const __half2* o2 = reinterpret_cast<const __half2*>(&old[j]);
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                v[8 * j + 2 * t] = r16(v[8 * j + 2 * t]) + __low2float(o2[t]);
                v[8 * j + 2 * t + 1] = r16(v[8 * j + 2 * t + 1]) + __high2float(o2[t]);
              }
            }
          }
          if (lane == 0) tma_store_wait_read<1>();   // the buffer written two chunks ago has been read out
          __syncwarp();
          uint8_t* buf = stg + sbuf * 4096 + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint4 pk;
            uint32_t* pw = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const __half2 h2 = __floats2half2_rn(v[8 * j + 2 * t], v[8 * j + 2 * t + 1]);
              pw[t] = *reinterpret_cast<const uint32_t*>(&h2);
            }
            *reinterpret_cast<uint4*>(buf + ((j ^ sw) << 4)) = pk;
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0 && col0 < nlim && row0 < M) {
            tma_store_2d(&mapO, stg + sbuf * 4096, col0, row0);
            tma_store_commit();
          }
          sbuf ^= 1;
        }
      } else {
        // ---------------- fp32 outputs: 32-column chunks (128 B rows)
        // Residual by TMA (mode 2, BN = 128: a warp's half tile is exactly two chunks = its two staging buffers).
        // The thread = row loads of the first version touched 32 different lines per instruction, half of every
        // 32 B sector unused (L1 is ~30 KB next to 193 KB of shared memory), and only the first chunk's latency
        // overlapped the main loop: the fp32 residual epilogue cost 8 us per launch over a plain fp16 store
        // (tools/gemm_depth_probe.py: out-proj 21.3 vs 13 us, fc2 39.9 vs 30).  Now lane 0 requests both
        // 32 x 32 fp32 boxes of the tile at tile start through the output's own tensor map; they land, swizzled
        // like the store wants them, while the main loop of the tile runs; the thread adds into its row in place.
        // BN = 256 (four chunks per warp): chunks 2 and 3 are requested as soon as the stores of chunks 0 and 1 have
        // been read out of their buffers.
        constexpr bool kTmaResid = (mode == 2);
        if constexpr (kTmaResid) {
          if (lane == 0) {
            tma_store_wait_read<0>();             // the previous tile's stores have left both buffers
#pragma unroll
            for (int b = 0; b < 2; ++b) {
              if (colh + 32 * b < N && row0 < M) {
                mbar_arrive_expect_tx(&xres_bar[ew][b], 4096);
                tma_load_2d(stg + b * 4096, &mapO, &xres_bar[ew][b], colh + 32 * b, row0);
              }
            }
          }
          __syncwarp();
          sbuf = 0;
        }
        float g[32];
        bool have_gate = false;
#pragma unroll 1
        for (int c0 = 0; c0 < HALF; c0 += 32) {
          const int col0 = colh + c0;
          float4 x[8];
          if (mode == 2) {
            if constexpr (!kTmaResid) {
              const float* orow = reinterpret_cast<const float*>(ep.out) + (size_t)row * ep.ldo + col0;
#pragma unroll
              for (int j = 0; j < 8; ++j)
                x[j] = (row_ok && col0 + 4 * j < N) ? *reinterpret_cast<const float4*>(orow + 4 * j)
                                                    : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            have_gate = ep.gate != nullptr;
            if (have_gate) {
              const __half* gp = ep.gate + (size_t)((row_ok ? row : 0) / ep.rows_per_batch) * ep.gate_stride + col0;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 gg = make_uint4(0u, 0u, 0u, 0u);
                if (col0 + 8 * j < N) gg = *reinterpret_cast<const uint4*>(gp + 8 * j);
                const __half2* g2 = reinterpret_cast<const __half2*>(&gg);
#pragma unroll
                for (int t = 0; t < 4; ++t) { g[8 * j + 2 * t] = __low2float(g2[t]); g[8 * j + 2 * t + 1] = __high2float(g2[t]); }
              }
            }
          }
          if (c0 == 0) {
            mbar_wait(&tfull_bar[acc], (lt >> 1) & 1);
            tc_fence_after();
          }
          float v[32];
          {
            uint32_t rr[32];
            tmem_ld_x32(tacc + c0, rr);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rr[j]) + sb[half * HALF + c0 + j];
          }
          if constexpr (kTmaResid) {
            if (col0 < N && row0 < M) {
              mbar_wait(&xres_bar[ew][sbuf], (xph >> sbuf) & 1u);   // own phase bits: ragged tiles skip boxes
              xph ^= 1u << sbuf;
              const uint8_t* xb = stg + sbuf * 4096 + lane * 128;
#pragma unroll
              for (int j = 0; j < 8; ++j) x[j] = *reinterpret_cast<const float4*>(xb + ((j ^ sw) << 4));
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) x[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
          if (mode == 2) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float* xf = reinterpret_cast<float*>(&x[j]);
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                float h = r16(v[4 * j + t]);
                if (have_gate) h = r16(h * g[4 * j + t]);
                v[4 * j + t] = xf[t] + h;
              }
            }
          }
          if constexpr (!kTmaResid) {
            if (lane == 0) tma_store_wait_read<1>();
            __syncwarp();
          }
          uint8_t* buf = stg + sbuf * 4096 + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(buf + ((j ^ sw) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0 && col0 < N && row0 < M) {
            if (mode == 4 && ksplit > 1) tma_reduce_add_2d(&mapO, stg + sbuf * 4096, col0, row0);
            else tma_store_2d(&mapO, stg + sbuf * 4096, col0, row0);
            tma_store_commit();
          }
          if constexpr (kTmaResid && HALF > 64) {
            if (lane == 0 && c0 + 64 < HALF && col0 + 64 < N && row0 < M) {
              tma_store_wait_read<0>();           // this chunk's store has left the buffer
              mbar_arrive_expect_tx(&xres_bar[ew][sbuf], 4096);
              tma_load_2d(stg + sbuf * 4096, &mapO, &xres_bar[ew][sbuf], col0 + 64, row0);
            }
          }
          sbuf ^= 1;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (NCTA == 2) mbar_arrive_leader(&tempty_bar[acc]); else mbar_arrive(&tempty_bar[acc]);
      }
    }
    if (lane == 0) tma_store_wait_all();
  }
  pdl_launch_dependents();          // this CTA only has its teardown left
  tc_fence_before();
  if constexpr (NCTA == 2) cluster_sync_all();      // nobody signals the peer's shared memory after this point
  else __syncthreads();
  if (warp == 2) {
    if constexpr (NCTA == 2) tmem_dealloc_2cta(tmem, 2 * BN); else tmem_dealloc(tmem, 2 * BN);
  }
}

template <int BN, int STAGES, int MODE, int NCTA = 1, bool TRANS = false, bool GATHER = false, bool TRANSB = false>
static int launch_gemm_ws(const CUtensorMap& mA, const CUtensorMap& mW, const CUtensorMap& mO, int M, int N, int K,
                          const GemmEpi& ep, cudaStream_t st, int ksplit = 1) {
  constexpr int SMEM = STAGES * (kBM * kBK * 2 + (BN / NCTA) * kBK * 2) + 8 * 2 * 4096 + 1024;
  static bool configured = false;
  static int num_sms = 0;
  if (!configured) {
    if (cudaFuncSetAttribute(gemm_ws_kernel<BN, STAGES, MODE, NCTA, TRANS, GATHER, TRANSB>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) !=
        cudaSuccess)
      return GVF_ERR_CUDA;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    configured = true;
  }
  const int tiles = ((N + BN - 1) / BN) * ((M + NCTA * kBM - 1) / (NCTA * kBM)) * ksplit;
  const int slots = num_sms / NCTA;
  const int grid = (tiles < slots ? tiles : slots) * NCTA;
  if constexpr (NCTA == 1) {
    return launch_pdl(gemm_ws_kernel<BN, STAGES, MODE, 1, TRANS, GATHER, TRANSB>, dim3(grid), dim3(320), SMEM, st, mA, mW, mO, M, N, K, ep, ksplit) == cudaSuccess
               ? GVF_OK : GVF_ERR_CUDA;
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(320);
    cfg.dynamicSmemBytes = SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, gemm_ws_kernel<BN, STAGES, MODE, 2>, mA, mW, mO, M, N, K, ep, ksplit) == cudaSuccess ? GVF_OK
                                                                                                                    : GVF_ERR_CUDA;
  }
}

// ---------------------------------------------------------------------------------------
// Residual GEMM fused with the LayerNorm (+ adaLN modulate / affine) that consumes its result:
//     X[M,512] += gate * fp16(A W^T + b)          (fp32 residual stream, as mode 2 above)
//     Y[M,512]  = fp16( LN(X) * (1 + scale) + shift )   or   fp16( LN(X) * w + b )
// reference model/dit.py:246-277: every attention out-projection / MLP fc2 is followed by the norm of the
// next sub-block.  As two kernels that was 21 us + 10 us of mostly launch, ramp and a second pass over X.
// One CTA owns complete rows (a 128 x 512 tile: the whole TMEM as one fp32 accumulator, two N = 256 MMAs
// per k-step), so the row statistics never leave the CTA:
//   pass 1  (thread = row, each of the two warps of a row quarter owns 256 columns) accumulator + bias ->
//           fp16 rounding -> gate -> + residual; the new X goes to global memory through the TMA staging
//           tiles and back into TMEM; shifted sums give (mean, M2) of the warp's half row, Chan's formula
//           merges the two halves through shared memory
//   pass 2  X re-read from TMEM -> normalise -> modulate -> fp16 -> TMA store of Y.
// 96 CTAs for M = 12288: not a full wave, but these are latency-bound launches and one of them disappears.
struct LnEpi {
  const float* bias;                // [512] or null
  const __half* gate;               // [batches, gate_stride] or null
  int gate_stride, rows_per_batch;
  float* x;                         // [M, ldx] fp32 in / out
  int ldx;
  const float* ln_w;                // affine LayerNorm ([512], [512]) or null
  const float* ln_b;
  const __half* shift;              // adaLN modulation ([batches, mod_stride]) or null
  const __half* scale;
  int mod_stride;
  float eps;
};

constexpr int kLnBN = 512, kLnStages = 2;

__global__ void __launch_bounds__(320, 1)
gemm_ln_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW,
               const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapY, int M, int K,
               LnEpi ep) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[kLnStages], empty_bar[kLnStages], tfull_bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int A_BYTES = kBM * kBK * 2, W_BYTES = kLnBN * kBK * 2, STAGE_BYTES = A_BYTES + W_BYTES;
  uint8_t* stg_base = smem + kLnStages * STAGE_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kblocks = (K + kBK - 1) / kBK;
  const int tile_m = blockIdx.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kLnStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tfull_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapW);
    tma_prefetch_desc(&mapX);
    tma_prefetch_desc(&mapY);
  }
  if (warp == 2) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < kblocks; ++kb) {
        const int s = kb % kLnStages;
        mbar_wait(&empty_bar[s], ((kb / kLnStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
        uint8_t* st = smem + s * STAGE_BYTES;
        tma_load_2d(st, &mapA, &full_bar[s], kb * kBK, tile_m * kBM);
        tma_load_2d(st + A_BYTES, &mapW, &full_bar[s], kb * kBK, 0);              // W rows 0..255
        tma_load_2d(st + A_BYTES + W_BYTES / 2, &mapW, &full_bar[s], kb * kBK, 256);   // W rows 256..511
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(kBM, 256, 0, 0);
      for (int kb = 0; kb < kblocks; ++kb) {
        const int s = kb % kLnStages;
        mbar_wait(&full_bar[s], (kb / kLnStages) & 1);
        tc_fence_after();
        const uint32_t a0 = smem_u32(smem + s * STAGE_BYTES), w0 = a0 + A_BYTES;
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) {
          const uint64_t ad = make_smem_desc(a0 + k * 32, 16, 1024, SWZ_128B);
          mma_ss(tmem, ad, make_smem_desc(w0 + k * 32, 16, 1024, SWZ_128B), idesc, (kb | k) != 0);
          mma_ss(tmem + 256, ad, make_smem_desc(w0 + W_BYTES / 2 + k * 32, 16, 1024, SWZ_128B), idesc, (kb | k) != 0);
        }
        tc_commit(&empty_bar[s]);
      }
      tc_commit(&tfull_bar);
    }
  } else {
    const int ew = warp - 2, q = warp & 3, half = ew >> 2;
    uint8_t* stg = stg_base + ew * 2 * 4096;
    const uint32_t sw = (uint32_t)(lane & 7);
    const int row0 = tile_m * kBM + q * 32, row = row0 + lane;
    const bool row_ok = row < M;
    const int colh = half * 256;
    const uint32_t tacc = tmem + colh + ((uint32_t)(q * 32) << 16);
    const int rsafe = row_ok ? row : 0;
    const float* xrow = ep.x + (size_t)rsafe * ep.ldx + colh;
    const __half* grow = ep.gate ? ep.gate + (size_t)(rsafe / ep.rows_per_batch) * ep.gate_stride + colh : nullptr;
    int sbuf = 0;
    float x0 = 0.f, s1 = 0.f, s2 = 0.f;
    // residual rows of the first chunk travel while the main loop runs
    float4 xr[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) xr[j] = row_ok ? *reinterpret_cast<const float4*>(xrow + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
    mbar_wait(&tfull_bar, 0);
    tc_fence_after();
    // ---------------- pass 1: new X (global + TMEM) and the half-row statistics
#pragma unroll 1
    for (int c0 = 0; c0 < 256; c0 += 32) {
      float v[32];
      {
        uint32_t rr[32];
        tmem_ld_x32(tacc + c0, rr);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rr[j]);
      }
      if (ep.bias) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(ep.bias + colh + c0) + j);
          v[4 * j] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w;
        }
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = r16(v[j]);
      if (grow) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 gg = *reinterpret_cast<const uint4*>(grow + c0 + 8 * j);
          const __half2* g2 = reinterpret_cast<const __half2*>(&gg);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            v[8 * j + 2 * t] = r16(v[8 * j + 2 * t] * __low2float(g2[t]));
            v[8 * j + 2 * t + 1] = r16(v[8 * j + 2 * t + 1] * __high2float(g2[t]));
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[4 * j] += xr[j].x; v[4 * j + 1] += xr[j].y; v[4 * j + 2] += xr[j].z; v[4 * j + 3] += xr[j].w;
      }
      if (c0 + 32 < 256) {                          // next chunk's residual
#pragma unroll
        for (int j = 0; j < 8; ++j)
          xr[j] = row_ok ? *reinterpret_cast<const float4*>(xrow + c0 + 32 + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (c0 == 0) x0 = v[0];
#pragma unroll
      for (int j = 0; j < 32; ++j) { const float d = v[j] - x0; s1 += d; s2 = fmaf(d, d, s2); }
      {
        uint32_t rr[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) rr[j] = __float_as_uint(v[j]);
        tmem_st_x32(tacc + c0, rr);
      }
      if (lane == 0) tma_store_wait_read<1>();
      __syncwarp();
      uint8_t* buf = stg + sbuf * 4096 + lane * 128;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(buf + ((j ^ sw) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0 && row0 < M) {
        tma_store_2d(&mapX, stg + sbuf * 4096, colh + c0, row0);
        tma_store_commit();
      }
      sbuf ^= 1;
    }
    tmem_st_wait();
    // ---------------- merge the two half rows (Chan et al.): n_a = n_b = 256
    const float mean_h = x0 + s1 * (1.0f / 256.0f);
    const float m2_h = fmaxf(s2 - s1 * s1 * (1.0f / 256.0f), 0.f);
    float2* s_stat = reinterpret_cast<float2*>(smem);   // the pipeline stages are idle once tfull_bar has fired
    s_stat[half * kBM + q * 32 + lane] = make_float2(mean_h, m2_h);
    asm volatile("bar.sync 2, 256;" ::: "memory");
    const float2 other = s_stat[(half ^ 1) * kBM + q * 32 + lane];
    const float dm = mean_h - other.x;
    const float mean = 0.5f * (mean_h + other.x);
    const float var = (m2_h + other.y + dm * dm * 128.0f) * (1.0f / 512.0f);
    const float rstd = rsqrtf(var + ep.eps);
    const __half* shrow = ep.shift ? ep.shift + (size_t)(rsafe / ep.rows_per_batch) * ep.mod_stride + colh : nullptr;
    const __half* scrow = ep.scale ? ep.scale + (size_t)(rsafe / ep.rows_per_batch) * ep.mod_stride + colh : nullptr;
    // ---------------- pass 2: normalise + modulate -> fp16
#pragma unroll 1
    for (int c0 = 0; c0 < 256; c0 += 64) {
      float v[64];
      {
        uint32_t rr[64];
        tmem_ld_x32(tacc + c0, *reinterpret_cast<uint32_t(*)[32]>(&rr[0]));
        tmem_ld_x32(tacc + c0 + 32, *reinterpret_cast<uint32_t(*)[32]>(&rr[32]));
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 64; ++j) v[j] = (__uint_as_float(rr[j]) - mean) * rstd;
      }
      if (ep.ln_w) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 ww = __ldg(reinterpret_cast<const float4*>(ep.ln_w + colh + c0) + j);
          const float4 bb = __ldg(reinterpret_cast<const float4*>(ep.ln_b + colh + c0) + j);
          v[4 * j] = v[4 * j] * ww.x + bb.x; v[4 * j + 1] = v[4 * j + 1] * ww.y + bb.y;
          v[4 * j + 2] = v[4 * j + 2] * ww.z + bb.z; v[4 * j + 3] = v[4 * j + 3] * ww.w + bb.w;
        }
      }
      if (scrow) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint4 sc = *reinterpret_cast<const uint4*>(scrow + c0 + 8 * j);
          const uint4 sh = *reinterpret_cast<const uint4*>(shrow + c0 + 8 * j);
          const __half2* sc2 = reinterpret_cast<const __half2*>(&sc);
          const __half2* sh2 = reinterpret_cast<const __half2*>(&sh);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            v[8 * j + 2 * t] = v[8 * j + 2 * t] * (1.0f + __low2float(sc2[t])) + __low2float(sh2[t]);
            v[8 * j + 2 * t + 1] = v[8 * j + 2 * t + 1] * (1.0f + __high2float(sc2[t])) + __high2float(sh2[t]);
          }
        }
      }
      if (lane == 0) tma_store_wait_read<1>();
      __syncwarp();
      uint8_t* buf = stg + sbuf * 4096 + lane * 128;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint4 pk;
        uint32_t* pw = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const __half2 h2 = __floats2half2_rn(v[8 * j + 2 * t], v[8 * j + 2 * t + 1]);
          pw[t] = *reinterpret_cast<const uint32_t*>(&h2);
        }
        *reinterpret_cast<uint4*>(buf + ((j ^ sw) << 4)) = pk;
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0 && row0 < M) {
        tma_store_2d(&mapY, stg + sbuf * 4096, colh + c0, row0);
        tma_store_commit();
      }
      sbuf ^= 1;
    }
    if (lane == 0) tma_store_wait_all();
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

template <int BN, int STAGES>
static int launch_gemm_persistent(const CUtensorMap& mA, const CUtensorMap& mW, int M, int N, int K,
                                  const GemmEpi& ep, cudaStream_t st) {
  constexpr int SMEM = STAGES * (kBM * kBK * 2 + BN * kBK * 2) + 4 * 32 * kStgLd * 4 + 1024;
  static bool configured = false;
  static int num_sms = 0;
  if (!configured) {
    if (cudaFuncSetAttribute(gemm_persistent_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             SMEM) != cudaSuccess)
      return GVF_ERR_CUDA;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    configured = true;
  }
  const int tiles = ((N + BN - 1) / BN) * ((M + kBM - 1) / kBM);
  const int grid = tiles < num_sms ? tiles : num_sms;
  gemm_persistent_kernel<BN, STAGES><<<grid, 192, SMEM, st>>>(mA, mW, M, N, K, ep);
  return cudaGetLastError() == cudaSuccess ? GVF_OK : GVF_ERR_CUDA;
}

template <int BN, int STAGES, int EC = kEpiCols, int MINB = 2>
static int launch_gemm(const CUtensorMap& mA, const CUtensorMap& mW, int M, int N, int K,
                       const GemmEpi& ep, cudaStream_t st) {
  constexpr int SMEM = STAGES * (kBM * kBK * 2 + BN * kBK * 2) + 1024;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(gemm_kernel<BN, STAGES, EC, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             SMEM) != cudaSuccess)
      return GVF_ERR_CUDA;
    configured = true;
  }
  dim3 grid((N + BN - 1) / BN, (M + kBM - 1) / kBM);
  gemm_kernel<BN, STAGES, EC, MINB><<<grid, 192, SMEM, st>>>(mA, mW, M, N, K, ep);
  return cudaGetLastError() == cudaSuccess ? GVF_OK : GVF_ERR_CUDA;
}

}  // namespace gvf

using namespace gvf;

static int g_gemm_variant = -1;
static int g_gemm_ksplit = 0;      // 0 automatic, -1 never, n > 0 forced (fp32-store epilogue only)

// Split-K factor of the fp32-store GEMMs (weight gradients).  Work items = tiles x ksplit run in waves of one item per
// SM, so what matters is the wave count: 18 tiles x 8 = 144 items are one wave, x 10 = 180 are two with the second 78 %
// empty (768 x 768 x 393216: 407 us at 8, 623 us at 10, 409 us at 16 -- tools/gemm_tn_ksplit_sweep.py,
// profiles/r02_gemm_tn_ksplit.txt).  Cost model fitted to that sweep, in us: waves x (k-blocks per item x 0.5 + 3.7
// epilogue incl. the TMA reduce-add) + 3 for the zero fill when split; an un-split kernel on at most half the SMs sees
// 0.42 us per k-block (the operand stream L2 -> SM is the limit, fewer CTAs get more of it).  It picks the measured optimum on
// every shape of the two training steps.
static int sm_count() {             // one device kind per process: queried once (the attribute call costs microseconds)
  static int sms = 0;
  if (sms == 0) {
    int dev = 0, n = 148;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    sms = n;
  }
  return sms;
}
static int pick_ksplit(long long tiles, int kblocks, int sms) {
  if (kblocks < 8) return 1;
  const int kmax = kblocks / 4 < 32 ? kblocks / 4 : 32;
  double cost[33];
  double best = 1e30;
  for (int ks = 1; ks <= 32; ++ks) {
    cost[ks] = 1e30;
    if (ks > (kmax > 1 ? kmax : 1)) continue;
    const int per = (kblocks + ks - 1) / ks;
    if (ks > 1 && (kblocks + per - 1) / per != ks) continue;        // would leave an empty split
    const long long waves = (tiles * ks + sms - 1) / sms;
    const double a = (ks == 1 && tiles * 2 <= sms) ? 0.42 : 0.5;
    cost[ks] = waves * (per * a + 3.7) + (ks > 1 ? 3.0 : 0.0);
    if (cost[ks] < best) best = cost[ks];
  }
  // within 2.5 % of the optimum prefer the finest split: shorter fp32 accumulation chains in the tensor core (which
  // truncates), e.g. 24 x 16384 instead of 8 x 49152 products for the decoder's 393216-row weight gradients
  for (int ks = 32; ks >= 1; --ks)
    if (cost[ks] <= best * 1.025) return ks;
  return 1;
}

static int gemm_impl(const void* A, int lda, const void* W, int ldw, int M, int N, int K, int epilogue,
                     const float* bias, void* out, int ldo, const void* gate, int gate_stride,
                     int rows_per_batch, const float* gamma_q, const float* gamma_k, int norm_cols,
                     void* stream) {
  if (!A || !W || !out || M <= 0 || N <= 0 || K <= 0) return GVF_ERR_INVALID;
  if ((N % 8) || (K % 8) || (lda % 8) || (ldw % 8) || epilogue < 0 || epilogue > 8) return GVF_ERR_INVALID;
  if (epilogue == 8 && !gate) return GVF_ERR_INVALID;
  if (epilogue == 6 && (!gamma_q || !gamma_k || norm_cols <= 0 || (norm_cols % 64) || norm_cols > N))
    return GVF_ERR_INVALID;
  if ((epilogue == 0 || epilogue == 1 || epilogue == 3 || epilogue == 6 || epilogue == 7 || epilogue == 8) && (ldo % 2))
    return GVF_ERR_INVALID;
  if (epilogue == 7 && (N % 256)) return GVF_ERR_UNSUPPORTED;      // GEGLU: whole 256-row weight tiles [128 value | 128 gate]
  if ((epilogue == 2 || epilogue == 4) && (ldo % 2)) return GVF_ERR_INVALID;
  if (gate && (gate_stride % 2)) return GVF_ERR_INVALID;
  if (((uintptr_t)A | (uintptr_t)W | (uintptr_t)out) & 15) return GVF_ERR_INVALID;
  if (epilogue == 2 && gate && rows_per_batch <= 0) return GVF_ERR_INVALID;
  // variant: 0 = one tile per CTA, 2 CTAs / SM, 3 stages; 1 = persistent 128x128; 2 = persistent 128x256;
  //          3 = one tile per CTA, 3 CTAs / SM, 2 stages, 32-column epilogue chunks;
  //          4 / 5 = generation 2 (8 epilogue warps, TMA stores) 128x128 / 128x256; 6 / 7 = generation 2 on CTA pairs,
  //          256x128 / 256x256 pair tiles
  int variant = g_gemm_variant;
  if (variant < 0) {
    // measured on B200 (tools/gemm_bench.py, profiles/r01_gemm_variants.txt): the generation-2 kernels win on
    // every shape of the path; 128 x 256 tiles whenever N allows them and there is more than one column of
    // tiles per row block to amortise the wider epilogue (N >= 768), or the epilogue is a plain fp16 store
    variant = (N % 256 == 0 && (N >= 768 || epilogue == 0 || epilogue == 1)) ? 5 : 4;
    // long-K residual shapes (fc2: N = 512, K = 2048) stay on 128 x 128 tiles: since the MMA issue loop became
    // warp-uniform (round 2) they run 30.1 us against 31.5 (128 x 256) / 31.4 (256 x 256 pairs) -- the wide tiles'
    // 96 / 192 work items quantise badly on 148 SMs, which used to be hidden behind the slow issue loop
    // CTA pairs (tcgen05.mma.cta_group::2, 256 x 256 tiles, each CTA stages half of W): -6..-11 % on the long-K
    // motion-VAE shapes (ff1 101.9 -> 91.0 us, cuBLAS 92.1), nothing on the K = 512 DiT shapes, which are bound by
    // ramp-up and epilogue latency rather than by L2 -> SM operand traffic
    if (variant == 5 && K >= 768 && (long long)((M + 2 * kBM - 1) / (2 * kBM)) * (N / 256) >= 74) variant = 7;
  }
  if (epilogue == 7 && variant != 5 && variant != 7) variant = 5;   // the interleaving is defined on 256-column tiles
  // generation-2 kernels need a TMA-storable output (16 B aligned rows) and do not do the compact mode 5
  if (variant >= 4 && variant <= 9 &&
      (epilogue == 5 || (ldo * ((epilogue == 2 || epilogue == 4) ? 4 : 2)) % 16 != 0 ||
       (gate && ((gate_stride % 8) || ((uintptr_t)gate & 15)))))
    variant = 0;
  if (epilogue == 7 && variant != 5 && variant != 7) return GVF_ERR_UNSUPPORTED;
  // the pre-activation side output of mode 1 and the GELU' epilogue exist in the generation-2 kernels only
  if ((epilogue == 8 || (epilogue == 1 && gate)) && !(variant >= 4 && variant <= 9)) return GVF_ERR_UNSUPPORTED;
  // 8 / 9: pipeline-depth experiments (128x128 with 5 stages, 256x128 pair tiles with 6 stages)
  const int BN = (variant == 2 || variant == 5 || variant == 7) ? 256 : 128;
  const int wbox = (variant == 6 || variant == 7 || variant == 9) ? BN / 2 : BN;      // W rows one CTA stages per k-block
  CUtensorMap mA, mW;
  const uint64_t dA[2] = {(uint64_t)K, (uint64_t)M}, sA[2] = {1, (uint64_t)lda};
  const uint64_t dW[2] = {(uint64_t)K, (uint64_t)N}, sW[2] = {1, (uint64_t)ldw};
  const uint32_t bA[2] = {kBK, kBM}, bW[2] = {kBK, (uint32_t)wbox};
  if (!make_tmap_f16(&mA, A, 2, dA, sA, bA, CU_TENSOR_MAP_SWIZZLE_128B)) return GVF_ERR_CUDA;
  if (!make_tmap_f16(&mW, W, 2, dW, sW, bW, CU_TENSOR_MAP_SWIZZLE_128B)) return GVF_ERR_CUDA;
  GemmEpi ep;
  ep.mode = epilogue; ep.bias = bias; ep.out = out; ep.gate = (const __half*)gate;
  ep.gate_stride = gate_stride; ep.rows_per_batch = rows_per_batch > 0 ? rows_per_batch : 1;
  ep.ldo = ldo;
  ep.gamma_q = gamma_q; ep.gamma_k = gamma_k; ep.norm_cols = norm_cols;
  if (variant >= 4 && variant <= 9) {
    const bool out16 = (epilogue == 0 || epilogue == 1 || epilogue == 3 || epilogue == 6 || epilogue == 7 || epilogue == 8);
    CUtensorMap mO;
    if (!make_tmap_2d(&mO, out, out16 ? 2 : 4, (uint64_t)(epilogue == 7 ? N / 2 : N), (uint64_t)M, (uint64_t)ldo,
                      out16 ? 64 : 32, 32))
      return GVF_ERR_CUDA;
    cudaStream_t cs = (cudaStream_t)stream;
#define GVF_WS(MODE)                                                                      \
    case MODE:                                                                            \
      return variant == 9   ? launch_gemm_ws<128, 6, MODE, 2>(mA, mW, mO, M, N, K, ep, cs) \
             : variant == 8 ? launch_gemm_ws<128, 5, MODE, 1>(mA, mW, mO, M, N, K, ep, cs)   \
             : variant == 7 ? launch_gemm_ws<256, 4, MODE, 2>(mA, mW, mO, M, N, K, ep, cs) \
             : variant == 6 ? launch_gemm_ws<128, 5, MODE, 2>(mA, mW, mO, M, N, K, ep, cs) \
             : variant == 5 ? launch_gemm_ws<256, 3, MODE>(mA, mW, mO, M, N, K, ep, cs)    \
                            : launch_gemm_ws<128, 4, MODE>(mA, mW, mO, M, N, K, ep, cs);
    if (epilogue == 4) {
      // fp32 outputs with few tiles and a long reduction (the wgrad GEMMs of the training step: 768 x 768 outputs
      // over K = 12288 are 18 tiles of 192 k-blocks): split K so that about one wave of work items exists; the
      // partial tiles are summed by TMA reduce-add stores into the zeroed output
      const int NC = (variant == 6 || variant == 7 || variant == 9) ? 2 : 1;
      const long long tiles = (long long)((N + BN - 1) / BN) * ((M + NC * kBM - 1) / (NC * kBM));
      const int kblocks = (K + kBK - 1) / kBK;
      const int sms = sm_count();
      int ksplit = 1;
      if (g_gemm_ksplit > 0) ksplit = g_gemm_ksplit;
      else if (g_gemm_ksplit == 0 && tiles * 2 <= sms / NC && kblocks >= 32) ksplit = pick_ksplit(tiles, kblocks, sms / NC);
      if (ksplit > 1) {
        const int per = (kblocks + ksplit - 1) / ksplit;
        ksplit = (kblocks + per - 1) / per;           // no empty split
      }
      if (ksplit > 1 &&
          cudaMemset2DAsync(out, (size_t)ldo * 4, 0, (size_t)N * 4, (size_t)M, cs) != cudaSuccess)
        return GVF_ERR_CUDA;
      return variant == 9   ? launch_gemm_ws<128, 6, 4, 2>(mA, mW, mO, M, N, K, ep, cs, ksplit)
             : variant == 8 ? launch_gemm_ws<128, 5, 4, 1>(mA, mW, mO, M, N, K, ep, cs, ksplit)
             : variant == 7 ? launch_gemm_ws<256, 4, 4, 2>(mA, mW, mO, M, N, K, ep, cs, ksplit)
             : variant == 6 ? launch_gemm_ws<128, 5, 4, 2>(mA, mW, mO, M, N, K, ep, cs, ksplit)
             : variant == 5 ? launch_gemm_ws<256, 3, 4>(mA, mW, mO, M, N, K, ep, cs, ksplit)
                            : launch_gemm_ws<128, 4, 4>(mA, mW, mO, M, N, K, ep, cs, ksplit);
    }
    switch (epilogue) {
      GVF_WS(0) GVF_WS(1) GVF_WS(2) GVF_WS(3) GVF_WS(6)
      case 7:
        return variant == 7 ? launch_gemm_ws<256, 4, 7, 2>(mA, mW, mO, M, N, K, ep, cs)
                            : launch_gemm_ws<256, 3, 7>(mA, mW, mO, M, N, K, ep, cs);
      case 8:
        return variant == 7   ? launch_gemm_ws<256, 4, 8, 2>(mA, mW, mO, M, N, K, ep, cs)
               : variant == 5 ? launch_gemm_ws<256, 3, 8>(mA, mW, mO, M, N, K, ep, cs)
               : variant == 4 ? launch_gemm_ws<128, 4, 8>(mA, mW, mO, M, N, K, ep, cs)
                              : GVF_ERR_UNSUPPORTED;
      default: return GVF_ERR_INVALID;
    }
#undef GVF_WS
  }
  if (variant == 2) return launch_gemm_persistent<256, 3>(mA, mW, M, N, K, ep, (cudaStream_t)stream);
  if (variant == 1) return launch_gemm_persistent<128, 4>(mA, mW, M, N, K, ep, (cudaStream_t)stream);
  if (variant == 3) return launch_gemm<128, 2, 32, 3>(mA, mW, M, N, K, ep, (cudaStream_t)stream);   // 3 CTAs / SM
  return launch_gemm<128, 3>(mA, mW, M, N, K, ep, (cudaStream_t)stream);
}

// tuning hook for the benchmarks: -1 automatic, 0 v1 (one tile per CTA), 1 persistent 128x128, 2 persistent 128x256
extern "C" GVF_API void gvf_gemm_set_variant(int v) { g_gemm_variant = v; }
extern "C" GVF_API void gvf_set_pdl(int on) { gvf::g_pdl_enabled = on ? 1 : 0; }
extern "C" GVF_API void gvf_gemm_set_ksplit(int k) { g_gemm_ksplit = k; }

extern "C" GVF_API int gvf_gemm_f16(const void* A, int lda, const void* W, int ldw, int M, int N,
                                    int K, int epilogue, const float* bias, void* out, int ldo,
                                    const void* gate, int gate_stride, int rows_per_batch,
                                    void* stream) {
  if (epilogue == 6 || epilogue == 7) return GVF_ERR_INVALID;
  return gemm_impl(A, lda, W, ldw, M, N, K, epilogue, bias, out, ldo, gate, gate_stride, rows_per_batch,
                   nullptr, nullptr, 0, stream);
}

// FeedForward's first Linear with GEGLU in the epilogue (reference model/autoencoder.py:90-107):
// out[M, N/2] = fp16(value * gelu_erf(gate)).  W / bias rows must be interleaved per 256-row tile:
// rows [256 t, 256 t + 128) = value rows [128 t, 128 t + 128) of net.0, rows [256 t + 128, 256 t + 256) = gate rows
// [N/2 + 128 t, ...) (gvfdiffusion_b200/ops.py: geglu_interleave).  N % 256 == 0.
extern "C" GVF_API int gvf_gemm_geglu_f16(const void* A, int lda, const void* W, int ldw, int M, int N, int K,
                                          const float* bias, void* out, int ldo, void* stream) {
  return gemm_impl(A, lda, W, ldw, M, N, K, 7, bias, out, ldo, nullptr, 0, 0, nullptr, nullptr, 0, stream);
}

// QKV projection with MultiHeadRMSNorm fused into the epilogue (reference
// model/attention/modules.py:113-125): out fp16 [M,N]; columns [0, norm_cols) are split into
// 32-wide heads, first half normalised with gamma_q [norm_cols/64, 32], second half with gamma_k.
extern "C" GVF_API int gvf_gemm_qkv_rmsnorm_f16(const void* A, int lda, const void* W, int ldw, int M,
                                                int N, int K, const float* bias, void* out, int ldo,
                                                const float* gamma_q, const float* gamma_k,
                                                int norm_cols, void* stream) {
  return gemm_impl(A, lda, W, ldw, M, N, K, 6, bias, out, ldo, nullptr, 0, 0, gamma_q, gamma_k, norm_cols,
                   stream);
}

// Residual Linear + the LayerNorm of the next sub-block in one kernel (see gemm_ln_kernel): N is fixed to
// 512 (one CTA owns whole rows).  x fp32 [M, ldx] in/out; y fp16 [M, ldy] = LN(x_new) * (1 + scale) + shift
// (shift/scale fp16 [batches, mod_stride], batch = row / rows_per_batch) or * ln_w + ln_b (fp32 [512]) or plain.
extern "C" GVF_API int gvf_gemm_resid_ln_f16(const void* A, int lda, const void* W, int ldw, int M, int N, int K,
                                             const float* bias, float* x, int ldx, const void* gate,
                                             int gate_stride, int rows_per_batch, const float* ln_w,
                                             const float* ln_b, const void* shift, const void* scale,
                                             int mod_stride, float eps, void* y, int ldy, void* stream) {
  if (!A || !W || !x || !y || M <= 0 || K <= 0) return GVF_ERR_INVALID;
  if (N != kLnBN) return GVF_ERR_UNSUPPORTED;
  if ((K % 8) || (lda % 8) || (ldw % 8) || (ldx % 4) || (ldy % 8)) return GVF_ERR_INVALID;
  if ((ln_w == nullptr) != (ln_b == nullptr) || (shift == nullptr) != (scale == nullptr)) return GVF_ERR_INVALID;
  if (((uintptr_t)A | (uintptr_t)W | (uintptr_t)x | (uintptr_t)y | (uintptr_t)bias | (uintptr_t)ln_w |
       (uintptr_t)ln_b | (uintptr_t)gate | (uintptr_t)shift | (uintptr_t)scale) & 15)
    return GVF_ERR_INVALID;
  if ((gate && (gate_stride % 8)) || (shift && (mod_stride % 8))) return GVF_ERR_INVALID;
  if ((gate || shift) && rows_per_batch <= 0) return GVF_ERR_INVALID;
  CUtensorMap mA, mW, mX, mY;
  const uint64_t dA[2] = {(uint64_t)K, (uint64_t)M}, sA[2] = {1, (uint64_t)lda};
  const uint64_t dW[2] = {(uint64_t)K, (uint64_t)N}, sW[2] = {1, (uint64_t)ldw};
  const uint32_t bA[2] = {kBK, kBM}, bW[2] = {kBK, 256};
  if (!make_tmap_f16(&mA, A, 2, dA, sA, bA, CU_TENSOR_MAP_SWIZZLE_128B)) return GVF_ERR_CUDA;
  if (!make_tmap_f16(&mW, W, 2, dW, sW, bW, CU_TENSOR_MAP_SWIZZLE_128B)) return GVF_ERR_CUDA;
  if (!make_tmap_2d(&mX, x, 4, (uint64_t)N, (uint64_t)M, (uint64_t)ldx, 32, 32)) return GVF_ERR_CUDA;
  if (!make_tmap_2d(&mY, y, 2, (uint64_t)N, (uint64_t)M, (uint64_t)ldy, 64, 32)) return GVF_ERR_CUDA;
  LnEpi ep;
  ep.bias = bias; ep.gate = (const __half*)gate; ep.gate_stride = gate_stride;
  ep.rows_per_batch = rows_per_batch > 0 ? rows_per_batch : M;
  ep.x = x; ep.ldx = ldx; ep.ln_w = ln_w; ep.ln_b = ln_b;
  ep.shift = (const __half*)shift; ep.scale = (const __half*)scale; ep.mod_stride = mod_stride; ep.eps = eps;
  constexpr int SMEM = kLnStages * (kBM * kBK * 2 + kLnBN * kBK * 2) + 8 * 2 * 4096 + 1024;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(gemm_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess)
      return GVF_ERR_CUDA;
    configured = true;
  }
  return launch_pdl(gemm_ln_kernel, dim3((M + kBM - 1) / kBM), dim3(320), SMEM, (cudaStream_t)stream, mA, mW, mX, mY,
                    M, K, ep) == cudaSuccess ? GVF_OK : GVF_ERR_CUDA;
}

// Weight gradient without transposed copies: out[M, N] fp32 (+ nothing) = A[R, M]^T W[R, N], A / W fp16 row-major
// activations (dY and X of a Linear: dW[out, in] = dY^T X).  Both operands enter the tensor core MN-major; split-K over R
// like the fp32-store epilogue of gvf_gemm_f16.  M, N multiples of 8 (16 B rows), any R.
extern "C" GVF_API int gvf_gemm_tn_f16(const void* A, int lda, const void* W, int ldw, int M, int N, int R, float* out,
                                       int ldo, void* stream) {
  if (!A || !W || !out || M <= 0 || N <= 0 || R <= 0) return GVF_ERR_INVALID;
  if ((M % 8) || (N % 8) || (lda % 8) || (ldw % 8) || (ldo % 4) || lda < M || ldw < N || ldo < N) return GVF_ERR_INVALID;
  if (((uintptr_t)A | (uintptr_t)W | (uintptr_t)out) & 15) return GVF_ERR_INVALID;
  CUtensorMap mA, mW, mO;
  const uint64_t dA[2] = {(uint64_t)M, (uint64_t)R}, sA[2] = {1, (uint64_t)lda};
  const uint64_t dW[2] = {(uint64_t)N, (uint64_t)R}, sW[2] = {1, (uint64_t)ldw};
  const uint32_t box[2] = {64, kBK};
  if (!make_tmap_f16(&mA, A, 2, dA, sA, box, CU_TENSOR_MAP_SWIZZLE_128B)) return GVF_ERR_CUDA;
  if (!make_tmap_f16(&mW, W, 2, dW, sW, box, CU_TENSOR_MAP_SWIZZLE_128B)) return GVF_ERR_CUDA;
  if (!make_tmap_2d(&mO, out, 4, (uint64_t)N, (uint64_t)M, (uint64_t)ldo, 32, 32)) return GVF_ERR_CUDA;
  GemmEpi ep;
  ep.mode = 4; ep.bias = nullptr; ep.out = out; ep.gate = nullptr; ep.gate_stride = 0; ep.rows_per_batch = 1; ep.ldo = ldo;
  ep.gamma_q = nullptr; ep.gamma_k = nullptr; ep.norm_cols = 0;
  cudaStream_t cs = (cudaStream_t)stream;
  const bool wide = (N % 256) == 0;
  const int BN = wide ? 256 : 128;
  const long long tiles = (long long)((N + BN - 1) / BN) * ((M + kBM - 1) / kBM);
  const int kblocks = (R + kBK - 1) / kBK;
  const int sms = sm_count();
  int ksplit = 1;
  if (g_gemm_ksplit > 0) ksplit = g_gemm_ksplit;
  else if (g_gemm_ksplit == 0) ksplit = pick_ksplit(tiles, kblocks, sms);
  if (ksplit > 1) {
    const int per = (kblocks + ksplit - 1) / ksplit;
    ksplit = (kblocks + per - 1) / per;
  }
  if (ksplit > 1 && cudaMemset2DAsync(out, (size_t)ldo * 4, 0, (size_t)N * 4, (size_t)M, cs) != cudaSuccess) return GVF_ERR_CUDA;
  return wide ? launch_gemm_ws<256, 3, 4, 1, true>(mA, mW, mO, M, N, R, ep, cs, ksplit)
              : launch_gemm_ws<128, 4, 4, 1, true>(mA, mW, mO, M, N, R, ep, cs, ksplit);
}

// Input gradient without a transposed weight copy: out[M, N] fp16 = epilogue(A[M, R] W[R, N]), A fp16 row-major (dY), W fp16
// row-major [R, N] = the Linear's own [out_features, in_features] weight.  epilogue 0 (fp16 store) or 8 (GELU': out =
// fp16(acc) * gelu_tanh'(gate[m, n]), gate = the saved fp16 pre-activation [M, gate_stride]).  N, R multiples of 8.
extern "C" GVF_API int gvf_gemm_nn_f16(const void* A, int lda, const void* W, int ldw, int M, int N, int R, int epilogue, void* out,
                                       int ldo, const void* gate, int gate_stride, void* stream) {
  if (!A || !W || !out || M <= 0 || N <= 0 || R <= 0 || (epilogue != 0 && epilogue != 8)) return GVF_ERR_INVALID;
  if ((N % 8) || (R % 8) || (lda % 8) || (ldw % 8) || (ldo % 8) || lda < R || ldw < N || ldo < N) return GVF_ERR_INVALID;
  if (epilogue == 8 && (!gate || (gate_stride % 8) || ((uintptr_t)gate & 15))) return GVF_ERR_INVALID;
  if (((uintptr_t)A | (uintptr_t)W | (uintptr_t)out) & 15) return GVF_ERR_INVALID;
  const bool wide = (N % 256) == 0;
  const int BN = wide ? 256 : 128;
  CUtensorMap mA, mW, mO;
  const uint64_t dA[2] = {(uint64_t)R, (uint64_t)M}, sA[2] = {1, (uint64_t)lda};
  const uint64_t dW[2] = {(uint64_t)N, (uint64_t)R}, sW[2] = {1, (uint64_t)ldw};
  const uint32_t bA[2] = {kBK, kBM}, bW[2] = {64, kBK};
  if (!make_tmap_f16(&mA, A, 2, dA, sA, bA, CU_TENSOR_MAP_SWIZZLE_128B)) return GVF_ERR_CUDA;
  if (!make_tmap_f16(&mW, W, 2, dW, sW, bW, CU_TENSOR_MAP_SWIZZLE_128B)) return GVF_ERR_CUDA;
  if (!make_tmap_2d(&mO, out, 2, (uint64_t)N, (uint64_t)M, (uint64_t)ldo, 64, 32)) return GVF_ERR_CUDA;
  GemmEpi ep;
  ep.mode = epilogue; ep.bias = nullptr; ep.out = out; ep.gate = (const __half*)gate; ep.gate_stride = gate_stride;
  ep.rows_per_batch = 1; ep.ldo = ldo; ep.gamma_q = nullptr; ep.gamma_k = nullptr; ep.norm_cols = 0;
  cudaStream_t cs = (cudaStream_t)stream;
  if (epilogue == 8)
    return wide ? launch_gemm_ws<256, 3, 8, 1, false, false, true>(mA, mW, mO, M, N, R, ep, cs)
                : launch_gemm_ws<128, 4, 8, 1, false, false, true>(mA, mW, mO, M, N, R, ep, cs);
  return wide ? launch_gemm_ws<256, 3, 0, 1, false, false, true>(mA, mW, mO, M, N, R, ep, cs)
              : launch_gemm_ws<128, 4, 0, 1, false, false, true>(mA, mW, mO, M, N, R, ep, cs);
}

// Submanifold sparse convolution as ONE kernel (reference sparse/conv/conv_spconv.py:6-15 ->
// spconv.SubMConv3d): out[n, :] = epilogue(sum_k W[:, k, :] x[nbr[n, k]] + bias).  x fp16 [N, Cin] (row stride ldx), nbr
// int32 [N, K3] from gvf_sparse_neighbor_map (-1 = absent), W fp16 [Cout, K3 * Cin].  The im2col operand of
// gvf_sparse_im2col_f16 + gvf_gemm_f16 is never written: the GEMM's TMA producer gathers the neighbour rows itself
// (tile::gather4).  Cin % 64 == 0; epilogue 0 (fp16 store), 3 (fp16 residual: out += fp16(conv + bias), in place) or 4 (fp32
// store).  Bit-identical to the two-kernel path.
extern "C" GVF_API int gvf_sparse_conv_gemm_f16(const void* x, int ldx, const int* nbr, int N, int K3, int Cin, const void* W,
                                                int ldw, int Cout, const float* bias, void* out, int ldo, int epilogue,
                                                void* stream) {
  if (!x || !nbr || !W || !out || N <= 0 || K3 <= 0 || Cin <= 0 || Cout <= 0) return GVF_ERR_INVALID;
  if ((Cin % 64) || (Cout % 8) || (ldx % 8) || (ldw % 8) || ldx < Cin || ldw < K3 * Cin) return GVF_ERR_UNSUPPORTED;
  if (epilogue != 0 && epilogue != 3 && epilogue != 4) return GVF_ERR_UNSUPPORTED;
  if (((uintptr_t)x | (uintptr_t)W | (uintptr_t)out) & 15) return GVF_ERR_INVALID;
  if ((ldo * (epilogue == 4 ? 4 : 2)) % 16) return GVF_ERR_INVALID;
  const int K = K3 * Cin;
  CUtensorMap mA, mW, mO;
  const uint64_t dA[2] = {(uint64_t)Cin, (uint64_t)N}, sA[2] = {1, (uint64_t)ldx};
  const uint64_t dW[2] = {(uint64_t)K, (uint64_t)Cout}, sW[2] = {1, (uint64_t)ldw};
  const bool wide = (Cout % 256) == 0;
  const uint32_t bA[2] = {kBK, 1}, bW[2] = {kBK, (uint32_t)(wide ? 256 : 128)};
  if (!make_tmap_f16(&mA, x, 2, dA, sA, bA, CU_TENSOR_MAP_SWIZZLE_128B)) return GVF_ERR_CUDA;
  if (!make_tmap_f16(&mW, W, 2, dW, sW, bW, CU_TENSOR_MAP_SWIZZLE_128B)) return GVF_ERR_CUDA;
  if (!make_tmap_2d(&mO, out, epilogue == 4 ? 4 : 2, (uint64_t)Cout, (uint64_t)N, (uint64_t)ldo, epilogue == 4 ? 32 : 64, 32))
    return GVF_ERR_CUDA;
  GemmEpi ep;
  ep.mode = epilogue; ep.bias = bias; ep.out = out; ep.gate = nullptr; ep.gate_stride = 0; ep.rows_per_batch = 1; ep.ldo = ldo;
  ep.gamma_q = nullptr; ep.gamma_k = nullptr; ep.norm_cols = 0;
  ep.gather_idx = nbr; ep.gather_k3 = K3; ep.gather_cb = Cin / kBK;
  cudaStream_t cs = (cudaStream_t)stream;
  if (epilogue == 0)
    return wide ? launch_gemm_ws<256, 3, 0, 1, false, true>(mA, mW, mO, N, Cout, K, ep, cs)
                : launch_gemm_ws<128, 4, 0, 1, false, true>(mA, mW, mO, N, Cout, K, ep, cs);
  if (epilogue == 3)      // out (fp16, pre-filled with the skip path) += fp16(conv): SparseResBlock3d's `h + skip_connection(x)`
    return wide ? launch_gemm_ws<256, 3, 3, 1, false, true>(mA, mW, mO, N, Cout, K, ep, cs)
                : launch_gemm_ws<128, 4, 3, 1, false, true>(mA, mW, mO, N, Cout, K, ep, cs);
  return wide ? launch_gemm_ws<256, 3, 4, 1, false, true>(mA, mW, mO, N, Cout, K, ep, cs)
              : launch_gemm_ws<128, 4, 4, 1, false, true>(mA, mW, mO, N, Cout, K, ep, cs);
}
