// attn_bwd.cu -- dense multi-head attention BACKWARD on tcgen05 tensor cores (sm_100a).
//
// Replaces the backward of `flash_attn.flash_attn_func` as reached through autograd from the reference's training
// step (train_vae.py:293-353 -> model/autoencoder.py:109-163 `Attention`, head dim 64; BASELINE configs[2] and [4]).
// fp16 operands, fp32 accumulation and statistics; the forward (csrc/attn.cu, gvf_attn_fwd_lse_f16) leaves
// LSE2[b,h,q] = log2(sum_k exp(scale * s_qk)) per query row, so P is recomputed exactly (no running maximum here).
//
//   D[q]  = sum_d dO[q,d] O[q,d]                                   (attn_bwd_prep_kernel, HBM pass)
//   P     = exp2(scale*log2e * Q K^T - LSE2),   dP = dO V^T,   dS = P o (dP - D)
//   dV    = P^T dO,   dK = scale * dS^T Q,   dQ = scale * dS K
//
// Two kernels, no atomics, deterministic:
//   attn_bwd_dkdv_kernel   CTA = (batch, head, 256 keys = two 128-key tiles); walks all query blocks.  Tensor-core
//                          work is done TRANSPOSED so that the accumulator rows (TMEM lanes) are keys:
//                            S^T = K Q^T, dP^T = V dO^T   (SS, both operands K-major from TMA tiles)
//                            dV += P^T dO, dK += dS^T Q   (TS: A = P^T / dS^T written to TMEM as fp16 by the
//                                                          softmax threads, B = the dO / Q block MN-major)
//   attn_bwd_dq_kernel     CTA = (batch, head, 256 queries = two tiles); walks all key blocks (and, when q is shared
//                          by the batch as in the motion-VAE decoder, all batch entries: dQ sums over them)
//                            S = Q K^T, dP = dO V^T (SS);  dQ += dS K (TS, B = K block MN-major)
//   Both: 12 warps like the forward -- warp 0 TMA producer, warps 1/2 single-thread tcgen05.mma issuers of tile A/B,
//   warp 3 TMEM allocation, warps 4-7 / 8-11 the element-wise warpgroups (thread = accumulator row = TMEM lane).
//   The fp16 P / dS tiles overwrite the fp32 S / dP columns they were computed from.
// Zero-filled TMA tails need no masks: padded key rows have K = V = 0 (their dS meets K = 0 in dQ, and their dK / dV
// rows are never stored); padded query rows have LSE2 = +inf (P = 0) and D = 0 by construction of the buffers.
#include "../../include/gvf_b200.h"
#include "tc_common.cuh"
#include "tma_host.h"

namespace gvf {
using namespace tc;

__device__ __forceinline__ float bwd_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct AttnBwdArgs {
  __half *dq, *dk, *dv;
  long long dq_sb, dq_sl, dq_sh, dk_sb, dk_sl, dk_sh, dv_sb, dv_sl, dv_sh;   // element strides (batch, seq, head)
  const float* lse2;      // [Nb, H, Lqp]
  const float* dsum;      // [Nb, H, Lqp]
  int Lq, Lk, H, Lqp, Nb;
  int q_batch_mul, kv_batch_mul;   // 0: tensor shared across the batch
  float scale, scale_log2e;
  // 1: the issuer waits for the TS MMAs of block i (which read the fp16 P / dS columns) before it issues the SS MMAs
  // of block i + 1 (which overwrite them as fp32 S / dP).  0 (default): issue them back to back -- tcgen05.mma
  // instructions of one thread execute in issue order, so the hazard is resolved inside the tensor pipe and one
  // commit -> mbarrier -> wait hop per block disappears (gvf_attn_bwd_set_serial for A/B runs).
  int serial;
};

// D[b,h,q] = sum_d dO O (fp32); rows q in [Lq, Lqp) are zeroed.  One thread per (b, q, h).
template <int D>
__global__ void __launch_bounds__(256) attn_bwd_prep_kernel(const __half* __restrict__ o, const __half* __restrict__ dout,
                                                            long long o_sb, long long o_sl, long long o_sh,
                                                            long long g_sb, long long g_sl, long long g_sh, int Nb,
                                                            int Lq, int Lqp, int H, float* __restrict__ dsum) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)Nb * Lqp * H;
  if (gid >= n) return;
  const int h = (int)(gid % H);
  const int q = (int)((gid / H) % Lqp);
  const int b = (int)(gid / ((long long)H * Lqp));
  float s = 0.f;
  if (q < Lq) {
    const uint4* po = reinterpret_cast<const uint4*>(o + b * o_sb + q * o_sl + h * o_sh);
    const uint4* pg = reinterpret_cast<const uint4*>(dout + b * g_sb + q * g_sl + h * g_sh);
#pragma unroll
    for (int j = 0; j < D / 8; ++j) {
      const uint4 a = __ldg(po + j), g = __ldg(pg + j);
      const __half2* a2 = reinterpret_cast<const __half2*>(&a);
      const __half2* g2 = reinterpret_cast<const __half2*>(&g);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 x = __half22float2(a2[t]), y = __half22float2(g2[t]);
        s = fmaf(x.x, y.x, s);
        s = fmaf(x.y, y.y, s);
      }
    }
  }
  dsum[((long long)b * H + h) * Lqp + q] = s;
}

constexpr int kBwdQStages = 3;     // dK/dV kernel: ring of (Q tile, dO tile, LSE2, D) stages
constexpr int kBwdKVStages = 3;    // dQ kernel: ring of (K tile, V tile) stages

// ------------------------------------------------------------------------------------------------ dK / dV
template <int D>
__global__ void __launch_bounds__(384, 1)
attn_bwd_dkdv_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                     const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapDO,
                     const AttnBwdArgs a) {
  constexpr int ROWB = D * 2, TILE_BYTES = 128 * ROWB;
  constexpr uint64_t SWZ = (D == 32) ? SWZ_64B : SWZ_128B;
  constexpr uint32_t SBO = 8 * ROWB;
  constexpr int S = kBwdQStages;
  constexpr int STAGE_BYTES = 2 * TILE_BYTES + 1024;          // Q, dO, LSE2[128], D[128]
  constexpr uint32_t TM_S = 0, TM_DP = 64, TM_DV = 128, TM_DK = 128 + D, TM_STRIDE = 128 + 2 * D;
  static_assert(2 * TM_STRIDE <= 512, "TMEM budget");

  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t kv_full, q_full[S], q_empty[S], s_full[2], p_full[2], o_done[2];
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sK = smem;                                  // 2 tiles
  uint8_t* sV = smem + 2 * TILE_BYTES;                 // 2 tiles
  uint8_t* sQ = smem + 4 * TILE_BYTES;                 // S stages

  const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
  const int kblk = blockIdx.x, h = blockIdx.y, nb = blockIdx.z;
  const int k0 = kblk * 256;
  const int nkt = (a.Lk - k0 > 128) ? 2 : 1;           // live key tiles of this CTA
  const int n_qt = (a.Lq + 127) / 128;                 // query tiles (ring items)
  const int n_blk = 2 * n_qt;                          // 64-query blocks (padded rows are inert)

  if (threadIdx.x == 0) {
    mbar_init(&kv_full, 1);
    for (int s = 0; s < S; ++s) { mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], nkt * 5); }     // 1 commit + 4 warps per tile
    for (int x = 0; x < 2; ++x) { mbar_init(&s_full[x], 1); mbar_init(&p_full[x], 4); mbar_init(&o_done[x], 1); }
    fence_barrier_init();
    tma_prefetch_desc(&mapQ); tma_prefetch_desc(&mapK); tma_prefetch_desc(&mapV); tma_prefetch_desc(&mapDO);
  }
  if (warp == 3) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&kv_full, nkt * 2 * TILE_BYTES);
      for (int x = 0; x < nkt; ++x) {
        tma_load_4d(sK + x * TILE_BYTES, &mapK, &kv_full, 0, h, k0 + x * 128, nb * a.kv_batch_mul);
        tma_load_4d(sV + x * TILE_BYTES, &mapV, &kv_full, 0, h, k0 + x * 128, nb * a.kv_batch_mul);
      }
      const float* lse = a.lse2 + ((long long)nb * a.H + h) * a.Lqp;
      const float* dsm = a.dsum + ((long long)nb * a.H + h) * a.Lqp;
      for (int j = 0; j < n_qt; ++j) {
        const int s = j % S;
        mbar_wait(&q_empty[s], ((j / S) & 1) ^ 1);
        uint8_t* st = sQ + s * STAGE_BYTES;
        mbar_arrive_expect_tx(&q_full[s], STAGE_BYTES);
        tma_load_4d(st, &mapQ, &q_full[s], 0, h, j * 128, nb * a.q_batch_mul);
        tma_load_4d(st + TILE_BYTES, &mapDO, &q_full[s], 0, h, j * 128, nb);
        bulk_load_1d(st + 2 * TILE_BYTES, lse + j * 128, 512, &q_full[s]);
        bulk_load_1d(st + 2 * TILE_BYTES + 512, dsm + j * 128, 512, &q_full[s]);
      }
    }
  } else if (warp == 1 || warp == 2) {
    // MMA issuer of key tile x: the loop runs with all 32 lanes on warp-uniform values and only the tcgen05 instructions
    // sit behind elect.sync -- inside `if (lane == 0)` ptxas wraps every tcgen05.mma / commit in an ELECT / R2UR /
    // BRA.U.ANY loop with dependent descriptor arithmetic (the finding behind attention v8, csrc/attn.cu)
    const int x = warp - 1;
    if (x < nkt) {
      const uint32_t idesc_s = make_idesc_f16(128, 64, 0, 0);
      const uint32_t idesc_o = make_idesc_f16(128, D, 0, 1);
      const uint32_t ka = smem_u32(sK) + x * TILE_BYTES, va = smem_u32(sV) + x * TILE_BYTES;
      const uint32_t tX = tmem + x * TM_STRIDE;
      mbar_wait(&kv_full, 0);
      tc_fence_after();
      for (int i = 0; i < n_blk; ++i) {
        const int j = i >> 1, s = j % S;
        if ((i & 1) == 0) { mbar_wait(&q_full[s], (j / S) & 1); tc_fence_after(); }
        if (i > 0 && a.serial) { mbar_wait(&o_done[x], (i - 1) & 1); tc_fence_after(); }   // P / dS columns are free again
        const uint32_t qa = smem_u32(sQ) + s * STAGE_BYTES + (i & 1) * 64 * ROWB;
        const uint32_t da = qa + TILE_BYTES;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < D / 16; ++k)             // S^T = K Q^T
            mma_ss(tX + TM_S, make_smem_desc(ka + k * 32, 16, SBO, SWZ), make_smem_desc(qa + k * 32, 16, SBO, SWZ),
                   idesc_s, k != 0);
#pragma unroll
          for (int k = 0; k < D / 16; ++k)             // dP^T = V dO^T
            mma_ss(tX + TM_DP, make_smem_desc(va + k * 32, 16, SBO, SWZ), make_smem_desc(da + k * 32, 16, SBO, SWZ),
                   idesc_s, k != 0);
          tc_commit(&s_full[x]);
        }
        __syncwarp();
        mbar_wait(&p_full[x], i & 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)                  // dV += P^T dO   (K = 64 queries)
            mma_ts(tX + TM_DV, tX + TM_S + k * 8, make_smem_desc(da + k * 16 * ROWB, SBO, SBO, SWZ), idesc_o,
                   (i | k) != 0);
#pragma unroll
          for (int k = 0; k < 4; ++k)                  // dK += dS^T Q
            mma_ts(tX + TM_DK, tX + TM_DP + k * 8, make_smem_desc(qa + k * 16 * ROWB, SBO, SBO, SWZ), idesc_o,
                   (i | k) != 0);
          tc_commit(&o_done[x]);
          if (i & 1) tc_commit(&q_empty[s]);           // this tile's operands of the stage are consumed
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    const int x = (warp - 4) >> 2;
    if (x < nkt) {
      const int quarter = warp & 3;
      const int row = quarter * 32 + lane;             // key row inside the tile == TMEM lane
      const uint32_t tX = tmem + x * TM_STRIDE + ((uint32_t)(quarter * 32) << 16);
      const float c = a.scale_log2e;
      for (int i = 0; i < n_blk; ++i) {
        const int j = i >> 1, s = j % S;
        const float* st = reinterpret_cast<const float*>(sQ + s * STAGE_BYTES + 2 * TILE_BYTES) + (i & 1) * 64;
        if ((i & 1) == 0) mbar_wait(&q_full[s], (j / S) & 1);   // LSE2 / D of the stage (bulk copies) are visible
        mbar_wait(&s_full[x], i & 1);
        tc_fence_after();
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t sv[32], dp[32];
          tmem_ld_x32(tX + TM_S + 32 * hf, sv);
          tmem_ld_x32(tX + TM_DP + 32 * hf, dp);
          tmem_ld_wait();
          uint32_t pp[16], ds[16];
#pragma unroll
          for (int k = 0; k < 32; k += 4) {
            const float4 l4 = *reinterpret_cast<const float4*>(st + 32 * hf + k);
            const float4 d4 = *reinterpret_cast<const float4*>(st + 128 + 32 * hf + k);
            const float p0 = bwd_exp2(fmaf(__uint_as_float(sv[k]), c, -l4.x));
            const float p1 = bwd_exp2(fmaf(__uint_as_float(sv[k + 1]), c, -l4.y));
            const float p2 = bwd_exp2(fmaf(__uint_as_float(sv[k + 2]), c, -l4.z));
            const float p3 = bwd_exp2(fmaf(__uint_as_float(sv[k + 3]), c, -l4.w));
            const float e0 = p0 * (__uint_as_float(dp[k]) - d4.x), e1 = p1 * (__uint_as_float(dp[k + 1]) - d4.y);
            const float e2 = p2 * (__uint_as_float(dp[k + 2]) - d4.z), e3 = p3 * (__uint_as_float(dp[k + 3]) - d4.w);
            __half2 t;
            t = __floats2half2_rn(p0, p1); pp[k >> 1] = *reinterpret_cast<uint32_t*>(&t);
            t = __floats2half2_rn(p2, p3); pp[(k >> 1) + 1] = *reinterpret_cast<uint32_t*>(&t);
            t = __floats2half2_rn(e0, e1); ds[k >> 1] = *reinterpret_cast<uint32_t*>(&t);
            t = __floats2half2_rn(e2, e3); ds[(k >> 1) + 1] = *reinterpret_cast<uint32_t*>(&t);
          }
          // fp16 pairs of queries [32 hf, 32 hf + 32) -> 16 columns; they land on S / dP columns already read
          tmem_st_x16(tX + TM_S + 16 * hf, pp);
          tmem_st_x16(tX + TM_DP + 16 * hf, ds);
        }
        tmem_st_wait();                                // warp-wide: every lane's P / dS columns are written
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&p_full[x]);
          if (i & 1) mbar_arrive(&q_empty[s]);         // LSE2 / D of this stage are no longer needed by this warp
        }
      }
      mbar_wait(&o_done[x], (n_blk - 1) & 1);
      tc_fence_after();
      const int ki = k0 + x * 128 + row;
#pragma unroll
      for (int which = 0; which < 2; ++which) {        // 0: dV, 1: dK (scaled)
        const float mul = which ? a.scale : 1.0f;
        __half* base = which ? a.dk + (long long)nb * a.dk_sb + (long long)ki * a.dk_sl + (long long)h * a.dk_sh
                             : a.dv + (long long)nb * a.dv_sb + (long long)ki * a.dv_sl + (long long)h * a.dv_sh;
#pragma unroll
        for (int d0 = 0; d0 < D; d0 += 32) {
          uint32_t r[32];
          tmem_ld_x32(tX + (which ? TM_DK : TM_DV) + d0, r);
          tmem_ld_wait();
          if (ki < a.Lk) {
#pragma unroll
            for (int k = 0; k < 32; k += 8) {
              __align__(16) __half hh[8];
#pragma unroll
              for (int t = 0; t < 8; ++t) hh[t] = __float2half_rn(__uint_as_float(r[k + t]) * mul);
              *reinterpret_cast<uint4*>(base + d0 + k) = *reinterpret_cast<uint4*>(hh);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 3) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------ dQ
template <int D>
__global__ void __launch_bounds__(384, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                   const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapDO,
                   const AttnBwdArgs a, int reps) {
  constexpr int ROWB = D * 2, TILE_BYTES = 128 * ROWB;
  constexpr uint64_t SWZ = (D == 32) ? SWZ_64B : SWZ_128B;
  constexpr uint32_t SBO = 8 * ROWB;
  constexpr int S = kBwdKVStages;
  constexpr int DO_BYTES = 2 * TILE_BYTES + 2048;             // dO tiles A/B, LSE2[256], D[256]
  constexpr uint32_t TM_S = 0, TM_DP = 64, TM_DQ = 128, TM_STRIDE = 128 + D;

  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t q_full, do_full[2], do_empty[2], kv_full[S], kv_empty[S], s_full[2], p_full[2], dq_done[2];
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                                  // 2 tiles
  uint8_t* sDO = smem + 2 * TILE_BYTES;                // 2 stages
  uint8_t* sKV = sDO + 2 * DO_BYTES;                   // S x (K tile, V tile)

  const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
  const int qblk = blockIdx.x, h = blockIdx.y, nbq = blockIdx.z;
  const int q0 = qblk * 256;
  const int nq = (a.Lq - q0 > 128) ? 2 : 1;
  const int n_kv = (a.Lk + 127) / 128;                 // K/V tiles per batch entry
  const int n_blk = (a.Lk + 63) / 64;                  // 64-key blocks per batch entry

  if (threadIdx.x == 0) {
    mbar_init(&q_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&do_full[s], 1); mbar_init(&do_empty[s], nq * 5); }
    for (int s = 0; s < S; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], nq); }
    for (int x = 0; x < 2; ++x) { mbar_init(&s_full[x], 1); mbar_init(&p_full[x], 4); mbar_init(&dq_done[x], 1); }
    fence_barrier_init();
    tma_prefetch_desc(&mapQ); tma_prefetch_desc(&mapK); tma_prefetch_desc(&mapV); tma_prefetch_desc(&mapDO);
  }
  if (warp == 3) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&q_full, nq * TILE_BYTES);
      for (int x = 0; x < nq; ++x)
        tma_load_4d(sQ + x * TILE_BYTES, &mapQ, &q_full, 0, h, q0 + x * 128, nbq * a.q_batch_mul);
      int t = 0;                                       // running K/V tile counter
      for (int r = 0; r < reps; ++r) {
        const int nb = (a.q_batch_mul ? nbq : r);      // batch entry of dO / K / V for this pass
        const int ds = r & 1;
        mbar_wait(&do_empty[ds], ((r >> 1) & 1) ^ 1);
        uint8_t* st = sDO + ds * DO_BYTES;
        mbar_arrive_expect_tx(&do_full[ds], nq * TILE_BYTES + 2 * nq * 512);
        const float* lse = a.lse2 + ((long long)nb * a.H + h) * a.Lqp + q0;
        const float* dsm = a.dsum + ((long long)nb * a.H + h) * a.Lqp + q0;
        for (int x = 0; x < nq; ++x) {
          tma_load_4d(st + x * TILE_BYTES, &mapDO, &do_full[ds], 0, h, q0 + x * 128, nb);
          bulk_load_1d(st + 2 * TILE_BYTES + x * 512, lse + x * 128, 512, &do_full[ds]);
          bulk_load_1d(st + 2 * TILE_BYTES + 1024 + x * 512, dsm + x * 128, 512, &do_full[ds]);
        }
        for (int j = 0; j < n_kv; ++j, ++t) {
          const int s = t % S;
          mbar_wait(&kv_empty[s], ((t / S) & 1) ^ 1);
          mbar_arrive_expect_tx(&kv_full[s], 2 * TILE_BYTES);
          tma_load_4d(sKV + s * 2 * TILE_BYTES, &mapK, &kv_full[s], 0, h, j * 128, nb * a.kv_batch_mul);
          tma_load_4d(sKV + s * 2 * TILE_BYTES + TILE_BYTES, &mapV, &kv_full[s], 0, h, j * 128, nb * a.kv_batch_mul);
        }
      }
    }
  } else if (warp == 1 || warp == 2) {
    const int x = warp - 1;                            // warp-uniform issue loop, see the dK/dV kernel
    if (x < nq) {
      const uint32_t idesc_s = make_idesc_f16(128, 64, 0, 0);
      const uint32_t idesc_o = make_idesc_f16(128, D, 0, 1);
      const uint32_t qa = smem_u32(sQ) + x * TILE_BYTES;
      const uint32_t tX = tmem + x * TM_STRIDE;
      mbar_wait(&q_full, 0);
      tc_fence_after();
      int g = 0, t = 0;                                // running block / tile counters
      for (int r = 0; r < reps; ++r) {
        const int ds = r & 1;
        mbar_wait(&do_full[ds], (r >> 1) & 1);
        tc_fence_after();
        const uint32_t da = smem_u32(sDO) + ds * DO_BYTES + x * TILE_BYTES;
        for (int i = 0; i < n_blk; ++i, ++g) {
          const int s = t % S;
          if ((i & 1) == 0) { mbar_wait(&kv_full[s], (t / S) & 1); tc_fence_after(); }
          if (g > 0 && a.serial) { mbar_wait(&dq_done[x], (g - 1) & 1); tc_fence_after(); }   // dS columns are free again
          const uint32_t ka = smem_u32(sKV) + s * 2 * TILE_BYTES + (i & 1) * 64 * ROWB;
          const uint32_t va = ka + TILE_BYTES;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < D / 16; ++k)           // S = Q K^T
              mma_ss(tX + TM_S, make_smem_desc(qa + k * 32, 16, SBO, SWZ), make_smem_desc(ka + k * 32, 16, SBO, SWZ),
                     idesc_s, k != 0);
#pragma unroll
            for (int k = 0; k < D / 16; ++k)           // dP = dO V^T
              mma_ss(tX + TM_DP, make_smem_desc(da + k * 32, 16, SBO, SWZ), make_smem_desc(va + k * 32, 16, SBO, SWZ),
                     idesc_s, k != 0);
            tc_commit(&s_full[x]);
          }
          __syncwarp();
          mbar_wait(&p_full[x], g & 1);
          tc_fence_after();
          const bool tile_done = (i & 1) || i + 1 == n_blk;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)                // dQ += dS K   (K = 64 keys)
              mma_ts(tX + TM_DQ, tX + TM_S + k * 8, make_smem_desc(ka + k * 16 * ROWB, SBO, SBO, SWZ), idesc_o,
                     (g | k) != 0);
            tc_commit(&dq_done[x]);
            if (tile_done) tc_commit(&kv_empty[s]);
            if (i + 1 == n_blk) tc_commit(&do_empty[ds]);
          }
          __syncwarp();
          if (tile_done) ++t;
        }
      }
    }
  } else if (warp >= 4) {
    const int x = (warp - 4) >> 2;
    if (x < nq) {
      const int quarter = warp & 3;
      const int row = quarter * 32 + lane;
      const uint32_t tX = tmem + x * TM_STRIDE + ((uint32_t)(quarter * 32) << 16);
      const float c = a.scale_log2e;
      int g = 0;
      for (int r = 0; r < reps; ++r) {
        const int ds = r & 1;
        mbar_wait(&do_full[ds], (r >> 1) & 1);
        const float* st = reinterpret_cast<const float*>(sDO + ds * DO_BYTES + 2 * TILE_BYTES);
        const float lse = st[x * 128 + row], dsm = st[256 + x * 128 + row];
        __syncwarp();
        if (lane == 0) mbar_arrive(&do_empty[ds]);
        for (int i = 0; i < n_blk; ++i, ++g) {
          mbar_wait(&s_full[x], g & 1);
          tc_fence_after();
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            uint32_t sv[32], dp[32];
            tmem_ld_x32(tX + TM_S + 32 * hf, sv);
            tmem_ld_x32(tX + TM_DP + 32 * hf, dp);
            tmem_ld_wait();
            uint32_t dsv[16];
#pragma unroll
            for (int k = 0; k < 32; k += 2) {
              const float p0 = bwd_exp2(fmaf(__uint_as_float(sv[k]), c, -lse));
              const float p1 = bwd_exp2(fmaf(__uint_as_float(sv[k + 1]), c, -lse));
              const __half2 t2 = __floats2half2_rn(p0 * (__uint_as_float(dp[k]) - dsm),
                                                   p1 * (__uint_as_float(dp[k + 1]) - dsm));
              dsv[k >> 1] = *reinterpret_cast<const uint32_t*>(&t2);
            }
            tmem_st_x16(tX + TM_S + 16 * hf, dsv);
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[x]);
        }
      }
      mbar_wait(&dq_done[x], (g - 1) & 1);
      tc_fence_after();
      const int qi = q0 + x * 128 + row;
      __half* base = a.dq + (long long)nbq * a.dq_sb + (long long)qi * a.dq_sl + (long long)h * a.dq_sh;
#pragma unroll
      for (int d0 = 0; d0 < D; d0 += 32) {
        uint32_t rr[32];
        tmem_ld_x32(tX + TM_DQ + d0, rr);
        tmem_ld_wait();
        if (qi < a.Lq) {
#pragma unroll
          for (int k = 0; k < 32; k += 8) {
            __align__(16) __half hh[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) hh[t] = __float2half_rn(__uint_as_float(rr[k + t]) * a.scale);
            *reinterpret_cast<uint4*>(base + d0 + k) = *reinterpret_cast<uint4*>(hh);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 3) tmem_dealloc(tmem, 512);
}

template <int D>
static int launch_attn_bwd(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const CUtensorMap& mdo,
                           const AttnBwdArgs& a, int q_shared, cudaStream_t st) {
  constexpr int TILE = 128 * D * 2;
  constexpr int SMEM_KV = 4 * TILE + kBwdQStages * (2 * TILE + 1024) + 1024;
  constexpr int SMEM_Q = 2 * TILE + 2 * (2 * TILE + 2048) + kBwdKVStages * 2 * TILE + 1024;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(attn_bwd_dkdv_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_KV) != cudaSuccess ||
        cudaFuncSetAttribute(attn_bwd_dq_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_Q) != cudaSuccess)
      return GVF_ERR_CUDA;
    configured = true;
  }
  attn_bwd_dkdv_kernel<D><<<dim3((a.Lk + 255) / 256, a.H, a.Nb), 384, SMEM_KV, st>>>(mq, mk, mv, mdo, a);
  if (cudaGetLastError() != cudaSuccess) return GVF_ERR_CUDA;
  attn_bwd_dq_kernel<D><<<dim3((a.Lq + 255) / 256, a.H, q_shared ? 1 : a.Nb), 384, SMEM_Q, st>>>(mq, mk, mv, mdo, a,
                                                                                                  q_shared ? a.Nb : 1);
  return cudaGetLastError() == cudaSuccess ? GVF_OK : GVF_ERR_CUDA;
}

}  // namespace gvf

using namespace gvf;

static int g_attn_bwd_serial = 0;
extern "C" GVF_API void gvf_attn_bwd_set_serial(int v) { g_attn_bwd_serial = v; }

extern "C" GVF_API int gvf_attn_bwd_f16(const void* q, const void* k, const void* v, const void* o, const void* dout,
                                        const float* lse2, float* dsum, void* dq, void* dk, void* dv, int Nb, int Lq,
                                        int Lk, int H, int D, int lse_ld, const long long* q_strides,
                                        const long long* k_strides, const long long* v_strides,
                                        const long long* o_strides, const long long* do_strides,
                                        const long long* dq_strides, const long long* dk_strides,
                                        const long long* dv_strides, int q_shared, float scale, void* stream) {
  if (!q || !k || !v || !o || !dout || !lse2 || !dsum || !dq || !dk || !dv) return GVF_ERR_INVALID;
  if (!q_strides || !k_strides || !v_strides || !o_strides || !do_strides || !dq_strides || !dk_strides || !dv_strides)
    return GVF_ERR_INVALID;
  if (Nb <= 0 || Lq <= 0 || Lk <= 0 || H <= 0) return GVF_ERR_INVALID;
  if (D != 32 && D != 64) return GVF_ERR_UNSUPPORTED;
  if (lse_ld < ((Lq + 127) / 128) * 128 || (lse_ld % 128)) return GVF_ERR_INVALID;
  for (int i = 0; i < 3; ++i)
    if ((q_strides[i] % 8) || (k_strides[i] % 8) || (v_strides[i] % 8) || (o_strides[i] % 8) || (do_strides[i] % 8) ||
        (dq_strides[i] % 8) || (dk_strides[i] % 8) || (dv_strides[i] % 8))
      return GVF_ERR_INVALID;
  if (((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)o | (uintptr_t)dout | (uintptr_t)dq | (uintptr_t)dk |
       (uintptr_t)dv | (uintptr_t)lse2 | (uintptr_t)dsum) & 15)
    return GVF_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  {
    const long long n = (long long)Nb * lse_ld * H;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (D == 64)
      attn_bwd_prep_kernel<64><<<blocks, 256, 0, st>>>((const __half*)o, (const __half*)dout, o_strides[0], o_strides[1],
                                                        o_strides[2], do_strides[0], do_strides[1], do_strides[2], Nb, Lq,
                                                        lse_ld, H, dsum);
    else
      attn_bwd_prep_kernel<32><<<blocks, 256, 0, st>>>((const __half*)o, (const __half*)dout, o_strides[0], o_strides[1],
                                                        o_strides[2], do_strides[0], do_strides[1], do_strides[2], Nb, Lq,
                                                        lse_ld, H, dsum);
    if (cudaGetLastError() != cudaSuccess) return GVF_ERR_CUDA;
  }
  CUtensorMap mq, mk, mv, mdo;
  const CUtensorMapSwizzle sw = swizzle_for_bytes(D * 2);
  const uint32_t box[4] = {(uint32_t)D, 1, 128, 1};
  auto mk_map = [&](CUtensorMap* m, const void* p, int L, int nbt, const long long* s) {
    const uint64_t dims[4] = {(uint64_t)D, (uint64_t)H, (uint64_t)L, (uint64_t)nbt};
    const uint64_t strides[4] = {1, (uint64_t)s[2], (uint64_t)s[1], (uint64_t)(nbt > 1 ? s[0] : (long long)L * s[1])};
    return make_tmap_f16(m, p, 4, dims, strides, box, sw);
  };
  if (!mk_map(&mq, q, Lq, q_shared ? 1 : Nb, q_strides)) return GVF_ERR_CUDA;
  if (!mk_map(&mk, k, Lk, Nb, k_strides)) return GVF_ERR_CUDA;
  if (!mk_map(&mv, v, Lk, Nb, v_strides)) return GVF_ERR_CUDA;
  if (!mk_map(&mdo, dout, Lq, Nb, do_strides)) return GVF_ERR_CUDA;
  AttnBwdArgs a;
  a.dq = (__half*)dq; a.dk = (__half*)dk; a.dv = (__half*)dv;
  a.dq_sb = dq_strides[0]; a.dq_sl = dq_strides[1]; a.dq_sh = dq_strides[2];
  a.dk_sb = dk_strides[0]; a.dk_sl = dk_strides[1]; a.dk_sh = dk_strides[2];
  a.dv_sb = dv_strides[0]; a.dv_sl = dv_strides[1]; a.dv_sh = dv_strides[2];
  a.lse2 = lse2; a.dsum = dsum;
  a.Lq = Lq; a.Lk = Lk; a.H = H; a.Lqp = lse_ld; a.Nb = Nb;
  a.q_batch_mul = q_shared ? 0 : 1; a.kv_batch_mul = 1;
  a.scale = scale; a.scale_log2e = scale * 1.4426950408889634f;
  a.serial = g_attn_bwd_serial;
  return D == 64 ? launch_attn_bwd<64>(mq, mk, mv, mdo, a, q_shared, st) : launch_attn_bwd<32>(mq, mk, mv, mdo, a, q_shared, st);
}
