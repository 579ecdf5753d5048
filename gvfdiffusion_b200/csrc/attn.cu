// attn.cu -- dense multi-head attention forward on tcgen05 tensor cores (sm_100a).
//
// Replaces `flash_attn.flash_attn_{func,kvpacked_func,qkvpacked_func}` (FA2, mma.sync) as
// reached from reference model/attention/full_attn.py:74-140 (DiT spatial / image-cross /
// static-cross attention, head dim 32) and model/autoencoder.py:132-144 (motion-VAE self /
// decoder cross attention, head dim 64).  fp16 in / out, fp32 softmax statistics, no mask,
// softmax scale passed in.  Layout [N, L, H, d] with arbitrary (16 B aligned) strides, so
// packed qkv / kv tensors and the temporal (strided) view are addressed in place by TMA.
//
// One CTA = one (batch, head) x up to two 128-row query tiles ("A" and "B", ping-pong):
//   (12 warps: 0 TMA, 1-2 MMA issuers, 3 TMEM allocator, 4-7 / 8-11 the two softmax warpgroups)
//   warp 0      TMA producer: Q tiles once, then K/V tiles [128 x d] through a 4-stage
//               mbarrier ring (SWIZZLE_64B for d=32, SWIZZLE_128B for d=64)
//   warps 1,2   single-thread tcgen05.mma issuers, one per query tile:
//                 S = Q K^T   (SS: A,B K-major from smem, M=128 N=64 K=d)  -> TMEM, 2 buffers
//                 O' = P V    (TS: A = P from TMEM, B = V rows MN-major, M=128 N=d K=64)
//               scores are computed two 64-column blocks ahead of the softmax (double-buffered S):
//               the MMA -> commit -> mbarrier -> tcgen05.ld hand-off latency (~1.5k cycles, measured:
//               the loop ran at 85 % of its time with all softmax math removed) is off the critical path
//   warps 4-7   softmax of tile A, warps 8-11 softmax of tile B: thread = one query row
//               (TMEM lane); tcgen05.ld S -> running max / exp2 / row sum in registers ->
//               P (fp16 pairs) to its own TMEM columns with tcgen05.st; the per-tile partial
//               product O' is read back from TMEM one iteration later and accumulated in
//               registers (O <- (O + O') * alpha), so no TMEM read-modify-write and no
//               correction warpgroup is needed for d <= 64.
// With d = 32 the MUFU exp2 (1 per score) bounds the kernel, not the MMA pipe: 128x128 scores
// cost 1024 MUFU cycles/SM vs 256 tensor cycles -- see DESIGN.md for the roofline.
#include "../../include/gvf_b200.h"
#include "tc_common.cuh"
#include "tma_host.h"
#include "launch.h"

namespace gvf {
using namespace tc;

constexpr int kAttnStages = 4;

// Packed fp32 pairs (FFMA2 / FADD2 / FMUL2, new on sm_100): one issue slot for two lanes of work.
// The softmax loop is issue- as much as MUFU-bound, so halving its FMA-pipe instruction count matters.
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {
  asm("{\n.reg .b64 ra, rb, rc, rd;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\nmov.b64 rc, {%6, %7};\n"
      "fma.rn.f32x2 rd, ra, rb, rc;\nmov.b64 {%0, %1}, rd;\n}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
__device__ __forceinline__ void fadd2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{\n.reg .b64 ra, rb, rd;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\n"
      "add.rn.f32x2 rd, ra, rb;\nmov.b64 {%0, %1}, rd;\n}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void fmul2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{\n.reg .b64 ra, rb, rd;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\n"
      "mul.rn.f32x2 rd, ra, rb;\nmov.b64 {%0, %1}, rd;\n}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}

__device__ __forceinline__ float fast_exp2(float x) {   // MUFU.EX2; exp2(-inf) = 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnArgs {
  __half* o;
  long long o_stride_b, o_stride_l, o_stride_h;   // elements
  int Lq, Lk, H;
  int q_batch_mul, kv_batch_mul;                  // 0: tensor shared across the batch
  float scale_log2e;
  long long* trace;   // optional clock64 trace of CTA (0,0,0) (tools/attn_experiments.py)
  int stagger;        // clocks tile B's softmax warpgroup starts late (see the softmax branch)
  int dbg;            // bit 0x20: MUFU ping-pong between the two softmax warpgroups (v4)
  // v6 work list: items [0, split_from) are whole (query block, head, batch) units; every unit from
  // split_from on is cut into `nsplit` key ranges whose partial (O, m, l) go to ws_o / ws_ml (see
  // attn_merge_kernel).  gx = query blocks per (head, batch).
  int split_from, nsplit, gx;
  float* ws_o;        // [split units][nsplit][512 rows][32]
  float* ws_ml;       // [split units][nsplit][512 rows][2]
  // training: LSE2[nb, h, q] = log2(sum_k exp(scale s_qk)) for the backward kernels (csrc/attn_bwd.cu); rows in
  // [Lq, lse_ld) get +inf.  Only attn_fwd_kernel writes it (gvf_attn_fwd_lse_f16 selects that kernel).
  float* lse = nullptr;
  int lse_ld = 0;
};

// POLY = n > 0: every n-th pair of exponentials is evaluated on the FMA pipe (Cody-Waite split plus a
// degree-3 polynomial, relative error 7.5e-5 -- a sixth of the fp16 rounding P gets anyway) instead of
// MUFU.EX2, which at 16 results / clk / SM is the unit that bounds this kernel at d = 32.
template <int D, int BLK, int NSBUF, int POLY>
__global__ void __launch_bounds__(384, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                const __grid_constant__ CUtensorMap mapV, const AttnArgs a) {
  constexpr int ROWB = D * 2;                      // bytes per smem row
  constexpr int TILE_BYTES = 128 * ROWB;           // one [128 x D] fp16 tile
  constexpr uint64_t SWZ = (D == 32) ? SWZ_64B : SWZ_128B;
  constexpr uint32_t SBO = 8 * ROWB;
  constexpr int S = kAttnStages;
  // BLK = key/value columns per softmax block (a whole 128-row TMA tile at d=32, half of one at d=64
  // where the 64 accumulator registers leave no room for a 128-wide score row), NSBUF score buffers.
  // TMEM columns per query tile: NSBUF x S (fp32), P (fp16 pairs) and the partial product O'
  constexpr int BPT = 128 / BLK;                   // blocks per TMA tile
  constexpr uint32_t TM_STRIDE = NSBUF * BLK + BLK / 2 + D;  // 224 for (32,128,1) and (64,64,2): 2 tiles <= 512
  constexpr uint32_t TM_P = NSBUF * BLK, TM_O = NSBUF * BLK + BLK / 2;
  static_assert(2 * TM_STRIDE <= 512, "TMEM budget");

  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t q_full, kv_full[S], kv_empty[S], s_full[2][2], s_free[2][2], p_full[2], o_full[2];
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                              // 2 tiles
  uint8_t* sKV = smem + 2 * TILE_BYTES;            // S x (K tile, V tile)

  const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
  const int qblk = blockIdx.x, h = blockIdx.y, nb = blockIdx.z;
  const int cta_lin = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  if (a.trace && threadIdx.x == 0 && cta_lin < 1024) {   // per-CTA (smid, start ns) for schedule studies
    unsigned long long t; unsigned sm;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    asm volatile("mov.u32 %0, %smid;" : "=r"(sm));
    a.trace[256 + 3 * cta_lin] = sm; a.trace[256 + 3 * cta_lin + 1] = (long long)t;
  }
  const int q0 = qblk * 256;
  const int nq = (a.Lq - q0 > 128) ? 2 : 1;
  const int n_kv = (a.Lk + 127) / 128;             // TMA tiles
  const int n_blk = (a.Lk + BLK - 1) / BLK;        // softmax blocks

  if (threadIdx.x == 0) {
    mbar_init(&q_full, 1);
    for (int s = 0; s < S; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], nq); }
    for (int x = 0; x < 2; ++x) {
      for (int b = 0; b < 2; ++b) { mbar_init(&s_full[x][b], 1); mbar_init(&s_free[x][b], 4); }   // one arrival per softmax warp
      mbar_init(&p_full[x], 4);
      mbar_init(&o_full[x], 1);
    }
    fence_barrier_init();
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapV);
  }
  if (warp == 3) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  // Register re-distribution between warpgroups (setmaxnreg): warps 0-3 only issue TMA / MMA and keep
  // 56 registers; the softmax warpgroups (warps 4-7, 8-11) grow to 216.  Each setmaxnreg sits at the top
  // of the branch it governs so ptxas budgets that region only.
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(&q_full, nq * TILE_BYTES);
      for (int x = 0; x < nq; ++x)
        tma_load_4d(sQ + x * TILE_BYTES, &mapQ, &q_full, 0, h, q0 + x * 128, nb * a.q_batch_mul);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j % S;
        mbar_wait(&kv_empty[s], ((j / S) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], 2 * TILE_BYTES);
        tma_load_4d(sKV + s * 2 * TILE_BYTES, &mapK, &kv_full[s], 0, h, j * 128, nb * a.kv_batch_mul);
        tma_load_4d(sKV + s * 2 * TILE_BYTES + TILE_BYTES, &mapV, &kv_full[s], 0, h, j * 128,
                    nb * a.kv_batch_mul);
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ------------------------------------------------------------------ MMA issuer of tile x
    // warp-uniform issue loop, tcgen05 instructions behind elect.sync (see the v8 kernel below for the finding)
    const int x = (warp == 1) ? 0 : 1;
    if (x < nq) {
      const uint32_t idesc_qk = make_idesc_f16(128, BLK, 0, 0);
      const uint32_t idesc_pv = make_idesc_f16(128, D, 0, 1);
      const uint32_t qa = smem_u32(sQ) + x * TILE_BYTES, skv = smem_u32(sKV);
      const uint32_t tX = tmem + x * TM_STRIDE;
      auto issue_qk = [&](int i) {                 // S[i % NSBUF] = Q K(block i)^T
        const uint32_t ka = skv + ((i / BPT) % S) * 2 * TILE_BYTES + (i % BPT) * BLK * ROWB;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < D / 16; ++k)
            mma_ss(tX + (i % NSBUF) * BLK, make_smem_desc(qa + k * 32, 16, SBO, SWZ),
                   make_smem_desc(ka + k * 32, 16, SBO, SWZ), idesc_qk, k != 0);
          tc_commit(&s_full[x][i % NSBUF]);
        }
        __syncwarp();
      };
      auto issue_pv = [&](int i) {                 // O' = P(block i) V(block i)
        const uint32_t va = skv + ((i / BPT) % S) * 2 * TILE_BYTES + TILE_BYTES + (i % BPT) * BLK * ROWB;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BLK / 16; ++k)
            mma_ts(tX + TM_O, tX + TM_P + k * 8, make_smem_desc(va + k * 16 * ROWB, SBO, SBO, SWZ), idesc_pv, k != 0);
          tc_commit(&o_full[x]);
          if ((i % BPT) == BPT - 1 || i + 1 == n_blk) tc_commit(&kv_empty[(i / BPT) % S]);   // tile fully consumed
        }
        __syncwarp();
      };
      const bool tr = a.trace && lane == 0 && x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
      mbar_wait(&q_full, 0);
      int tiles_waited = 0;                        // K/V tiles whose arrival this thread has observed
      auto need_tile = [&](int t) {
        while (tiles_waited <= t) {
          mbar_wait(&kv_full[tiles_waited % S], (tiles_waited / S) & 1);
          ++tiles_waited;
        }
        tc_fence_after();
      };
      for (int i = 0; i < NSBUF && i < n_blk; ++i) {
        need_tile(i / BPT);
        issue_qk(i);
      }
      for (int i = 0; i < n_blk; ++i) {
        if (i + NSBUF < n_blk) {
          // next scores for this buffer as soon as the softmax warps have pulled S(i) into registers
          need_tile((i + NSBUF) / BPT);
          mbar_wait(&s_free[x][i % NSBUF], (i / NSBUF) & 1);
          tc_fence_after();
          issue_qk(i + NSBUF);
          if (tr && i < 16) a.trace[i * 16 + 10] = clock64();
        }
        mbar_wait(&p_full[x], i & 1);              // P(i) is in TMEM
        tc_fence_after();
        if (tr && i < 16) a.trace[i * 16 + 8] = clock64();
        issue_pv(i);
        if (tr && i < 16) a.trace[i * 16 + 9] = clock64();
      }
    }
  }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    const int x = (warp - 4) >> 2;                 // 0: tile A (warps 4-7), 1: tile B (warps 8-11)
    if (x < nq) {
      const int quarter = warp & 3;
      const int row = quarter * 32 + lane;         // row inside the tile == TMEM lane
      const uint32_t tX = tmem + x * TM_STRIDE + ((uint32_t)(quarter * 32) << 16);
      const float c = a.scale_log2e;
      float O[D];
#pragma unroll
      for (int i = 0; i < D; ++i) O[i] = 0.f;
      float m = -INFINITY, l = 0.f;
      // The two softmax warpgroups share each SM sub-partition's MUFU unit.  Started together they stay in
      // lock-step: both in the exponential phase (MUFU oversubscribed), then both in the load / max / fold
      // phase (MUFU idle).  Delaying tile B by about half a block period interleaves the phases.
      if (x == 1 && a.stagger > 0) {
        const long long t0 = clock64();
        while (clock64() - t0 < a.stagger) {}
      }
      const bool pingpong = (a.dbg & 0x20) && nq == 2;
      if (pingpong && x == 1) asm volatile("bar.arrive 1, 256;" ::: "memory");   // tile A goes first

      auto fold_o = [&](float alpha) {             // O <- (O + O'(i-1)) * alpha
#pragma unroll
        for (int d0 = 0; d0 < D; d0 += 32) {
          uint32_t r[32];
          tmem_ld_x32(tX + TM_O + d0, r);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; ++k) O[d0 + k] = (O[d0 + k] + __uint_as_float(r[k])) * alpha;
        }
      };

      const bool tr = a.trace && warp == 4 && lane == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
      for (int i = 0; i < n_blk; ++i) {
        const int valid = a.Lk - i * BLK;          // columns >= valid are padding (last block)
        const int b = i % NSBUF;
        if (tr && i < 32) a.trace[(i & 15) * 16 + (i < 16 ? 0 : 12)] = clock64();
        if (a.trace && warp == 8 && lane == 0 && cta_lin == 0 && i < 16) a.trace[i * 16 + 14] = clock64();
        mbar_wait(&s_full[x][b], (i / NSBUF) & 1);
        tc_fence_after();
        if (tr && i < 16) a.trace[i * 16 + 1] = clock64();
        uint32_t sv[BLK];
#pragma unroll
        for (int c0 = 0; c0 < BLK; c0 += 32) tmem_ld_x32(tX + b * BLK + c0, *reinterpret_cast<uint32_t(*)[32]>(&sv[c0]));
        tmem_ld_wait();
        if (tr && i < 16) a.trace[i * 16 + 2] = clock64();
        tc_fence_before();
        __syncwarp();                              // tcgen05.wait::ld is warp-wide: every lane's scores are in registers
        if (lane == 0) mbar_arrive(&s_free[x][b]); // the MMA warp may overwrite this score buffer
        if (valid < BLK) {
#pragma unroll
          for (int k = 0; k < BLK; ++k)
            if (k >= valid) sv[k] = 0xff800000u;     // -inf
        }
        // row max with 8 independent chains (a single fmax chain costs 64 x 6 cycles of pure latency)
        float m8[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) m8[k] = __uint_as_float(sv[k]);
#pragma unroll
        for (int k = 8; k < BLK; ++k) m8[k & 7] = fmaxf(m8[k & 7], __uint_as_float(sv[k]));
        const float mx = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])),
                               fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
        const float m_new = fmaxf(m, mx);
        const float alpha = fast_exp2((m - m_new) * c);
        m = m_new;
        const float mc = m_new * c;
        float ls[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // four packed (even, odd column) partial sums
        float nmc = -mc;
        auto exp_pair = [&](int k) {               // columns k, k+1 -> fp16 pair in sv[k / 2]
          float x0, x1;
          ffma2(x0, x1, __uint_as_float(sv[k]), __uint_as_float(sv[k + 1]), c, c, nmc, nmc);
          float p0, p1;
          if (POLY > 0 && ((k >> 1) % POLY) == POLY - 1) {
            // 2^x = 2^n * 2^f, n = round(x), f = x - n in [-0.5, 0.5]; n sits in the low mantissa bits of
            // t = x + 1.5 * 2^23 and is added straight into the exponent field of the polynomial's value
            x0 = fmaxf(x0, -120.f); x1 = fmaxf(x1, -120.f);
            float t0, t1, n0, n1, f0, f1;
            fadd2(t0, t1, x0, x1, 12582912.f, 12582912.f);
            fadd2(n0, n1, t0, t1, -12582912.f, -12582912.f);
            fadd2(f0, f1, x0, x1, -n0, -n1);
            ffma2(p0, p1, f0, f1, 0.05517164617776871f, 0.05517164617776871f, 0.2426111251115799f, 0.2426111251115799f);
            ffma2(p0, p1, p0, p1, f0, f1, 0.6932609677314758f, 0.6932609677314758f);
            ffma2(p0, p1, p0, p1, f0, f1, 0.9999280571937561f, 0.9999280571937561f);
            p0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
            p1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
          } else {
            p0 = fast_exp2(x0); p1 = fast_exp2(x1);
          }
          const int j = (k >> 1) & 3;
          fadd2(ls[2 * j], ls[2 * j + 1], ls[2 * j], ls[2 * j + 1], p0, p1);
          const __half2 hp = __floats2half2_rn(p0, p1);
          sv[k >> 1] = *reinterpret_cast<const uint32_t*>(&hp);
        };
        uint32_t ro[D];
        if (pingpong) {
          // MUFU ping-pong: the two warpgroups take turns in the exponential phase (named barriers 1 / 2),
          // so one runs its loads / max / fold / stores while the other owns the MUFU unit
          if (i > 0) {
            mbar_wait(&o_full[x], (i - 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int d0 = 0; d0 < D; d0 += 32) tmem_ld_x32(tX + TM_O + d0, *reinterpret_cast<uint32_t(*)[32]>(&ro[d0]));
          }
          asm volatile("bar.sync %1, 256;" : "+f"(nmc) : "r"(1 + x) : "memory");
          if (tr && i < 16) a.trace[i * 16 + 3] = clock64();
#pragma unroll
          for (int k = 0; k < BLK; k += 2) exp_pair(k);
          float lsum = ((ls[0] + ls[1]) + (ls[2] + ls[3])) + ((ls[4] + ls[5]) + (ls[6] + ls[7]));
          if (i + 1 < n_blk || x == 0) asm volatile("bar.arrive %1, 256;" : "+f"(lsum) : "r"(2 - x) : "memory");
          if (tr && i < 16) a.trace[i * 16 + 4] = clock64();
          l = l * alpha + lsum;
        } else {
        // first half of the exponentials; packed fp16 pairs overwrite sv[0..BLK/2) in place
#pragma unroll
        for (int k = 0; k < BLK / 2; k += 2) exp_pair(k);
        if (tr && i < 16) a.trace[i * 16 + 3] = clock64();
        // PV(i-1) has had half a block to finish: start pulling its product O' while the second half
        // of the exponentials is computed
        if (i > 0) {
          mbar_wait(&o_full[x], (i - 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int d0 = 0; d0 < D; d0 += 32) tmem_ld_x32(tX + TM_O + d0, *reinterpret_cast<uint32_t(*)[32]>(&ro[d0]));
        }
        if (tr && i < 16) a.trace[i * 16 + 4] = clock64();
#pragma unroll
        for (int k = BLK / 2; k < BLK; k += 2) exp_pair(k);
        l = l * alpha + (((ls[0] + ls[1]) + (ls[2] + ls[3])) + ((ls[4] + ls[5]) + (ls[6] + ls[7])));
        }
        if (i > 0) {                               // O <- (O + O'(i-1)) * alpha ; the P region is free again
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < D; k += 2) {
            fadd2(O[k], O[k + 1], O[k], O[k + 1], __uint_as_float(ro[k]), __uint_as_float(ro[k + 1]));
            fmul2(O[k], O[k + 1], O[k], O[k + 1], alpha, alpha);
          }
        }
        if (tr && i < 16) a.trace[i * 16 + 5] = clock64();
#pragma unroll
        for (int c0 = 0; c0 < BLK / 2; c0 += 16) tmem_st_x16(tX + TM_P + c0, *reinterpret_cast<uint32_t(*)[16]>(&sv[c0]));
        tmem_st_wait();
        if (tr && i < 16) a.trace[i * 16 + 6] = clock64();
        tc_fence_before();
        __syncwarp();                              // tcgen05.wait::st is warp-wide
        if (lane == 0) mbar_arrive(&p_full[x]);
      }
      if (tr) a.trace[13] = clock64();
      // last partial product
      mbar_wait(&o_full[x], (n_blk - 1) & 1);
      tc_fence_after();
      fold_o(1.0f);
      const int qi = q0 + x * 128 + row;
      if (a.lse && qi < a.lse_ld)
        a.lse[((long long)nb * a.H + h) * a.lse_ld + qi] = qi < a.Lq ? fmaf(m, c, log2f(l)) : INFINITY;
      if (qi < a.Lq) {
        const float inv = 1.0f / l;
        __half* op = a.o + (long long)nb * a.o_stride_b + (long long)qi * a.o_stride_l + (long long)h * a.o_stride_h;
#pragma unroll
        for (int k = 0; k < D; k += 8) {
          __align__(16) __half hh[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) hh[t] = __float2half_rn(O[k + t] * inv);
          *reinterpret_cast<uint4*>(op + k) = *reinterpret_cast<uint4*>(hh);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 3) tmem_dealloc(tmem, 512);
  if (a.trace && threadIdx.x == 0 && cta_lin < 1024) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    a.trace[256 + 3 * cta_lin + 2] = (long long)t;
  }
}

// ---------------------------------------------------------------------------------------
// v6 (d = 32): FOUR query tiles and four softmax warpgroups per CTA.  Measurements behind it
// (tools/attn_experiments.py, tools/probe/mufu_probe3.cu): the MUFU pipe sustains one exp2 warp
// instruction per 8.2 clocks with two or more warps feeding it, but a softmax thread's block is a
// serial chain (barrier wait -> tcgen05.ld -> max -> exponentials -> P store -> arrive) and with only
// two softmax warps per SM sub-partition their stalls coincide more often than not (v4: 2850 clocks
// per 2 x 128 x 128 scores against 2112 of MUFU work; staggering or ping-ponging the two recovers a few
// per cent).  Four warps per sub-partition keep the pipe fed.  To fit four tiles in 512 TMEM columns
// and 104 registers per softmax thread: 64-key blocks, single S and P buffers per tile, O accumulated
// in TMEM by the tensor core itself, and lazy rescaling (the reference maximum only moves when a block
// exceeds it by more than 2^8 -- softmax is shift invariant -- so O is read-modified-written rarely).
//   20 warps: warp t < 4 = tcgen05.mma issuer of tile t (warp 0 also drives the K/V TMA ring, warp 1
//   owns the TMEM allocation), warps 4-19 = softmax warpgroups (thread = query row = TMEM lane).
//   TMEM columns per tile: S (64 fp32) | P (32 = 64 fp16 keys) | O (32).
template <int POLY, bool TRACE>
__global__ void __launch_bounds__(640, 1)
attn_fwd6_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                 const __grid_constant__ CUtensorMap mapV, const AttnArgs a) {
  constexpr int D = 32, BLK = 64, NT = 4;
  constexpr int ROWB = D * 2;
  constexpr int TILE_BYTES = 128 * ROWB;
  constexpr uint64_t SWZ = SWZ_64B;
  constexpr uint32_t SBO = 8 * ROWB;
  constexpr int S = kAttnStages;
  constexpr uint32_t TM_P = 64, TM_O = 96, TM_STRIDE = 128;

  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t q_full, kv_full[S], kv_empty[S], s_full[NT], s_free[NT], p_full[NT], o_full[NT];
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                              // NT tiles
  uint8_t* sKV = smem + NT * TILE_BYTES;           // S x (K tile, V tile)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // work item -> (unit, key range); units keep the (query block fastest, head, batch) order of the old 3-D grid
  const int item = blockIdx.x;
  int unit = item, part = -1;
  if (item >= a.split_from) {
    unit = a.split_from + (item - a.split_from) / a.nsplit;
    part = (item - a.split_from) % a.nsplit;
  }
  const int qblk = unit % a.gx, h = (unit / a.gx) % a.H, nb = unit / (a.gx * a.H);
  const int cta_lin = item;
  if (TRACE && a.trace && threadIdx.x == 0 && cta_lin < 1024) {
    unsigned long long t; unsigned sm;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    asm volatile("mov.u32 %0, %smid;" : "=r"(sm));
    a.trace[256 + 3 * cta_lin] = sm; a.trace[256 + 3 * cta_lin + 1] = (long long)t;
  }
  const int q0 = qblk * (128 * NT);
  const int nq = min(NT, (a.Lq - q0 + 127) / 128);
  const int n_kv_all = (a.Lk + 127) / 128;
  const int t0 = part < 0 ? 0 : (part * n_kv_all) / a.nsplit;              // first / past-the-last 128-key TMA tile
  const int t1 = part < 0 ? n_kv_all : ((part + 1) * n_kv_all) / a.nsplit;
  const int n_kv = t1 - t0;
  const int b0 = 2 * t0;                                                    // first 64-key softmax block (global)
  const int n_blk = min((a.Lk + BLK - 1) / BLK, 2 * t1) - b0;               // blocks of this CTA (local index i)

  if (threadIdx.x == 0) {
    mbar_init(&q_full, 1);
    for (int i = 0; i < S; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], nq); }
    for (int x = 0; x < NT; ++x) {
      mbar_init(&s_full[x], 1);
      mbar_init(&s_free[x], 128);
      mbar_init(&p_full[x], 128);
      mbar_init(&o_full[x], 1);
    }
    fence_barrier_init();
    // the Q tiles and the first K/V stages are requested right here, by the thread that has just created the
    // barriers: their L2 / HBM latency overlaps the TMEM allocation and the CTA-wide barrier below
    pdl_wait();
    mbar_arrive_expect_tx(&q_full, nq * TILE_BYTES);
    for (int t = 0; t < nq; ++t)
      tma_load_4d(sQ + t * TILE_BYTES, &mapQ, &q_full, 0, h, q0 + t * 128, nb * a.q_batch_mul);
    for (int j = 0; j < min(S, n_kv); ++j) {
      mbar_arrive_expect_tx(&kv_full[j], 2 * TILE_BYTES);
      tma_load_4d(sKV + j * 2 * TILE_BYTES, &mapK, &kv_full[j], 0, h, (t0 + j) * 128, nb * a.kv_batch_mul);
      tma_load_4d(sKV + j * 2 * TILE_BYTES + TILE_BYTES, &mapV, &kv_full[j], 0, h, (t0 + j) * 128, nb * a.kv_batch_mul);
    }
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem = tmem_base_s;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    const int x = warp;
    if (lane == 0 && x < nq) {
      const uint32_t idesc_qk = make_idesc_f16(128, BLK, 0, 0);
      const uint32_t idesc_pv = make_idesc_f16(128, D, 0, 1);
      const uint32_t qa = smem_u32(sQ) + x * TILE_BYTES, skv = smem_u32(sKV);
      const uint32_t tX = tmem + x * TM_STRIDE;
      // K/V ring, driven by tile 0's issuer between its MMA batches
      int loaded = min(S, n_kv);                   // the first stages were requested in the prologue
      auto load_kv = [&](int j) {
        const int s = j % S;
        mbar_arrive_expect_tx(&kv_full[s], 2 * TILE_BYTES);
        tma_load_4d(sKV + s * 2 * TILE_BYTES, &mapK, &kv_full[s], 0, h, (t0 + j) * 128, nb * a.kv_batch_mul);
        tma_load_4d(sKV + s * 2 * TILE_BYTES + TILE_BYTES, &mapV, &kv_full[s], 0, h, (t0 + j) * 128, nb * a.kv_batch_mul);
      };
      auto pump = [&]() {                          // issue every load whose stage is already free
        while (loaded < n_kv && mbar_test_wait(&kv_empty[loaded % S], ((loaded / S) & 1) ^ 1)) load_kv(loaded++);
      };
      auto issue_qk = [&](int i) {                 // S = Q K(block i)^T
        const uint32_t ka = skv + ((i >> 1) % S) * 2 * TILE_BYTES + (i & 1) * BLK * ROWB;
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          mma_ss(tX, make_smem_desc(qa + k * 32, 16, SBO, SWZ), make_smem_desc(ka + k * 32, 16, SBO, SWZ),
                 idesc_qk, k != 0);
      };
      auto issue_pv = [&](int i) {                 // O += P(block i) V(block i)
        const uint32_t va = skv + ((i >> 1) % S) * 2 * TILE_BYTES + TILE_BYTES + (i & 1) * BLK * ROWB;
#pragma unroll
        for (int k = 0; k < BLK / 16; ++k)
          mma_ts(tX + TM_O, tX + TM_P + k * 8, make_smem_desc(va + k * 16 * ROWB, SBO, SBO, SWZ), idesc_pv,
                 (i > 0 || k > 0) ? 1u : 0u);
      };
      int tiles_waited = 0;
      auto need_tile = [&](int t) {
        while (tiles_waited <= t) {
          if (x == 0)
            while (loaded <= tiles_waited) {       // this thread owes the load it is about to wait for
              mbar_wait(&kv_empty[loaded % S], ((loaded / S) & 1) ^ 1);
              load_kv(loaded++);
            }
          mbar_wait(&kv_full[tiles_waited % S], (tiles_waited / S) & 1);
          ++tiles_waited;
        }
        tc_fence_after();
      };
      mbar_wait(&q_full, 0);
      need_tile(0);
      issue_qk(0);
      tc_commit(&s_full[x]);
      for (int i = 0; i < n_blk; ++i) {
        if (i + 1 < n_blk) {
          need_tile((i + 1) >> 1);
          mbar_wait(&s_free[x], i & 1);            // the softmax warpgroup holds S(i) in registers
          tc_fence_after();
          issue_qk(i + 1);
          tc_commit(&s_full[x]);
        }
        if (x == 0) pump();
        mbar_wait(&p_full[x], i & 1);
        tc_fence_after();
        issue_pv(i);
        tc_commit(&o_full[x]);
        if ((i & 1) || i + 1 == n_blk) tc_commit(&kv_empty[(i >> 1) % S]);
        if (x == 0) pump();
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    const int x = (warp - 4) >> 2;
    if (x < nq) {
      const int quarter = warp & 3;
      const int row = quarter * 32 + lane;
      const uint32_t tX = tmem + x * TM_STRIDE + ((uint32_t)(quarter * 32) << 16);
      const float c = a.scale_log2e;
      float m = -INFINITY, l = 0.f;
      // barrier addresses live in registers: the generic->shared conversions were re-done every block
      const uint32_t b_sfull = smem_u32(&s_full[x]), b_sfree = smem_u32(&s_free[x]);
      const uint32_t b_pfull = smem_u32(&p_full[x]), b_ofull = smem_u32(&o_full[x]);
      const bool tr = TRACE && a.trace && lane == 0 && cta_lin == 0 && quarter == 0;
      if (TRACE && a.stagger > 0 && x > 0) {                // de-phase the warpgroups (see header)
        const long long t0 = clock64();
        while (clock64() - t0 < (long long)a.stagger * x) {}
      }
    for (int i = 0; i < n_blk; ++i) {
        if (tr && i < 16) a.trace[i * 16 + x] = clock64();
        uint32_t cur[BLK];
        mbar_wait_u32(b_sfull, i & 1);
        tc_fence_after();
        tmem_ld_x32(tX, *reinterpret_cast<uint32_t(*)[32]>(&cur[0]));
        tmem_ld_x32(tX + 32, *reinterpret_cast<uint32_t(*)[32]>(&cur[32]));
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive_u32(b_sfree);
        if (tr && i < 16 && x == 0) a.trace[i * 16 + 4] = clock64();
        const int valid = a.Lk - (b0 + i) * BLK;
        if (valid < BLK) {
#pragma unroll
          for (int k = 0; k < BLK; ++k)
            if (k >= valid) cur[k] = 0xff800000u;   // -inf
        }
        float m4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) m4[k] = fmaxf(__uint_as_float(cur[k]), __uint_as_float(cur[k + 4]));
#pragma unroll
        for (int k = 8; k < BLK; k += 8) {
#pragma unroll
          for (int t = 0; t < 4; ++t)
            m4[t] = fmaxf(m4[t], fmaxf(__uint_as_float(cur[k + t]), __uint_as_float(cur[k + t + 4])));
        }
        const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        bool o_ready = false;                      // has this thread already observed P V(i - 1)?
        if (__any_sync(0xffffffffu, (mx - m) * c > 8.0f)) {
          const float m_new = fmaxf(m, mx);
          const float alpha = fast_exp2((m - m_new) * c);      // first block: exp2(-inf) = 0
          m = m_new;
          l *= alpha;
          if (i > 0) {                             // O <- O * alpha in TMEM
            mbar_wait_u32(b_ofull, (i - 1) & 1);
            tc_fence_after();
            o_ready = true;
#pragma unroll
            for (int d0 = 0; d0 < D; d0 += 16) {
              uint32_t r[16];
              tmem_ld_x16(tX + TM_O + d0, r);
              tmem_ld_wait();
#pragma unroll
              for (int k = 0; k < 16; ++k) r[k] = __float_as_uint(__uint_as_float(r[k]) * alpha);
              tmem_st_x16(tX + TM_O + d0, r);
            }
          }
        }
        const float nmc = -m * c;
        float ls[8];
#pragma unroll
        for (int k = 0; k < BLK; k += 2) {
          float x0, x1, p0, p1;
          ffma2(x0, x1, __uint_as_float(cur[k]), __uint_as_float(cur[k + 1]), c, c, nmc, nmc);
          if (POLY > 0 && ((k >> 1) % POLY) == POLY - 1) {
            x0 = fmaxf(x0, -120.f); x1 = fmaxf(x1, -120.f);
            float t0, t1, n0, n1, f0, f1;
            fadd2(t0, t1, x0, x1, 12582912.f, 12582912.f);
            fadd2(n0, n1, t0, t1, -12582912.f, -12582912.f);
            fadd2(f0, f1, x0, x1, -n0, -n1);
            ffma2(p0, p1, f0, f1, 0.05517164617776871f, 0.05517164617776871f, 0.2426111251115799f, 0.2426111251115799f);
            ffma2(p0, p1, p0, p1, f0, f1, 0.6932609677314758f, 0.6932609677314758f);
            ffma2(p0, p1, p0, p1, f0, f1, 0.9999280571937561f, 0.9999280571937561f);
            p0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
            p1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
          } else {
            p0 = fast_exp2(x0); p1 = fast_exp2(x1);
          }
          const int j = (k >> 1) & 3;
          if (k < 8) { ls[2 * j] = p0; ls[2 * j + 1] = p1; }
          else fadd2(ls[2 * j], ls[2 * j + 1], ls[2 * j], ls[2 * j + 1], p0, p1);
          const __half2 hp = __floats2half2_rn(p0, p1);
          cur[k >> 1] = *reinterpret_cast<const uint32_t*>(&hp);
        }
        l += ((ls[0] + ls[1]) + (ls[2] + ls[3])) + ((ls[4] + ls[5]) + (ls[6] + ls[7]));
        if (tr && i < 16 && x == 0) a.trace[i * 16 + 5] = clock64();
        if (i > 0 && !o_ready) {                   // the P columns were last read by P V(i - 1)
          mbar_wait_u32(b_ofull, (i - 1) & 1);
          tc_fence_after();
        }
        tmem_st_x16(tX + TM_P, *reinterpret_cast<uint32_t(*)[16]>(&cur[0]));
        tmem_st_x16(tX + TM_P + 16, *reinterpret_cast<uint32_t(*)[16]>(&cur[16]));
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive_u32(b_pfull);
        if (tr && i < 16 && x == 0) a.trace[i * 16 + 6] = clock64();
      }
      if (tr && x == 0) a.trace[13] = clock64();
      mbar_wait_u32(b_ofull, (n_blk - 1) & 1);
      tc_fence_after();
      const int qi = q0 + x * 128 + row;
      if (part >= 0) {                             // partial result of one key range: unnormalised O, m, l
        const size_t pr = ((size_t)(unit - a.split_from) * a.nsplit + part) * (NT * 128) + x * 128 + row;
        float* wo = a.ws_o + pr * D;
#pragma unroll
        for (int d0 = 0; d0 < D; d0 += 16) {
          uint32_t r[16];
          tmem_ld_x16(tX + TM_O + d0, r);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 16; k += 4)
            *reinterpret_cast<uint4*>(wo + d0 + k) = make_uint4(r[k], r[k + 1], r[k + 2], r[k + 3]);
        }
        *reinterpret_cast<float2*>(a.ws_ml + pr * 2) = make_float2(m, l);
      } else {
      const float inv = 1.0f / l;
      __half* op = a.o + (long long)nb * a.o_stride_b + (long long)qi * a.o_stride_l + (long long)h * a.o_stride_h;
#pragma unroll
      for (int d0 = 0; d0 < D; d0 += 16) {
        uint32_t r[16];
        tmem_ld_x16(tX + TM_O + d0, r);
        tmem_ld_wait();
        if (qi < a.Lq) {
#pragma unroll
          for (int k = 0; k < 16; k += 8) {
            __align__(16) __half hh[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) hh[t] = __float2half_rn(__uint_as_float(r[k + t]) * inv);
            *reinterpret_cast<uint4*>(op + d0 + k) = *reinterpret_cast<uint4*>(hh);
          }
        }
      }
      }
    }
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
  if (TRACE && a.trace && threadIdx.x == 0 && cta_lin < 1024) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    a.trace[256 + 3 * cta_lin + 2] = (long long)t;
  }
}

// ---------------------------------------------------------------------------------------
// v8 (d = 32) = v6 with the MMA issue path made warp-uniform.  Finding behind it (tools/attn_trace.py, round 2):
// with ALL softmax arithmetic removed the v6 kernel still took 222 of its 255 us -- the single thread per tile
// that issues tcgen05.mma / commit / TMA was the bottleneck, not the MUFU pipe.  Inside `if (lane == 0)` every
// operand lives in per-thread registers, so ptxas wraps each tcgen05.mma, commit and TMA in an
// ELECT / R2UR.BROADCAST / BRA.U.ANY "waterfall" loop and recomputes descriptors with dependent integer chains:
// ~2 200 clocks per (QK, PV) pair of batches on a warp that shares its scheduler with four busy softmax warps.
// Here the issuer warps run their loop with all 32 lanes (every value provably warp-uniform: blockIdx, kernel
// arguments, a shuffled warp index), descriptors are built once and advanced by adding to their address field,
// only the tcgen05 / TMA instructions themselves sit behind elect.sync, and the softmax warps arrive on their barriers once per warp instead of once
// per thread (4 arrivals instead of 128 shared-memory atomics per hand-off).
//   20 warps: 0-3 MMA issuers (tile = warp; warp 0 also keeps the K/V ring going), 4-19 softmax (thread = query
//   row = TMEM lane).
template <int PMASK, int DROP, bool TRACE, int NT>
__global__ void __launch_bounds__(NT * 160, NT == 2 ? 2 : 1)
attn_fwd8_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                 const __grid_constant__ CUtensorMap mapV, const AttnArgs a) {
  constexpr int D = 32, BLK = 64;
  constexpr int ROWB = D * 2;
  constexpr int TILE_BYTES = 128 * ROWB;
  constexpr uint64_t SWZ = SWZ_64B;
  constexpr uint32_t SBO = 8 * ROWB;
  constexpr int S = kAttnStages;
  constexpr uint32_t TM_P = 64, TM_O = 96, TM_STRIDE = 128;

  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t q_full, kv_full[S], kv_empty[S], s_full[NT], s_free[NT], p_full[NT], o_full[NT];
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                              // NT tiles
  uint8_t* sKV = smem + NT * TILE_BYTES;           // S x (K tile, V tile)

  const int warp = uniform_warp_id();
  const int lane = threadIdx.x & 31;
  const int item = blockIdx.x;
  int unit = item, part = -1;
  if (item >= a.split_from) {
    unit = a.split_from + (item - a.split_from) / a.nsplit;
    part = (item - a.split_from) % a.nsplit;
  }
  const int qblk = unit % a.gx, h = (unit / a.gx) % a.H, nb = unit / (a.gx * a.H);
  const int q0 = qblk * (128 * NT);
  const int nq = min(NT, (a.Lq - q0 + 127) / 128);
  const int n_kv_all = (a.Lk + 127) / 128;
  const int t0 = part < 0 ? 0 : (part * n_kv_all) / a.nsplit;
  const int t1 = part < 0 ? n_kv_all : ((part + 1) * n_kv_all) / a.nsplit;
  const int n_kv = t1 - t0;
  const int b0 = 2 * t0;
  const int n_blk = min((a.Lk + BLK - 1) / BLK, 2 * t1) - b0;

  if (TRACE && a.trace && threadIdx.x == 0 && item == 0) a.trace[240] = clock64();
  if (threadIdx.x == 0) {
    mbar_init(&q_full, 1);
    for (int i = 0; i < S; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], nq); }
    for (int x = 0; x < NT; ++x) {
      mbar_init(&s_full[x], 1);
      mbar_init(&s_free[x], 4);                    // one arrival per softmax warp of the tile
      mbar_init(&p_full[x], 4);
      mbar_init(&o_full[x], 1);
    }
    fence_barrier_init();
    pdl_wait();
    mbar_arrive_expect_tx(&q_full, nq * TILE_BYTES);
    for (int t = 0; t < nq; ++t)
      tma_load_4d(sQ + t * TILE_BYTES, &mapQ, &q_full, 0, h, q0 + t * 128, nb * a.q_batch_mul);
    for (int j = 0; j < min(S, n_kv); ++j) {
      mbar_arrive_expect_tx(&kv_full[j], 2 * TILE_BYTES);
      tma_load_4d(sKV + j * 2 * TILE_BYTES, &mapK, &kv_full[j], 0, h, (t0 + j) * 128, nb * a.kv_batch_mul);
      tma_load_4d(sKV + j * 2 * TILE_BYTES + TILE_BYTES, &mapV, &kv_full[j], 0, h, (t0 + j) * 128, nb * a.kv_batch_mul);
    }
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_s, NT * 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem = tmem_base_s;
  if (TRACE && a.trace && threadIdx.x == 0 && item == 0) a.trace[241] = clock64();

  if (warp < NT) {
    // ------------------------------------------------------------------ MMA issuer of tile x (all lanes, uniform)
    // (a 21st warp for the K/V ring does not fit: the register file is per SM sub-partition, and one issuer
    // at 32 plus four softmax warps at 112 registers already fill its 16 K entries -- tile 0's issuer keeps
    // the ring going between its MMA batches, with warp-uniform polling)
    // NT = 2 (two CTAs per SM): no register re-distribution -- a CTA's ten warps put three warps on two of the SM
    // sub-partitions and two on the others, and the per-sub-partition register pool of the pair without an
    // issuer warp has nothing to hand to its softmax warps (setmaxnreg.inc never returns); 96 registers for all.
    if constexpr (NT == 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    const int x = warp;
    if (x < nq) {
      const uint32_t idesc_qk = make_idesc_f16(128, BLK, 0, 0);
      const uint32_t idesc_pv = make_idesc_f16(128, D, 0, 1);
      const uint32_t tX = tmem + x * TM_STRIDE;
      // descriptors: high words are constants, low words = (address >> 4) | (LBO >> 4) << 16; a block further
      // on in shared memory is reached by adding (bytes >> 4) to the low word (addresses stay below 2^18)
      const uint32_t hi = (uint32_t)((make_smem_desc(0, 0, SBO, SWZ)) >> 32);
      const uint32_t q_lo = (uint32_t)make_smem_desc(smem_u32(sQ) + x * TILE_BYTES, 16, SBO, SWZ);
      const uint32_t k_lo0 = (uint32_t)make_smem_desc(smem_u32(sKV), 16, SBO, SWZ);                  // stage 0, first half
      const uint32_t v_lo0 = (uint32_t)make_smem_desc(smem_u32(sKV) + TILE_BYTES, SBO, SBO, SWZ);
      constexpr uint32_t STAGE16 = (2 * TILE_BYTES) >> 4, HALF16 = (BLK * ROWB) >> 4;
      auto desc = [&](uint32_t lo) { return ((uint64_t)hi << 32) | lo; };
      uint32_t qk_off = 0, pv_off = 0;             // (stage, half) offsets of the next QK / PV block, in 16 B units
      int qk_stage = 0, pv_stage = 0, qk_half = 0, pv_half = 0;
      uint32_t kv_par = 0;                         // parity of kv_full[qk_stage] for the tile about to be used
      const uint32_t b_kvfull = smem_u32(&kv_full[0]), b_kvempty = smem_u32(&kv_empty[0]);
      const uint32_t b_sfull = smem_u32(&s_full[x]), b_sfree = smem_u32(&s_free[x]);
      const uint32_t b_pfull = smem_u32(&p_full[x]), b_ofull = smem_u32(&o_full[x]);
      auto issue_qk = [&]() {                      // S = Q K(next block)^T, then advance
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < D / 16; ++k)
            mma_ss(tX, desc(q_lo + 2 * k), desc(k_lo0 + qk_off + 2 * k), idesc_qk, k != 0);
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b_sfull) : "memory");
        }
        __syncwarp();
        qk_half ^= 1;
        qk_off += HALF16;
        if (qk_half == 0) {
          qk_off += STAGE16 - 2 * HALF16;
          if (++qk_stage == S) { qk_stage = 0; qk_off = 0; kv_par ^= 1; }
        }
      };
      // K/V ring (tile 0's issuer): tiles [0, min(S, n_kv)) were requested in the prologue
      int loaded = min(S, n_kv), ld_stage = 0;
      uint32_t ld_par = 0;                         // parity to wait for on kv_empty[ld_stage]: first re-use waits for phase 0
      auto load_next = [&]() {                     // caller has made sure stage ld_stage is free
        if (elect_one()) {
          const uint32_t dst = smem_u32(sKV) + ld_stage * 2 * TILE_BYTES;
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b_kvfull + 8 * ld_stage), "r"(2 * TILE_BYTES) : "memory");
          asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                       ::"r"(dst), "l"(&mapK), "r"(b_kvfull + 8 * ld_stage), "r"(0), "r"(h), "r"((t0 + loaded) * 128), "r"(nb * a.kv_batch_mul) : "memory");
          asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                       ::"r"(dst + TILE_BYTES), "l"(&mapV), "r"(b_kvfull + 8 * ld_stage), "r"(0), "r"(h), "r"((t0 + loaded) * 128), "r"(nb * a.kv_batch_mul) : "memory");
        }
        __syncwarp();
        ++loaded;
        if (++ld_stage == S) { ld_stage = 0; ld_par ^= 1; }
      };
      auto pump = [&]() {                          // request every tile whose stage is already free
        while (loaded < n_kv) {
          uint32_t ok;
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                       : "=r"(ok) : "r"(b_kvempty + 8 * ld_stage), "r"(ld_par) : "memory");
          if (!__shfl_sync(0xffffffffu, ok, 0)) break;
          load_next();
        }
      };
      int tiles_used = 1;                          // K/V tiles this issuer has started to consume
      auto need_tile = [&]() {                     // before the first block of K/V tile number `tiles_used`
        if (x == 0)
          while (loaded <= tiles_used) {           // this warp owes the load it is about to wait for
            mbar_wait_u32(b_kvempty + 8 * ld_stage, ld_par);
            load_next();
          }
        mbar_wait_u32(b_kvfull + 8 * qk_stage, kv_par);
        ++tiles_used;
      };
      mbar_wait(&q_full, 0);
      mbar_wait_u32(b_kvfull, 0);
      tc_fence_after();
      if (TRACE && a.trace && item == 0 && x == 0 && lane == 0) a.trace[242] = clock64();
      issue_qk();
      const bool tri = TRACE && a.trace && item == 0 && x == 0 && lane == 0;
      for (int i = 0; i < n_blk; ++i) {
        if (i + 1 < n_blk) {
          if (qk_half == 0) need_tile();           // first block of a new K/V tile
          if (tri && i < 16) a.trace[i * 16 + 9] = clock64();
          mbar_wait_u32(b_sfree, i & 1);           // the softmax warpgroup holds S(i) in registers
          tc_fence_after();
          if (tri && i < 16) a.trace[i * 16 + 7] = clock64();
          issue_qk();
        }
        if (x == 0) pump();
        mbar_wait_u32(b_pfull, i & 1);
        tc_fence_after();
        if (tri && i < 16) a.trace[i * 16 + 8] = clock64();
        const bool last_of_tile = pv_half == 1 || i + 1 == n_blk;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BLK / 16; ++k)
            mma_ts(tX + TM_O, tX + TM_P + k * 8, desc(v_lo0 + pv_off + k * ((16 * ROWB) >> 4)), idesc_pv,
                   (i > 0 || k > 0) ? 1u : 0u);
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b_ofull) : "memory");
          if (last_of_tile)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b_kvempty + 8 * pv_stage) : "memory");
        }
        __syncwarp();
        pv_half ^= 1;
        pv_off += HALF16;
        if (pv_half == 0) {
          pv_off += STAGE16 - 2 * HALF16;
          if (++pv_stage == S) { pv_stage = 0; pv_off = 0; }
        }
        if (x == 0) pump();
      }
    }
  } else {
    if constexpr (NT == 4) asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    const int x = (warp - NT) >> 2;               // NT = 2: warps 2-5 / 6-9 = tiles 0 / 1, TMEM lane quarter = warp & 3
    if (x < nq) {
      const int quarter = warp & 3;
      const int row = quarter * 32 + lane;
      const uint32_t tX = tmem + x * TM_STRIDE + ((uint32_t)(quarter * 32) << 16);
      const float c = a.scale_log2e;
      float m = -INFINITY, l = 0.f;
      const uint32_t b_sfull = smem_u32(&s_full[x]), b_sfree = smem_u32(&s_free[x]);
      const uint32_t b_pfull = smem_u32(&p_full[x]), b_ofull = smem_u32(&o_full[x]);
      const bool tr = TRACE && a.trace && lane == 0 && item == 0 && quarter == 0;
      for (int i = 0; i < n_blk; ++i) {
        uint32_t cur[BLK];
        if (tr && i < 16) a.trace[i * 16 + x] = clock64();
        mbar_wait_u32(b_sfull, i & 1);
        tc_fence_after();
        if (tr && i < 16 && x == 0) a.trace[i * 16 + 10] = clock64();
        tmem_ld_x32(tX, *reinterpret_cast<uint32_t(*)[32]>(&cur[0]));
        tmem_ld_x32(tX + 32, *reinterpret_cast<uint32_t(*)[32]>(&cur[32]));
        tmem_ld_wait();
        tc_fence_before();
        if (lane == 0) mbar_arrive_u32(b_sfree);   // tcgen05.wait::ld is warp-wide: every lane's scores are in registers
        if (tr && i < 16 && x == 0) a.trace[i * 16 + 4] = clock64();
        const int valid = a.Lk - (b0 + i) * BLK;
        if (valid < BLK) {
#pragma unroll
          for (int k = 0; k < BLK; ++k)
            if (k >= valid) cur[k] = 0xff800000u;   // -inf
        }
        float m4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) m4[k] = fmaxf(__uint_as_float(cur[k]), __uint_as_float(cur[k + 4]));
        if (!(DROP & 4)) {
#pragma unroll
        for (int k = 8; k < BLK; k += 8) {
#pragma unroll
          for (int t = 0; t < 4; ++t)
            m4[t] = fmaxf(m4[t], fmaxf(__uint_as_float(cur[k + t]), __uint_as_float(cur[k + t + 4])));
        }
        }
        const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        bool o_ready = false;                      // has this thread already observed P V(i - 1)?
        if (__any_sync(0xffffffffu, (mx - m) * c > 8.0f)) {
          const float m_new = fmaxf(m, mx);
          const float alpha = fast_exp2((m - m_new) * c);      // first block: exp2(-inf) = 0
          m = m_new;
          l *= alpha;
          if (i > 0) {                             // O <- O * alpha in TMEM
            mbar_wait_u32(b_ofull, (i - 1) & 1);
            tc_fence_after();
            o_ready = true;
#pragma unroll
            for (int d0 = 0; d0 < D; d0 += 16) {
              uint32_t r[16];
              tmem_ld_x16(tX + TM_O + d0, r);
              tmem_ld_wait();
#pragma unroll
              for (int k = 0; k < 16; ++k) r[k] = __float_as_uint(__uint_as_float(r[k]) * alpha);
              tmem_st_x16(tX + TM_O + d0, r);
            }
          }
        }
        const float nmc = -m * c;
        if (tr && i < 16 && x == 0) a.trace[i * 16 + 12] = clock64();
        float ls[8];
#pragma unroll
        for (int k = 0; k < BLK; k += 2) {
          float x0, x1, p0, p1;
          if (DROP & 1) { x0 = __uint_as_float(cur[k]); x1 = __uint_as_float(cur[k + 1]); }
          else ffma2(x0, x1, __uint_as_float(cur[k]), __uint_as_float(cur[k + 1]), c, c, nmc, nmc);
          if (DROP & 16) { p0 = x0; p1 = x1; } else
          if ((PMASK >> ((k >> 1) & 15)) & 1) {
            x0 = fmaxf(x0, -120.f); x1 = fmaxf(x1, -120.f);
            float t0_, t1_, n0, n1, f0, f1;
            fadd2(t0_, t1_, x0, x1, 12582912.f, 12582912.f);
            fadd2(n0, n1, t0_, t1_, -12582912.f, -12582912.f);
            fadd2(f0, f1, x0, x1, -n0, -n1);
            ffma2(p0, p1, f0, f1, 0.05517164617776871f, 0.05517164617776871f, 0.2426111251115799f, 0.2426111251115799f);
            ffma2(p0, p1, p0, p1, f0, f1, 0.6932609677314758f, 0.6932609677314758f);
            ffma2(p0, p1, p0, p1, f0, f1, 0.9999280571937561f, 0.9999280571937561f);
            p0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0_) << 23));
            p1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1_) << 23));
          } else {
            p0 = fast_exp2(x0); p1 = fast_exp2(x1);
          }
          const int j = (k >> 1) & 3;
          if (k < 8) { ls[2 * j] = p0; ls[2 * j + 1] = p1; }
          else if (!(DROP & 2)) fadd2(ls[2 * j], ls[2 * j + 1], ls[2 * j], ls[2 * j + 1], p0, p1);
          if (DROP & 8) { cur[k >> 1] = __float_as_uint(p0) ^ __float_as_uint(p1); } else {
          const __half2 hp = __floats2half2_rn(p0, p1);
          cur[k >> 1] = *reinterpret_cast<const uint32_t*>(&hp);
          }
        }
        l += ((ls[0] + ls[1]) + (ls[2] + ls[3])) + ((ls[4] + ls[5]) + (ls[6] + ls[7]));
        if (tr && i < 16 && x == 0) a.trace[i * 16 + 5] = clock64();
        if (i > 0 && !o_ready) {                   // the P columns were last read by P V(i - 1)
          mbar_wait_u32(b_ofull, (i - 1) & 1);
          tc_fence_after();
        }
        if (tr && i < 16 && x == 0) a.trace[i * 16 + 11] = clock64();
        tmem_st_x16(tX + TM_P, *reinterpret_cast<uint32_t(*)[16]>(&cur[0]));
        tmem_st_x16(tX + TM_P + 16, *reinterpret_cast<uint32_t(*)[16]>(&cur[16]));
        tmem_st_wait();
        tc_fence_before();
        if (lane == 0) mbar_arrive_u32(b_pfull);   // tcgen05.wait::st is warp-wide
        if (tr && i < 16 && x == 0) a.trace[i * 16 + 6] = clock64();
      }
      if (tr && x == 0) a.trace[13] = clock64();
      mbar_wait_u32(b_ofull, (n_blk - 1) & 1);
      tc_fence_after();
      const int qi = q0 + x * 128 + row;
      if (part >= 0) {                             // partial result of one key range: unnormalised O, m, l
        const size_t pr = ((size_t)(unit - a.split_from) * a.nsplit + part) * (NT * 128) + x * 128 + row;
        float* wo = a.ws_o + pr * D;
#pragma unroll
        for (int d0 = 0; d0 < D; d0 += 16) {
          uint32_t r[16];
          tmem_ld_x16(tX + TM_O + d0, r);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 16; k += 4)
            *reinterpret_cast<uint4*>(wo + d0 + k) = make_uint4(r[k], r[k + 1], r[k + 2], r[k + 3]);
        }
        *reinterpret_cast<float2*>(a.ws_ml + pr * 2) = make_float2(m, l);
      } else {
        const float inv = 1.0f / l;
        __half* op = a.o + (long long)nb * a.o_stride_b + (long long)qi * a.o_stride_l + (long long)h * a.o_stride_h;
#pragma unroll
        for (int d0 = 0; d0 < D; d0 += 16) {
          uint32_t r[16];
          tmem_ld_x16(tX + TM_O + d0, r);
          tmem_ld_wait();
          if (qi < a.Lq) {
#pragma unroll
            for (int k = 0; k < 16; k += 8) {
              __align__(16) __half hh[8];
#pragma unroll
              for (int t = 0; t < 8; ++t) hh[t] = __float2half_rn(__uint_as_float(r[k + t]) * inv);
              *reinterpret_cast<uint4*>(op + d0 + k) = *reinterpret_cast<uint4*>(hh);
            }
          }
        }
      }
    }
  }
  if (TRACE && a.trace && threadIdx.x == 128 && item == 0) a.trace[243] = clock64();   // tile 0's first softmax thread: output stored
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, NT * 128);
  if (TRACE && a.trace && threadIdx.x == 0 && item == 0) a.trace[244] = clock64();
}

// ---------------------------------------------------------------------------------------
// Short-sequence attention on CUDA cores (L <= 32): the DiT temporal self-attention
// (reference model/dit.py:254-260: sequences of T = 24 frames per latent token, d = 32).
// 0.6 GFLOP per block -- latency-, not tensor-bound; a 128-row MMA tile would be 81 % padding.
// One warp per (batch, head); lane = query; K/V rows staged in shared memory.
template <int D>
__global__ void __launch_bounds__(256) attn_small_kernel(const __half* __restrict__ q, const __half* __restrict__ k,
                                                         const __half* __restrict__ v, __half* __restrict__ o,
                                                         long long nbh, int H, int L, long long sb, long long sl,
                                                         long long sh, long long osb, long long osl, long long osh,
                                                         float scale) {
  __shared__ __align__(16) __half sK[8][32][D];
  __shared__ __align__(16) __half sV[8][32][D];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long bh = (long long)blockIdx.x * 8 + w;
  if (bh >= nbh) return;
  const long long b = bh / H;
  const int h = (int)(bh - b * H);
  const __half* qb = q + b * sb + (long long)h * sh;
  const __half* kb = k + b * sb + (long long)h * sh;
  const __half* vb = v + b * sb + (long long)h * sh;
  float qv[D];
  if (lane < L) {
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
      const uint4 tq = *reinterpret_cast<const uint4*>(qb + (long long)lane * sl + i * 8);
      const __half* hq = reinterpret_cast<const __half*>(&tq);
#pragma unroll
      for (int t = 0; t < 8; ++t) qv[i * 8 + t] = __half2float(hq[t]);
      *reinterpret_cast<uint4*>(&sK[w][lane][i * 8]) = *reinterpret_cast<const uint4*>(kb + (long long)lane * sl + i * 8);
      *reinterpret_cast<uint4*>(&sV[w][lane][i * 8]) = *reinterpret_cast<const uint4*>(vb + (long long)lane * sl + i * 8);
    }
  }
  __syncwarp();
  if (lane >= L) return;
  float sc[32];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < 32; ++j) {                           // unrolled so sc[] stays in registers
    if (j >= L) break;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {                      // one LDS.128 (broadcast) per 8 MACs
      const uint4 kk = *reinterpret_cast<const uint4*>(&sK[w][j][i * 8]);
      const __half2* k2 = reinterpret_cast<const __half2*>(&kk);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 kf = __half22float2(k2[t]);
        acc = fmaf(qv[i * 8 + 2 * t], kf.x, acc);
        acc = fmaf(qv[i * 8 + 2 * t + 1], kf.y, acc);
      }
    }
    sc[j] = acc * scale;
    mx = fmaxf(mx, sc[j]);
  }
  float l = 0.f;
  float ov[D];
#pragma unroll
  for (int i = 0; i < D; ++i) ov[i] = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (j >= L) break;
    const float p = __expf(sc[j] - mx);
    l += p;
    const float ph = __half2float(__float2half_rn(p));
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
      const uint4 vv = *reinterpret_cast<const uint4*>(&sV[w][j][i * 8]);
      const __half2* v2 = reinterpret_cast<const __half2*>(&vv);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 vf = __half22float2(v2[t]);
        ov[i * 8 + 2 * t] = fmaf(ph, vf.x, ov[i * 8 + 2 * t]);
        ov[i * 8 + 2 * t + 1] = fmaf(ph, vf.y, ov[i * 8 + 2 * t + 1]);
      }
    }
  }
  const float inv = 1.0f / l;
  __half* ob = o + b * osb + (long long)lane * osl + (long long)h * osh;
#pragma unroll
  for (int i = 0; i < D; i += 8) {
    __align__(16) __half hh[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) hh[t] = __float2half_rn(ov[i + t] * inv);
    *reinterpret_cast<uint4*>(ob + i) = *reinterpret_cast<uint4*>(hh);
  }
}

// Same computation when the H heads of a token are contiguous (q_strides[2] == D): one CTA per batch
// entry stages the [L x H*D] rows of q, k, v with fully coalesced 16 B loads (a row is H*D*2 = 1 KB
// contiguous even in the strided temporal view), one warp per head, lane = query; the output goes
// back through shared memory so stores are row-contiguous as well.
template <int D>
__global__ void __launch_bounds__(512) attn_small_rows_kernel(const __half* __restrict__ q, const __half* __restrict__ k,
                                                              const __half* __restrict__ v, __half* __restrict__ o,
                                                              int H, int L, long long sb, long long sl,
                                                              long long osb, long long osl, float scale) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  const int HD = H * D;
  __half* sQ = reinterpret_cast<__half*>(sm_raw);
  __half* sK = sQ + (size_t)L * HD;
  __half* sV = sK + (size_t)L * HD;
  const long long nb = blockIdx.x;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int vec_per_row = HD / 8;
  for (int idx = tid; idx < 3 * L * vec_per_row; idx += nthr) {
    const int which = idx / (L * vec_per_row), rem = idx - which * L * vec_per_row;
    const int l = rem / vec_per_row, c = (rem - l * vec_per_row) * 8;
    const __half* src = (which == 0 ? q : which == 1 ? k : v) + nb * sb + (long long)l * sl + c;
    *reinterpret_cast<uint4*>((which == 0 ? sQ : which == 1 ? sK : sV) + (size_t)l * HD + c) =
        *reinterpret_cast<const uint4*>(src);
  }
  __syncthreads();
  const int h = tid >> 5, lane = tid & 31;
  if (h < H && lane < L) {
    float qv[D];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
      const uint4 t = *reinterpret_cast<const uint4*>(sQ + (size_t)lane * HD + h * D + i * 8);
      const __half2* h2 = reinterpret_cast<const __half2*>(&t);
#pragma unroll
      for (int u = 0; u < 4; ++u) { const float2 f = __half22float2(h2[u]); qv[i * 8 + 2 * u] = f.x; qv[i * 8 + 2 * u + 1] = f.y; }
    }
    float sc[32];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (j >= L) break;
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < D / 8; ++i) {
        const uint4 kk = *reinterpret_cast<const uint4*>(sK + (size_t)j * HD + h * D + i * 8);
        const __half2* k2 = reinterpret_cast<const __half2*>(&kk);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float2 kf = __half22float2(k2[u]);
          acc = fmaf(qv[i * 8 + 2 * u], kf.x, acc);
          acc = fmaf(qv[i * 8 + 2 * u + 1], kf.y, acc);
        }
      }
      sc[j] = acc * scale;
      mx = fmaxf(mx, sc[j]);
    }
    float l = 0.f, ov[D];
#pragma unroll
    for (int i = 0; i < D; ++i) ov[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (j >= L) break;
      const float p = __expf(sc[j] - mx);
      l += p;
      const float ph = __half2float(__float2half_rn(p));
#pragma unroll
      for (int i = 0; i < D / 8; ++i) {
        const uint4 vv = *reinterpret_cast<const uint4*>(sV + (size_t)j * HD + h * D + i * 8);
        const __half2* v2 = reinterpret_cast<const __half2*>(&vv);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float2 vf = __half22float2(v2[u]);
          ov[i * 8 + 2 * u] = fmaf(ph, vf.x, ov[i * 8 + 2 * u]);
          ov[i * 8 + 2 * u + 1] = fmaf(ph, vf.y, ov[i * 8 + 2 * u + 1]);
        }
      }
    }
    const float inv = 1.0f / l;
#pragma unroll
    for (int i = 0; i < D; i += 8) {                     // park the output row in this thread's own q slot
      __align__(16) __half hh[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) hh[u] = __float2half_rn(ov[i + u] * inv);
      *reinterpret_cast<uint4*>(sQ + (size_t)lane * HD + h * D + i) = *reinterpret_cast<uint4*>(hh);
    }
  }
  __syncthreads();
  for (int idx = tid; idx < L * vec_per_row; idx += nthr) {
    const int l = idx / vec_per_row, c = (idx - l * vec_per_row) * 8;
    *reinterpret_cast<uint4*>(o + nb * osb + (long long)l * osl + c) =
        *reinterpret_cast<const uint4*>(sQ + (size_t)l * HD + c);
  }
}

// ---------------------------------------------------------------------------------------
// Short-sequence attention, warp-level tensor-core variant (L <= 32, d = 32, heads contiguous): the DiT
// temporal self-attention at the benchmark shape.  The CUDA-core kernel above spends its time on
// half->float conversions and shared-memory reads per FMA (48 us for 50 MB of traffic); here one CTA
// stages the q/k/v rows of HC heads of one sequence with cp.async -- whole (row, HC heads) segments of
// HC x 64 B, so the strided (B,T,N,C) temporal view is read in 512 B pieces -- and one warp per head runs
// S = Q K^T and O = P V as mma.sync.m16n8k16 on fragments fetched with ldmatrix (a 128-row tcgen05 tile
// would be 81 % padding at L = 24; the warp-level MMA shape fits).  Rows are 64 B with a 16 B-chunk XOR
// swizzle (chunk ^ (row >> 1 & 3)) that makes every ldmatrix phase conflict-free.  Same rounding points
// as the kernel above: fp32 scores and statistics, P rounded to fp16 for P V, l summed unrounded.
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int HC>
__global__ void __launch_bounds__(HC * 32) attn_small_mma_kernel(const __half* __restrict__ q,
                                                                 const __half* __restrict__ k,
                                                                 const __half* __restrict__ v, __half* __restrict__ o,
                                                                 int L, long long sb, long long sl, long long osb,
                                                                 long long osl, float scale_log2e) {
  extern __shared__ uint4 sm4[];                    // [HC][3 (q,k,v)][32 rows][4 chunks of 16 B]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int hb = blockIdx.x * HC;
  const long long nb = blockIdx.y;
  const uint32_t sbase = smem_u32(sm4);
  pdl_wait();
  auto slot = [](int hs, int ts, int row, int chunk) { return ((hs * 3 + ts) * 32 + row) * 4 + (chunk ^ ((row >> 1) & 3)); };
  // ---- stage: (tensor, row) segments of HC x 64 B contiguous in global memory
  const int per_t = L * HC * 4;
  for (int idx = tid; idx < 3 * per_t; idx += HC * 32) {
    const int ts = idx / per_t, r0 = idx - ts * per_t;
    const int t = r0 / (HC * 4), r1 = r0 - t * (HC * 4);
    const int hs = r1 >> 2, c = r1 & 3;
    const __half* src = (ts == 0 ? q : ts == 1 ? k : v) + nb * sb + (long long)t * sl + (hb + hs) * 32 + c * 8;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + slot(hs, ts, t, c) * 16), "l"(src) : "memory");
  }
  for (int idx = tid; idx < 3 * (32 - L) * HC * 4; idx += HC * 32) {    // zero the padding rows (0 x garbage = NaN)
    const int c = idx & 3, hs = (idx >> 2) % HC, r = (idx >> 2) / HC;
    const int ts = r / (32 - L), t = L + r - ts * (32 - L);
    sm4[slot(hs, ts, t, c)] = make_uint4(0u, 0u, 0u, 0u);
  }
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  const int hs = warp, g = lane >> 2, tg = lane & 3;
  const int n_mt = (L + 15) >> 4;
  for (int mt = 0; mt < n_mt; ++mt) {
    uint32_t qa[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
      ldsm_x4(qa[ks], sbase + slot(hs, 0, 16 * mt + (lane & 15), 2 * ks + (lane >> 4)) * 16);
    float sc[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      uint32_t kb[4];
      ldsm_x4(kb, sbase + slot(hs, 1, 8 * nt + (lane & 7), lane >> 3) * 16);
#pragma unroll
      for (int e = 0; e < 4; ++e) sc[nt][e] = 0.f;
      mma_16816(sc[nt], qa[0], kb[0], kb[1]);
      mma_16816(sc[nt], qa[1], kb[2], kb[3]);
    }
    // rows g (e = 0,1) and g + 8 (e = 2,3) of this 16-row tile; columns 8 nt + 2 tg + (e & 1)
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = 8 * nt + 2 * tg + (e & 1);
        sc[nt][e] = (col < L) ? sc[nt][e] * scale_log2e : -INFINITY;
        mx[e >> 1] = fmaxf(mx[e >> 1], sc[nt][e]);
      }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float l[2] = {0.f, 0.f};
    uint32_t pa[2][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      float p[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        p[e] = fast_exp2(sc[nt][e] - mx[e >> 1]);
        l[e >> 1] += p[e];
      }
      const __half2 lo = __floats2half2_rn(p[0], p[1]), hi = __floats2half2_rn(p[2], p[3]);
      pa[nt >> 1][(nt & 1) * 2] = *reinterpret_cast<const uint32_t*>(&lo);        // a0 / a2: row g
      pa[nt >> 1][(nt & 1) * 2 + 1] = *reinterpret_cast<const uint32_t*>(&hi);    // a1 / a3: row g + 8
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
      l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
    }
    float oc[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) oc[nt][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        uint32_t vb[4];
        ldsm_x4_t(vb, sbase + slot(hs, 2, 16 * ks + (lane & 7) + 8 * ((lane >> 3) & 1), 2 * c2 + (lane >> 4)) * 16);
        mma_16816(oc[2 * c2], pa[ks], vb[0], vb[1]);
        mma_16816(oc[2 * c2 + 1], pa[ks], vb[2], vb[3]);
      }
    // park the output rows in this head's own q slot (its fragments are already in registers)
    __syncwarp();
    const float inv[2] = {1.0f / l[0], 1.0f / l[1]};
    __half* sq = reinterpret_cast<__half*>(sm4);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int row = 16 * mt + g + 8 * r;
        *reinterpret_cast<__half2*>(sq + slot(hs, 0, row, nt) * 8 + 2 * tg) =
            __floats2half2_rn(oc[nt][2 * r] * inv[r], oc[nt][2 * r + 1] * inv[r]);
      }
  }
  pdl_launch_dependents();
  __syncthreads();
  for (int idx = tid; idx < per_t; idx += HC * 32) {
    const int t = idx / (HC * 4), r1 = idx - t * (HC * 4);
    const int h2 = r1 >> 2, c = r1 & 3;
    *reinterpret_cast<uint4*>(o + nb * osb + (long long)t * osl + (hb + h2) * 32 + c * 8) = sm4[slot(h2, 0, t, c)];
  }
}

// Pipelined form of the kernel above (the DiT's temporal attention: 512 tokens x 16 heads, sequences of 24, 50 MB per
// launch): a CTA walks over several sequences with two shared-memory buffers, the cp.async loads of sequence i + 1 issued
// before sequence i is computed, the grid sized so that every CTA gets the same number of sequences.  Identical bits.
// Measured SLOWER than one sequence per CTA (see the launch site) and kept opt-in only.
template <int HC>
__global__ void __launch_bounds__(HC * 32, 2) attn_small_mma_pipe_kernel(const __half* __restrict__ q,
                                                                        const __half* __restrict__ k,
                                                                        const __half* __restrict__ v, __half* __restrict__ o,
                                                                        int L, int NB, long long sb, long long sl,
                                                                        long long osb, long long osl, float scale_log2e) {
  extern __shared__ uint4 sm4_all[];                // 2 x [HC][3 (q,k,v)][32 rows][4 chunks of 16 B]
  constexpr int BUF = HC * 3 * 32 * 4;              // uint4 per buffer
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int hb = blockIdx.x * HC;
  pdl_wait();
  auto slot = [](int hs, int ts, int row, int chunk) { return ((hs * 3 + ts) * 32 + row) * 4 + (chunk ^ ((row >> 1) & 3)); };
  const int per_t = L * HC * 4;
  auto stage = [&](long long nb, int b) {
    const uint32_t sbase = smem_u32(sm4_all + b * BUF);
    for (int idx = tid; idx < 3 * per_t; idx += HC * 32) {
      const int ts = idx / per_t, r0 = idx - ts * per_t;
      const int t = r0 / (HC * 4), r1 = r0 - t * (HC * 4);
      const int hs = r1 >> 2, c = r1 & 3;
      const __half* src = (ts == 0 ? q : ts == 1 ? k : v) + nb * sb + (long long)t * sl + (hb + hs) * 32 + c * 8;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + slot(hs, ts, t, c) * 16), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // padding rows of both buffers: zero once (cp.async only ever writes rows < L; parked outputs only touch q slots)
  for (int idx = tid; idx < 2 * 3 * (32 - L) * HC * 4; idx += HC * 32) {
    const int b = idx / (3 * (32 - L) * HC * 4), i2 = idx - b * (3 * (32 - L) * HC * 4);
    const int c = i2 & 3, hs = (i2 >> 2) % HC, r = (i2 >> 2) / HC;
    const int ts = r / (32 - L), t = L + r - ts * (32 - L);
    sm4_all[b * BUF + slot(hs, ts, t, c)] = make_uint4(0u, 0u, 0u, 0u);
  }
  long long nb = blockIdx.y;
  if (nb < NB) stage(nb, 0);
  const int hs = warp, g = lane >> 2, tg = lane & 3;
  const int n_mt = (L + 15) >> 4;
  for (int it = 0; nb < NB; nb += gridDim.y, ++it) {
    const int b = it & 1;
    uint4* sm4 = sm4_all + b * BUF;
    const uint32_t sbase = smem_u32(sm4);
    const long long nxt = nb + gridDim.y;
    if (nxt < NB) {
      stage(nxt, b ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    for (int mt = 0; mt < n_mt; ++mt) {
      uint32_t qa[2][4];
  #pragma unroll
      for (int ks = 0; ks < 2; ++ks)
        ldsm_x4(qa[ks], sbase + slot(hs, 0, 16 * mt + (lane & 15), 2 * ks + (lane >> 4)) * 16);
      float sc[4][4];
  #pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        uint32_t kb[4];
        ldsm_x4(kb, sbase + slot(hs, 1, 8 * nt + (lane & 7), lane >> 3) * 16);
  #pragma unroll
        for (int e = 0; e < 4; ++e) sc[nt][e] = 0.f;
        mma_16816(sc[nt], qa[0], kb[0], kb[1]);
        mma_16816(sc[nt], qa[1], kb[2], kb[3]);
      }
      // rows g (e = 0,1) and g + 8 (e = 2,3) of this 16-row tile; columns 8 nt + 2 tg + (e & 1)
      float mx[2] = {-INFINITY, -INFINITY};
  #pragma unroll
      for (int nt = 0; nt < 4; ++nt)
  #pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int col = 8 * nt + 2 * tg + (e & 1);
          sc[nt][e] = (col < L) ? sc[nt][e] * scale_log2e : -INFINITY;
          mx[e >> 1] = fmaxf(mx[e >> 1], sc[nt][e]);
        }
  #pragma unroll
      for (int r = 0; r < 2; ++r) {
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      }
      float l[2] = {0.f, 0.f};
      uint32_t pa[2][4];
  #pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        float p[4];
  #pragma unroll
        for (int e = 0; e < 4; ++e) {
          p[e] = fast_exp2(sc[nt][e] - mx[e >> 1]);
          l[e >> 1] += p[e];
        }
        const __half2 lo = __floats2half2_rn(p[0], p[1]), hi = __floats2half2_rn(p[2], p[3]);
        pa[nt >> 1][(nt & 1) * 2] = *reinterpret_cast<const uint32_t*>(&lo);        // a0 / a2: row g
        pa[nt >> 1][(nt & 1) * 2 + 1] = *reinterpret_cast<const uint32_t*>(&hi);    // a1 / a3: row g + 8
      }
  #pragma unroll
      for (int r = 0; r < 2; ++r) {
        l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
        l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
      }
      float oc[4][4];
  #pragma unroll
      for (int nt = 0; nt < 4; ++nt)
  #pragma unroll
        for (int e = 0; e < 4; ++e) oc[nt][e] = 0.f;
  #pragma unroll
      for (int ks = 0; ks < 2; ++ks)
  #pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          uint32_t vb[4];
          ldsm_x4_t(vb, sbase + slot(hs, 2, 16 * ks + (lane & 7) + 8 * ((lane >> 3) & 1), 2 * c2 + (lane >> 4)) * 16);
          mma_16816(oc[2 * c2], pa[ks], vb[0], vb[1]);
          mma_16816(oc[2 * c2 + 1], pa[ks], vb[2], vb[3]);
        }
      // park the output rows in this head's own q slot (its fragments are already in registers)
      __syncwarp();
      const float inv[2] = {1.0f / l[0], 1.0f / l[1]};
      __half* sq = reinterpret_cast<__half*>(sm4);
  #pragma unroll
      for (int nt = 0; nt < 4; ++nt)
  #pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int row = 16 * mt + g + 8 * r;
          *reinterpret_cast<__half2*>(sq + slot(hs, 0, row, nt) * 8 + 2 * tg) =
              __floats2half2_rn(oc[nt][2 * r] * inv[r], oc[nt][2 * r + 1] * inv[r]);
        }
    }
    __syncthreads();
    for (int idx = tid; idx < per_t; idx += HC * 32) {
      const int t = idx / (HC * 4), r1 = idx - t * (HC * 4);
      const int h2 = r1 >> 2, c = r1 & 3;
      *reinterpret_cast<uint4*>(o + nb * osb + (long long)t * osl + (hb + h2) * 32 + c * 8) = sm4[slot(h2, 0, t, c)];
    }
    __syncthreads();                                // the buffer is refilled by the prefetch of the next iteration
  }
  pdl_launch_dependents();
}

template <int D, int POLY>
static int launch_attn(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const AttnArgs& a,
                       int Nb, cudaStream_t st) {
  constexpr int SMEM = (2 + 2 * kAttnStages) * 128 * D * 2 + 1024;
  constexpr int BLK = (D == 32) ? 128 : 64, NSBUF = (D == 32) ? 1 : 2;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(attn_fwd_kernel<D, BLK, NSBUF, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             SMEM) != cudaSuccess)
      return GVF_ERR_CUDA;
    configured = true;
  }
  dim3 grid((a.Lq + 255) / 256, a.H, Nb);
  attn_fwd_kernel<D, BLK, NSBUF, POLY><<<grid, 384, SMEM, st>>>(mq, mk, mv, a);
  return cudaGetLastError() == cudaSuccess ? GVF_OK : GVF_ERR_CUDA;
}

// ---------------------------------------------------------------------------------------
// v7 = v6 made persistent: one CTA per SM walks the work list (item = blockIdx.x, + gridDim.x, ...).  A
// (batch, head) unit is only 8 .. 64 softmax blocks long, and every v6 CTA paid ~5 us of launch, barrier /
// TMEM set-up, first-load latency, pipeline fill and drain around them (a third of the 16 us a spatial
// self-attention CTA lasts).  Here the barriers and the TMEM allocation are made once, the K/V ring simply
// continues into the next item's tiles, the next item's Q tiles are fetched into the other half of a double
// buffer while the current item still runs, and a softmax warpgroup goes from its epilogue straight to block 0
// of the next item, whose scores are already waiting in TMEM.  All mbarrier parities come from running
// counters (K/V tiles / softmax blocks consumed so far) instead of the item-local indices of v6.
// Requires Lq % 512 == 0 (all four query tiles of every item live, the barrier arrival counts are fixed).
struct AttnItem {
  int unit, part, h, nb, q0, t0, n_kv, b0, n_blk;
};
__device__ __forceinline__ AttnItem attn_item(const AttnArgs& a, int item) {
  AttnItem it;
  it.unit = item; it.part = -1;
  if (item >= a.split_from) {
    it.unit = a.split_from + (item - a.split_from) / a.nsplit;
    it.part = (item - a.split_from) % a.nsplit;
  }
  const int qblk = it.unit % a.gx;
  it.h = (it.unit / a.gx) % a.H;
  it.nb = it.unit / (a.gx * a.H);
  it.q0 = qblk * 512;
  const int n_kv_all = (a.Lk + 127) / 128;
  it.t0 = it.part < 0 ? 0 : (it.part * n_kv_all) / a.nsplit;
  const int t1 = it.part < 0 ? n_kv_all : ((it.part + 1) * n_kv_all) / a.nsplit;
  it.n_kv = t1 - it.t0;
  it.b0 = 2 * it.t0;
  it.n_blk = min((a.Lk + 63) / 64, 2 * t1) - it.b0;
  return it;
}

template <int POLY>
__global__ void __launch_bounds__(640, 1)
attn_fwd7_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                 const __grid_constant__ CUtensorMap mapV, const AttnArgs a, const int n_items) {
  constexpr int D = 32, BLK = 64, NT = 4;
  constexpr int ROWB = D * 2;
  constexpr int TILE_BYTES = 128 * ROWB;
  constexpr uint64_t SWZ = SWZ_64B;
  constexpr uint32_t SBO = 8 * ROWB;
  constexpr int S = kAttnStages;
  constexpr uint32_t TM_P = 64, TM_O = 96, TM_STRIDE = 128;

  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t q_full[2], kv_full[S], kv_empty[S], s_full[NT], s_free[NT], p_full[NT], o_full[NT];
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                              // 2 buffers x NT tiles
  uint8_t* sKV = smem + 2 * NT * TILE_BYTES;       // S x (K tile, V tile)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int first = blockIdx.x, step = gridDim.x;

  auto load_q = [&](const AttnItem& it, int buf) {
    mbar_arrive_expect_tx(&q_full[buf], NT * TILE_BYTES);
    for (int t = 0; t < NT; ++t)
      tma_load_4d(sQ + (buf * NT + t) * TILE_BYTES, &mapQ, &q_full[buf], 0, it.h, it.q0 + t * 128, it.nb * a.q_batch_mul);
  };
  auto load_kv = [&](const AttnItem& it, int j, int g) {      // local tile j of the item = running tile g of the ring
    const int s = g % S;
    mbar_arrive_expect_tx(&kv_full[s], 2 * TILE_BYTES);
    tma_load_4d(sKV + s * 2 * TILE_BYTES, &mapK, &kv_full[s], 0, it.h, (it.t0 + j) * 128, it.nb * a.kv_batch_mul);
    tma_load_4d(sKV + s * 2 * TILE_BYTES + TILE_BYTES, &mapV, &kv_full[s], 0, it.h, (it.t0 + j) * 128, it.nb * a.kv_batch_mul);
  };

  int pre = 0;                                     // K/V tiles requested in the prologue
  if (threadIdx.x == 0) {
    mbar_init(&q_full[0], 1);
    mbar_init(&q_full[1], 1);
    for (int i = 0; i < S; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], NT); }
    for (int x = 0; x < NT; ++x) {
      mbar_init(&s_full[x], 1);
      mbar_init(&s_free[x], 128);
      mbar_init(&p_full[x], 128);
      mbar_init(&o_full[x], 1);
    }
    fence_barrier_init();
    pdl_wait();
    const AttnItem it0 = attn_item(a, first);
    load_q(it0, 0);
    pre = min(S, it0.n_kv);
    for (int j = 0; j < pre; ++j) load_kv(it0, j, j);
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem = tmem_base_s;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");   // the issuers carry the work-list state: 32 spilled it
    const int x = warp;
    if (lane == 0) {
      const uint32_t idesc_qk = make_idesc_f16(128, BLK, 0, 0);
      const uint32_t idesc_pv = make_idesc_f16(128, D, 0, 1);
      const uint32_t skv = smem_u32(sKV);
      const uint32_t tX = tmem + x * TM_STRIDE;
      // ---- producer state (tile 0's issuer): the next K/V tile to request, possibly of a later item
      int p_item = first, p_u = 0, p_j = pre, loaded = pre;          // loaded = running count of requested tiles
      AttnItem p_it = attn_item(a, first);
      auto p_advance = [&]() {                     // -> true while there is a tile left to request
        while (p_item < n_items && p_j >= p_it.n_kv) {
          p_item += step; ++p_u; p_j = 0;
          if (p_item < n_items) p_it = attn_item(a, p_item);
        }
        return p_item < n_items;
      };
      auto p_load_next = [&]() {                   // caller has made sure the stage is free
        if (p_j == 0 && p_u > 0) load_q(p_it, p_u & 1);   // first tile of a later item: its Q tiles travel with it
        load_kv(p_it, p_j, loaded);
        ++p_j; ++loaded;
      };
      auto pump = [&]() {                          // request every tile whose stage is already free
        while (p_advance() && mbar_test_wait(&kv_empty[loaded % S], ((loaded / S) & 1) ^ 1)) p_load_next();
      };
      int tiles_waited = 0;                        // running count of K/V tiles this issuer has waited for
      auto need_tile = [&](int g) {
        while (tiles_waited <= g) {
          if (x == 0)
            while (loaded <= tiles_waited) {       // this thread owes the load it is about to wait for
              p_advance();
              mbar_wait(&kv_empty[loaded % S], ((loaded / S) & 1) ^ 1);
              p_load_next();
            }
          mbar_wait(&kv_full[tiles_waited % S], (tiles_waited / S) & 1);
          ++tiles_waited;
        }
        tc_fence_after();
      };
      int kvbase = 0, gblk = 0, u = 0;             // running K/V tiles / softmax blocks / items before this one
      for (int item = first; item < n_items; item += step, ++u) {
        const AttnItem it = attn_item(a, item);
        const uint32_t qa = smem_u32(sQ) + ((u & 1) * NT + x) * TILE_BYTES;
        auto issue_qk = [&](int i) {               // S = Q K(block i)^T
          const uint32_t ka = skv + ((kvbase + (i >> 1)) % S) * 2 * TILE_BYTES + (i & 1) * BLK * ROWB;
#pragma unroll
          for (int k = 0; k < D / 16; ++k)
            mma_ss(tX, make_smem_desc(qa + k * 32, 16, SBO, SWZ), make_smem_desc(ka + k * 32, 16, SBO, SWZ),
                   idesc_qk, k != 0);
        };
        auto issue_pv = [&](int i) {               // O (+)= P(block i) V(block i)
          const uint32_t va = skv + ((kvbase + (i >> 1)) % S) * 2 * TILE_BYTES + TILE_BYTES + (i & 1) * BLK * ROWB;
#pragma unroll
          for (int k = 0; k < BLK / 16; ++k)
            mma_ts(tX + TM_O, tX + TM_P + k * 8, make_smem_desc(va + k * 16 * ROWB, SBO, SBO, SWZ), idesc_pv,
                   (i > 0 || k > 0) ? 1u : 0u);
        };
        mbar_wait(&q_full[u & 1], (u >> 1) & 1);
        need_tile(kvbase);
        if (gblk > 0) {                            // S still holds the last block of the previous item
          mbar_wait(&s_free[x], (gblk - 1) & 1);
          tc_fence_after();
        }
        issue_qk(0);
        tc_commit(&s_full[x]);
        for (int i = 0; i < it.n_blk; ++i) {
          const int gi = gblk + i;
          if (i + 1 < it.n_blk) {
            need_tile(kvbase + ((i + 1) >> 1));
            mbar_wait(&s_free[x], gi & 1);         // the softmax warpgroup holds S(i) in registers
            tc_fence_after();
            issue_qk(i + 1);
            tc_commit(&s_full[x]);
          }
          if (x == 0) pump();
          mbar_wait(&p_full[x], gi & 1);
          tc_fence_after();
          issue_pv(i);
          tc_commit(&o_full[x]);
          if ((i & 1) || i + 1 == it.n_blk) tc_commit(&kv_empty[(kvbase + (i >> 1)) % S]);
          if (x == 0) pump();
        }
        kvbase += it.n_kv;
        gblk += it.n_blk;
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int x = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t tX = tmem + x * TM_STRIDE + ((uint32_t)(quarter * 32) << 16);
    const float c = a.scale_log2e;
    const uint32_t b_sfull = smem_u32(&s_full[x]), b_sfree = smem_u32(&s_free[x]);
    const uint32_t b_pfull = smem_u32(&p_full[x]), b_ofull = smem_u32(&o_full[x]);
    int gblk = 0;
    for (int item = first; item < n_items; item += step) {
      const AttnItem it = attn_item(a, item);
      float m = -INFINITY, l = 0.f;
      for (int i = 0; i < it.n_blk; ++i) {
        const int gi = gblk + i;
        uint32_t cur[BLK];
        mbar_wait_u32(b_sfull, gi & 1);
        tc_fence_after();
        tmem_ld_x32(tX, *reinterpret_cast<uint32_t(*)[32]>(&cur[0]));
        tmem_ld_x32(tX + 32, *reinterpret_cast<uint32_t(*)[32]>(&cur[32]));
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive_u32(b_sfree);
        const int valid = a.Lk - (it.b0 + i) * BLK;
        if (valid < BLK) {
#pragma unroll
          for (int k = 0; k < BLK; ++k)
            if (k >= valid) cur[k] = 0xff800000u;   // -inf
        }
        float m4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) m4[k] = fmaxf(__uint_as_float(cur[k]), __uint_as_float(cur[k + 4]));
#pragma unroll
        for (int k = 8; k < BLK; k += 8) {
#pragma unroll
          for (int t = 0; t < 4; ++t)
            m4[t] = fmaxf(m4[t], fmaxf(__uint_as_float(cur[k + t]), __uint_as_float(cur[k + t + 4])));
        }
        const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        bool o_ready = false;
        if (__any_sync(0xffffffffu, (mx - m) * c > 8.0f)) {
          const float m_new = fmaxf(m, mx);
          const float alpha = fast_exp2((m - m_new) * c);      // first block: exp2(-inf) = 0
          m = m_new;
          l *= alpha;
          if (i > 0) {                             // O <- O * alpha in TMEM
            mbar_wait_u32(b_ofull, (gi - 1) & 1);
            tc_fence_after();
            o_ready = true;
#pragma unroll
            for (int d0 = 0; d0 < D; d0 += 16) {
              uint32_t r[16];
              tmem_ld_x16(tX + TM_O + d0, r);
              tmem_ld_wait();
#pragma unroll
              for (int k = 0; k < 16; ++k) r[k] = __float_as_uint(__uint_as_float(r[k]) * alpha);
              tmem_st_x16(tX + TM_O + d0, r);
            }
          }
        }
        const float nmc = -m * c;
        float ls[8];
#pragma unroll
        for (int k = 0; k < BLK; k += 2) {
          float x0, x1, p0, p1;
          ffma2(x0, x1, __uint_as_float(cur[k]), __uint_as_float(cur[k + 1]), c, c, nmc, nmc);
          if (POLY > 0 && ((k >> 1) % POLY) == POLY - 1) {
            x0 = fmaxf(x0, -120.f); x1 = fmaxf(x1, -120.f);
            float t0, t1, n0, n1, f0, f1;
            fadd2(t0, t1, x0, x1, 12582912.f, 12582912.f);
            fadd2(n0, n1, t0, t1, -12582912.f, -12582912.f);
            fadd2(f0, f1, x0, x1, -n0, -n1);
            ffma2(p0, p1, f0, f1, 0.05517164617776871f, 0.05517164617776871f, 0.2426111251115799f, 0.2426111251115799f);
            ffma2(p0, p1, p0, p1, f0, f1, 0.6932609677314758f, 0.6932609677314758f);
            ffma2(p0, p1, p0, p1, f0, f1, 0.9999280571937561f, 0.9999280571937561f);
            p0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
            p1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
          } else {
            p0 = fast_exp2(x0); p1 = fast_exp2(x1);
          }
          const int j = (k >> 1) & 3;
          if (k < 8) { ls[2 * j] = p0; ls[2 * j + 1] = p1; }
          else fadd2(ls[2 * j], ls[2 * j + 1], ls[2 * j], ls[2 * j + 1], p0, p1);
          const __half2 hp = __floats2half2_rn(p0, p1);
          cur[k >> 1] = *reinterpret_cast<const uint32_t*>(&hp);
        }
        l += ((ls[0] + ls[1]) + (ls[2] + ls[3])) + ((ls[4] + ls[5]) + (ls[6] + ls[7]));
        if (i > 0 && !o_ready) {                   // the P columns were last read by P V(i - 1)
          mbar_wait_u32(b_ofull, (gi - 1) & 1);
          tc_fence_after();
        }
        tmem_st_x16(tX + TM_P, *reinterpret_cast<uint32_t(*)[16]>(&cur[0]));
        tmem_st_x16(tX + TM_P + 16, *reinterpret_cast<uint32_t(*)[16]>(&cur[16]));
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive_u32(b_pfull);
      }
      gblk += it.n_blk;
      mbar_wait_u32(b_ofull, (gblk - 1) & 1);
      tc_fence_after();
      const int qi = it.q0 + x * 128 + row;
      if (it.part >= 0) {                          // partial result of one key range: unnormalised O, m, l
        const size_t pr = ((size_t)(it.unit - a.split_from) * a.nsplit + it.part) * (NT * 128) + x * 128 + row;
        float* wo = a.ws_o + pr * D;
#pragma unroll
        for (int d0 = 0; d0 < D; d0 += 16) {
          uint32_t r[16];
          tmem_ld_x16(tX + TM_O + d0, r);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 16; k += 4)
            *reinterpret_cast<uint4*>(wo + d0 + k) = make_uint4(r[k], r[k + 1], r[k + 2], r[k + 3]);
        }
        *reinterpret_cast<float2*>(a.ws_ml + pr * 2) = make_float2(m, l);
      } else {
        const float inv = 1.0f / l;
        __half* op = a.o + (long long)it.nb * a.o_stride_b + (long long)qi * a.o_stride_l + (long long)it.h * a.o_stride_h;
#pragma unroll
        for (int d0 = 0; d0 < D; d0 += 16) {
          uint32_t r[16];
          tmem_ld_x16(tX + TM_O + d0, r);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 16; k += 8) {
            __align__(16) __half hh[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) hh[t] = __float2half_rn(__uint_as_float(r[k + t]) * inv);
            *reinterpret_cast<uint4*>(op + d0 + k) = *reinterpret_cast<uint4*>(hh);
          }
        }
      }
      // The first P V of the next item overwrites O: it is issued only after this warpgroup's next p_full
      // arrival, which follows these reads in program order -- tcgen05.wait::ld above has retired them.
      tc_fence_before();
    }
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// Merge of the key-range partials written by attn_fwd6_kernel for the split units: with m the lazily updated
// reference maximum of each part, O = sum_p 2^((m_p - M) c) O_p, l likewise, M = max_p m_p.  One thread per
// query row.  The split exists because 384 (batch, head) units on 148 SMs are 2.6 waves: the 88 units of the
// third wave are cut into three key ranges each (264 items = two rounds of one third), which ends the wave
// at 2/3 of its length; the merge costs ~5 us.
__global__ void __launch_bounds__(256) attn_merge_kernel(const AttnArgs a, int n_split_units) {
  // eight threads per query row, four output columns each: every partial row is read as one 128 B segment
  const int idx = blockIdx.x * 256 + threadIdx.x;
  const int rr = idx >> 3, c4 = (idx & 7) * 4;
  if (rr >= n_split_units * 512) return;
  const int su = rr >> 9, r = rr & 511;
  const int unit = a.split_from + su;
  const int qblk = unit % a.gx, h = (unit / a.gx) % a.H, nb = unit / (a.gx * a.H);
  const int qi = qblk * 512 + r;
  if (qi >= a.Lq) return;
  float2 ml[4];
  float M = -INFINITY;
  for (int p = 0; p < a.nsplit; ++p) {
    ml[p] = *reinterpret_cast<const float2*>(a.ws_ml + (((size_t)su * a.nsplit + p) * 512 + r) * 2);
    M = fmaxf(M, ml[p].x);
  }
  float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f, l = 0.f;
  for (int p = 0; p < a.nsplit; ++p) {
    const size_t pr = ((size_t)su * a.nsplit + p) * 512 + r;
    const float w = fast_exp2((ml[p].x - M) * a.scale_log2e);
    l = fmaf(w, ml[p].y, l);
    const float4 v = *reinterpret_cast<const float4*>(a.ws_o + pr * 32 + c4);
    o0 = fmaf(w, v.x, o0); o1 = fmaf(w, v.y, o1); o2 = fmaf(w, v.z, o2); o3 = fmaf(w, v.w, o3);
  }
  const float inv = 1.0f / l;
  __half* op = a.o + (long long)nb * a.o_stride_b + (long long)qi * a.o_stride_l + (long long)h * a.o_stride_h + c4;
  const __half2 lo = __floats2half2_rn(o0 * inv, o1 * inv), hi = __floats2half2_rn(o2 * inv, o3 * inv);
  uint2 pk;
  pk.x = *reinterpret_cast<const uint32_t*>(&lo);
  pk.y = *reinterpret_cast<const uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(op) = pk;
}

// caller-owned scratch for the split units (gvf_attn_set_workspace); no allocation happens in here
static float* g_attn_ws = nullptr;
static size_t g_attn_ws_bytes = 0;

template <int POLY, bool TRACE, bool PERSIST = false, int V8 = 0, int DROP = 0, int NT = 4>
static int launch_attn6(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const AttnArgs& a,
                        int Nb, cudaStream_t st) {
  constexpr int SMEM = ((PERSIST ? 8 : NT) + 2 * kAttnStages) * 128 * 32 * 2 + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e;
    if constexpr (PERSIST) e = cudaFuncSetAttribute(attn_fwd7_kernel<POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    else if constexpr (V8 != 0) e = cudaFuncSetAttribute(attn_fwd8_kernel<V8, DROP, TRACE, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    else e = cudaFuncSetAttribute(attn_fwd6_kernel<POLY, TRACE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return GVF_ERR_CUDA;
    configured = true;
  }
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  AttnArgs b = a;
  b.gx = (a.Lq + 128 * NT - 1) / (128 * NT);
  const int units = b.gx * a.H * Nb;
  b.split_from = units;
  b.nsplit = 1;
  b.ws_o = nullptr;
  b.ws_ml = nullptr;
  // tail of the last wave: cut its units into three key ranges when that shortens it (see attn_merge_kernel)
  const int rem = units % num_sms, n_kv = (a.Lk + 127) / 128;
  constexpr int NSPLIT = 3;
  int n_split_units = 0;
  // measured: pays for the 4096-key static cross-attention (280 -> 261 us), not for 1370 keys (110 -> 120 us:
  // three prologues and the merge cost more than the shorter wave saves)
  if (!TRACE && NT == 4 && g_attn_ws && units > num_sms && rem > 0 && rem * NSPLIT <= 2 * num_sms && n_kv >= 16) {
    const size_t rows = (size_t)rem * NSPLIT * 512;
    if (rows * (32 + 2) * sizeof(float) <= g_attn_ws_bytes) {
      n_split_units = rem;
      b.split_from = units - rem;
      b.nsplit = NSPLIT;
      b.ws_o = g_attn_ws;
      b.ws_ml = g_attn_ws + rows * 32;
    }
  }
  const int n_items = b.split_from + n_split_units * b.nsplit;
  if constexpr (PERSIST) {
    const dim3 grid(n_items < num_sms ? n_items : num_sms);
    if (launch_pdl(attn_fwd7_kernel<POLY>, grid, dim3(640), SMEM, st, mq, mk, mv, b, n_items) != cudaSuccess) return GVF_ERR_CUDA;
  } else if constexpr (V8 != 0) {
    const dim3 grid(n_items);
    if (launch_pdl(attn_fwd8_kernel<V8, DROP, TRACE, NT>, grid, dim3(NT * 160), SMEM, st, mq, mk, mv, b) != cudaSuccess) return GVF_ERR_CUDA;
  } else {
    const dim3 grid(n_items);
    if (launch_pdl(attn_fwd6_kernel<POLY, TRACE>, grid, dim3(640), SMEM, st, mq, mk, mv, b) != cudaSuccess) return GVF_ERR_CUDA;
  }
  if (n_split_units) {
    static_assert(NSPLIT <= 4, "attn_merge_kernel keeps the (m, l) pairs of at most four parts in registers");
    attn_merge_kernel<<<(n_split_units * 512 * 8 + 255) / 256, 256, 0, st>>>(b, n_split_units);
    if (cudaGetLastError() != cudaSuccess) return GVF_ERR_CUDA;
  }
  return GVF_OK;
}

}  // namespace gvf

using namespace gvf;

static int g_attn_dbg = 0;
static long long* g_attn_trace = nullptr;
extern "C" GVF_API void gvf_attn_set_trace(void* p) { g_attn_trace = (long long*)p; }
extern "C" GVF_API void gvf_attn_set_debug(int v) { g_attn_dbg = v; }
extern "C" GVF_API void gvf_attn_set_workspace(void* ws, size_t bytes) {
  gvf::g_attn_ws = (float*)ws;
  gvf::g_attn_ws_bytes = ws ? bytes : 0;
}
static int g_small_rows = 0;   // row-staged temporal variant: measured slower than the warp-per-(batch,head) one

// q [Nb_q, Lq, H, D], k/v [Nb_kv, Lk, H, D] fp16 with element strides (batch, seq, head);
// innermost dim contiguous.  Nb_q / Nb_kv may be 1 (tensor shared by all Nb batches).
static int attn_fwd_impl(const void* q, const void* k, const void* v, void* o, int Nb, int Lq,
                         int Lk, int H, int D, const long long* q_strides,
                         const long long* k_strides, const long long* v_strides,
                         const long long* o_strides, int q_shared, int kv_shared,
                         float scale, float* lse, int lse_ld, void* stream) {
  if (!q || !k || !v || !o || !q_strides || !k_strides || !v_strides || !o_strides) return GVF_ERR_INVALID;
  if (lse && (lse_ld < ((Lq + 127) / 128) * 128 || (lse_ld % 128) || ((uintptr_t)lse & 15))) return GVF_ERR_INVALID;
  if (Nb <= 0 || Lq <= 0 || Lk <= 0 || H <= 0) return GVF_ERR_INVALID;
  if (D != 32 && D != 64) return GVF_ERR_UNSUPPORTED;
  for (int i = 0; i < 3; ++i)
    if ((q_strides[i] % 8) || (k_strides[i] % 8) || (v_strides[i] % 8) || (o_strides[i] % 8)) return GVF_ERR_INVALID;
  if (((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)o) & 15) return GVF_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;

  if (!lse && D == 32 && Lq <= 32 && Lk == Lq && !q_shared && !kv_shared && q_strides[0] == k_strides[0] &&
      q_strides[1] == k_strides[1] && q_strides[2] == k_strides[2] && q_strides[0] == v_strides[0] &&
      q_strides[1] == v_strides[1] && q_strides[2] == v_strides[2]) {
    if ((g_attn_dbg & 0xf0) != 0x40 && q_strides[2] == D && o_strides[2] == D && H % 8 == 0) {
      // heads contiguous: warp-level tensor-core variant (one CTA = 8 heads of one sequence)
      constexpr int HC = 8, SMEM = HC * 3 * 32 * 64;
      static bool configured_mma = false;
      if (!configured_mma) {
        if (cudaFuncSetAttribute(attn_small_mma_kernel<HC>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) !=
            cudaSuccess) return GVF_ERR_CUDA;
        configured_mma = true;
      }
      if ((g_attn_dbg & 0x10000) && Nb >= 64) {
        // pipelined persistent form (opt-in, gvf_attn_set_debug(0x10000)): two buffers per CTA, two CTAs per SM, the same
        // number of sequences for every CTA.  MEASURED (tools/temporal_bench.py, cold L2): 22.5 us against 20.5 us for the
        // one-sequence-per-CTA kernel below, 22.5 against 18.5 us inside the NFE -- four independent 49 KB CTAs per SM
        // already overlap each other's loads and compute better than two double-buffered 98 KB ones.
        static bool configured_pipe = false;
        static int num_sms = 0;
        if (!configured_pipe) {
          if (cudaFuncSetAttribute(attn_small_mma_pipe_kernel<HC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * SMEM) !=
              cudaSuccess) return GVF_ERR_CUDA;
          int dev = 0;
          cudaGetDevice(&dev);
          cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
          configured_pipe = true;
        }
        const int gx = H / HC;
        const int cap = (2 * num_sms) / gx > 0 ? (2 * num_sms) / gx : 1;
        const int per = (Nb + cap - 1) / cap;
        const int gy = (Nb + per - 1) / per;
        return launch_pdl(attn_small_mma_pipe_kernel<HC>, dim3(gx, gy), dim3(HC * 32), 2 * SMEM, st, (const __half*)q,
                          (const __half*)k, (const __half*)v, (__half*)o, Lq, Nb, (long long)q_strides[0],
                          (long long)q_strides[1], (long long)o_strides[0], (long long)o_strides[1],
                          scale * 1.4426950408889634f) == cudaSuccess ? GVF_OK : GVF_ERR_CUDA;
      }
      return launch_pdl(attn_small_mma_kernel<HC>, dim3(H / HC, Nb), dim3(HC * 32), SMEM, st, (const __half*)q,
                        (const __half*)k, (const __half*)v, (__half*)o, Lq, (long long)q_strides[0],
                        (long long)q_strides[1], (long long)o_strides[0], (long long)o_strides[1],
                        scale * 1.4426950408889634f) == cudaSuccess ? GVF_OK : GVF_ERR_CUDA;
    }
    if (g_small_rows && q_strides[2] == D && o_strides[2] == D && H <= 16) {
      // heads contiguous: row-staged variant (coalesced even for the strided temporal view)
      const int smem = 3 * Lq * H * D * 2;
      static bool configured = false;
      if (!configured) {
        if (cudaFuncSetAttribute(attn_small_rows_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 3 * 32 * 16 * 32 * 2) != cudaSuccess) return GVF_ERR_CUDA;
        configured = true;
      }
      attn_small_rows_kernel<32><<<Nb, 512, smem, st>>>((const __half*)q, (const __half*)k, (const __half*)v,
                                                       (__half*)o, H, Lq, q_strides[0], q_strides[1],
                                                       o_strides[0], o_strides[1], scale);
      return cudaGetLastError() == cudaSuccess ? GVF_OK : GVF_ERR_CUDA;
    }
    const long long nbh = (long long)Nb * H;
    const unsigned blocks = (unsigned)((nbh + 7) / 8);
    attn_small_kernel<32><<<blocks, 256, 0, st>>>((const __half*)q, (const __half*)k, (const __half*)v, (__half*)o,
                                                  nbh, H, Lq, q_strides[0], q_strides[1], q_strides[2],
                                                  o_strides[0], o_strides[1], o_strides[2], scale);
    return cudaGetLastError() == cudaSuccess ? GVF_OK : GVF_ERR_CUDA;
  }

  CUtensorMap mq, mk, mv;
  const CUtensorMapSwizzle sw = swizzle_for_bytes(D * 2);
  const uint32_t box[4] = {(uint32_t)D, 1, 128, 1};
  auto mk_map = [&](CUtensorMap* m, const void* p, int L, int nbt, const long long* s) {
    const uint64_t dims[4] = {(uint64_t)D, (uint64_t)H, (uint64_t)L, (uint64_t)nbt};
    // a shared tensor has a single batch entry; give it a harmless non-zero stride
    const uint64_t strides[4] = {1, (uint64_t)s[2], (uint64_t)s[1], (uint64_t)(nbt > 1 ? s[0] : (long long)L * s[1])};
    return make_tmap_f16(m, p, 4, dims, strides, box, sw);
  };
  if (!mk_map(&mq, q, Lq, q_shared ? 1 : Nb, q_strides)) return GVF_ERR_CUDA;
  if (!mk_map(&mk, k, Lk, kv_shared ? 1 : Nb, k_strides)) return GVF_ERR_CUDA;
  if (!mk_map(&mv, v, Lk, kv_shared ? 1 : Nb, v_strides)) return GVF_ERR_CUDA;
  AttnArgs a;
  a.o = (__half*)o;
  a.o_stride_b = o_strides[0]; a.o_stride_l = o_strides[1]; a.o_stride_h = o_strides[2];
  a.Lq = Lq; a.Lk = Lk; a.H = H;
  a.q_batch_mul = q_shared ? 0 : 1;
  a.kv_batch_mul = kv_shared ? 0 : 1;
  a.scale_log2e = scale * 1.4426950408889634f;
  a.dbg = g_attn_dbg & 0xff;
  a.stagger = ((g_attn_dbg >> 8) & 0xff) * 100;
  a.trace = g_attn_trace;
  if (lse) {                     // training forward: the generic kernel is the one that also leaves LSE2 behind
    a.lse = lse; a.lse_ld = lse_ld; a.dbg = 0; a.stagger = 0; a.trace = nullptr;
    return D == 64 ? launch_attn<64, 0>(mq, mk, mv, a, Nb, st) : launch_attn<32, 0>(mq, mk, mv, a, Nb, st);
  }
  // Kernel choice.  d = 32 with more than two query tiles per (batch, head): v6 (four softmax warpgroups, a quarter
  // of the exponentials on the FMA pipe); otherwise v4 with the MUFU ping-pong.  gvf_attn_set_debug overrides
  // (tools/attn_experiments.py): 0x10 v4 plain, 0x30 v4 ping-pong, 0x80 v6 MUFU only, low nibble 4 = polynomial share.
  const int sel = g_attn_dbg & 0xf0;
  const bool poly = (g_attn_dbg & 0xf) == 4;
  if (D == 32 && (sel == 0x80 || sel == 0x90 || sel == 0xa0 || (sel == 0 && Lq > 256)))
  {
    if (a.trace && (sel == 0 || sel == 0xa0))        // instrumented v8 (tools/attn_trace.py), no key-range split
      return (g_attn_dbg & 0xf) == 1 ? launch_attn6<0, true, false, 0x10000>(mq, mk, mv, a, Nb, st)
                                     : launch_attn6<0, true, false, 0x18888>(mq, mk, mv, a, Nb, st);
    if (a.trace)   // instrumented build (tools/attn_experiments.py)
      return (poly || sel == 0) ? launch_attn6<4, true>(mq, mk, mv, a, Nb, st) : launch_attn6<0, true>(mq, mk, mv, a, Nb, st);
    const int share = g_attn_dbg & 0xf;
    // v8 (warp-uniform MMA issuers) is the default; 0x80 selects v6, 0x90 the persistent v7 (A/B runs).  Low nibble =
    // share of the exponentials evaluated on the FMA pipe.  MEASURED on B200 (tools/attn_variants.py, static / image /
    // spatial us): MUFU only 259 / 107 / 49.6, 1/8 238 / 101 / 47.6, 3/16 226.7 / 96.7 / 45.5, 1/4 227.7 / 96.5 / 45.5,
    // 5/16 231.9 / 98.5, 3/8 239 / 100.8, 1/2 253 / 105 (v6 at its best mix: 252 / 103 / 49.6).
    if (sel == 0 || sel == 0xa0) {
      const int drop = (g_attn_dbg >> 16) & 0xff;    // timing-only ablation builds (tools/attn_variants.py; WRONG results)
#define GVF_V8(M) return launch_attn6<0, false, false, 0x10000 | (M)>(mq, mk, mv, a, Nb, st)
#define GVF_V8D(DR) if (drop == DR) return launch_attn6<0, false, false, 0x18888, DR>(mq, mk, mv, a, Nb, st)
      if (g_attn_dbg & 0x10000000)                   // two-tile CTAs, two CTAs per SM
        return share == 1 ? launch_attn6<0, false, false, 0x10000, 0, 2>(mq, mk, mv, a, Nb, st)
                          : launch_attn6<0, false, false, 0x18888, 0, 2>(mq, mk, mv, a, Nb, st);
      if (drop == 31) return launch_attn6<0, false, false, 0x10000, 31>(mq, mk, mv, a, Nb, st);
      if (drop == 16) return launch_attn6<0, false, false, 0x10000, 16>(mq, mk, mv, a, Nb, st);
      GVF_V8D(1); GVF_V8D(2); GVF_V8D(3); GVF_V8D(4); GVF_V8D(7); GVF_V8D(8);
      switch (share) {
        case 1: GVF_V8(0x0000);
        case 2: GVF_V8(0xaaaa);     // 1/2
        case 3: GVF_V8(0xa8a8);     // 3/8
        case 5: GVF_V8(0xa888);     // 5/16
        case 6: GVF_V8(0x8880);     // 3/16
        case 8: GVF_V8(0x8080);     // 1/8
        default: GVF_V8(0x8888);    // 1/4
      }
#undef GVF_V8
#undef GVF_V8D
    }
    // default: every 8th pair of exponentials on the FMA pipe (bench: 282.0 ms / object; MUFU only 285.6; every
    // 4th pair 285.6).  Low nibble of the debug word: 1 MUFU only, 4 every 4th pair.
    // Short key ranges (spatial self-attention, 512 keys) stay MUFU only: 49.5 vs 53.5 us.
    // Persistent form (v7): only on request (debug selector 0x90).  MEASURED on B200 (bench.py): spatial 49.8 us
    // (v6 49.4), image 112.3 (104.8), static 281.7 (254.8) -- the work-list state costs the issuer warps registers
    // (64 instead of 32, so the softmax warps drop to 104) and the per-item fill / drain of the score pipeline,
    // not the CTA launch, turned out to be what a short item pays for.
    const bool persist = sel == 0x90 && (Lq % 512) == 0 && Lk >= 128 * kAttnStages;
    if (persist) {
      if (share == 4) return launch_attn6<4, false, true>(mq, mk, mv, a, Nb, st);
      if (share == 1 || (share == 0 && Lk < 1024)) return launch_attn6<0, false, true>(mq, mk, mv, a, Nb, st);
      return launch_attn6<8, false, true>(mq, mk, mv, a, Nb, st);
    }
    if (share == 4) return launch_attn6<4, false>(mq, mk, mv, a, Nb, st);
    if (share == 1 || (share == 0 && Lk < 1024)) return launch_attn6<0, false>(mq, mk, mv, a, Nb, st);
    return launch_attn6<8, false>(mq, mk, mv, a, Nb, st);
  }
  if (sel == 0) a.dbg |= 0x20;
  if (D == 64) return launch_attn<64, 0>(mq, mk, mv, a, Nb, st);
  return poly ? launch_attn<32, 4>(mq, mk, mv, a, Nb, st) : launch_attn<32, 0>(mq, mk, mv, a, Nb, st);
}

extern "C" GVF_API int gvf_attn_fwd_f16(const void* q, const void* k, const void* v, void* o, int Nb, int Lq,
                                        int Lk, int H, int D, const long long* q_strides,
                                        const long long* k_strides, const long long* v_strides,
                                        const long long* o_strides, int q_shared, int kv_shared,
                                        float scale, void* stream) {
  return attn_fwd_impl(q, k, v, o, Nb, Lq, Lk, H, D, q_strides, k_strides, v_strides, o_strides, q_shared, kv_shared,
                       scale, nullptr, 0, stream);
}

// Training forward: as above, and LSE2 [Nb, H, lse_ld] fp32 (lse_ld = Lq rounded up to 128) for gvf_attn_bwd_f16.
extern "C" GVF_API int gvf_attn_fwd_lse_f16(const void* q, const void* k, const void* v, void* o, float* lse2, int lse_ld,
                                            int Nb, int Lq, int Lk, int H, int D, const long long* q_strides,
                                            const long long* k_strides, const long long* v_strides,
                                            const long long* o_strides, int q_shared, int kv_shared, float scale,
                                            void* stream) {
  if (!lse2) return GVF_ERR_INVALID;
  return attn_fwd_impl(q, k, v, o, Nb, Lq, Lk, H, D, q_strides, k_strides, v_strides, o_strides, q_shared, kv_shared,
                       scale, lse2, lse_ld, stream);
}
