// sparse_attn_bwd.cu -- backward of the windowed / full sparse self-attention (csrc/sparse_attn.cu), for the training
// step of the static SparseTransformerVAE (reference sparse_transformer.py:172-192 under autograd,
// flash_attn_varlen_qkvpacked_func's backward at sparse/attention/windowed_attn.py:125-127).
//
// Same warp-level design as the forward (windows are 1..512 voxels, typically tens: a 128-row tcgen05 tile would be
// mostly padding): rows are gathered through the window list by the cp.async staging loads, `mma.sync.m16n8k16`, fp16
// operands, fp32 accumulation.  The forward leaves LSE2[t, h] = log2 sum_k exp(scale s_tk) per voxel row, so P is recomputed
// exactly.  Two kernels, no atomics:
//   sparse_attn_bwd_dq_kernel    CTA = (window, head, 64 query rows): S = Q K^T, dP = dO V^T per 64-key chunk,
//                                dS = P o (dP - D), dQ += dS K  (dS re-used as the A fragment, K via ldmatrix.trans)
//   sparse_attn_bwd_dkdv_kernel  CTA = (window, head, 64 key rows): the transposed problem -- S^T = K Q^T, dP^T = V dO^T
//                                per 64-query chunk (LSE2 / D of the chunk's queries in shared memory, indexed by column),
//                                dV += P^T dO, dK += dS^T Q
// D[t, h] = sum_d dO O is a small pre-pass.  Gradients land in a packed [T, 3, H, 64] tensor at the voxel's own row
// (the inverse permutation is fused like in the forward).
#include "../../include/gvf_b200.h"
#include "tc_common.cuh"

namespace gvf {
using namespace tc;

namespace {

constexpr int kD = 64;

__device__ __forceinline__ void b_ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void b_ldsm4t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void b_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float b_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t b_slot(int row, int chunk) { return (uint32_t)(row * 8 + (chunk ^ (row & 7))) * 16u; }

// rows [r0, r0 + 64) of the window list (clipped to len) -> swizzled [64 x 128 B] tile.  src row g: base + g * row_elems.
__device__ __forceinline__ void b_stage(uint32_t dst, const __half* __restrict__ base, long long row_elems,
                                        const int* __restrict__ idx, int beg, int r0, int len, int tid) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int e = it * 128 + tid, r = e >> 3, c = e & 7;
    const bool ok = r0 + r < len;
    const long long g = ok ? (idx ? (long long)__ldg(idx + beg + r0 + r) : (long long)(beg + r0 + r)) : 0;
    const __half* src = base + g * row_elems + c * 8;
    const int bytes = ok ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + b_slot(r, c)), "l"(src), "r"(bytes) : "memory");
  }
}
__device__ __forceinline__ long long b_row(const int* idx, int pos) { return idx ? (long long)__ldg(idx + pos) : (long long)pos; }

// 16 x 64 product of this warp's A fragments with a K-major 64-row tile: c[nt] (+)= A (16 x 64) . tile[8 nt .. 8 nt + 8)^T
__device__ __forceinline__ void b_mm_kmajor(float (&c)[8][4], const uint32_t (&a)[4][4], uint32_t tile, int lane) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
    for (int e = 0; e < 4; ++e) c[nt][e] = 0.f;
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      uint32_t kb[4];
      b_ldsm4(kb, tile + b_slot(8 * nt + (lane & 7), 4 * hf + (lane >> 3)));
      b_mma(c[nt], a[2 * hf], kb[0], kb[1]);
      b_mma(c[nt], a[2 * hf + 1], kb[2], kb[3]);
    }
  }
}
// acc (16 x 64 over d) += A (16 x 64 over the tile's rows) . tile (64 rows x 64 d), tile read with ldmatrix.trans
__device__ __forceinline__ void b_mm_rows(float (&acc)[8][4], const uint32_t (&a)[4][4], uint32_t tile, int lane) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
#pragma unroll
    for (int c2 = 0; c2 < 4; ++c2) {
      uint32_t vb[4];
      b_ldsm4t(vb, tile + b_slot(16 * ks + (lane & 7) + 8 * ((lane >> 3) & 1), 2 * c2 + (lane >> 4)));
      b_mma(acc[2 * c2], a[ks], vb[0], vb[1]);
      b_mma(acc[2 * c2 + 1], a[ks], vb[2], vb[3]);
    }
}
// C-fragment values (fp32, 16 x 64) -> A fragments (fp16) of the next product
__device__ __forceinline__ void b_pack(uint32_t (&a)[4][4], const float (&v)[8][4]) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const __half2 lo = __floats2half2_rn(v[nt][0], v[nt][1]), hi = __floats2half2_rn(v[nt][2], v[nt][3]);
    a[nt >> 1][(nt & 1) * 2] = *reinterpret_cast<const uint32_t*>(&lo);
    a[nt >> 1][(nt & 1) * 2 + 1] = *reinterpret_cast<const uint32_t*>(&hi);
  }
}
// normalise nothing, scale, park the warp's 16 x 64 tile in smem (own slice) and write whole 128 B rows to dst rows
__device__ __forceinline__ void b_store_rows(uint8_t* sm, const float (&acc)[8][4], float mul, __half* __restrict__ dst,
                                             long long row_elems, const int* idx, int beg, int r0, int len, int warp, int lane) {
  const int g = lane >> 2, tg = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int row = 16 * warp + g + 8 * r;
      *reinterpret_cast<__half2*>(sm + b_slot(row, nt) + 4 * tg) = __floats2half2_rn(acc[nt][2 * r] * mul, acc[nt][2 * r + 1] * mul);
    }
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int e = it * 32 + lane, r = 16 * warp + (e >> 3), c = e & 7;
    if (r0 + r < len) {
      const long long grow = b_row(idx, beg + r0 + r);
      *reinterpret_cast<uint4*>(dst + grow * row_elems + c * 8) = *reinterpret_cast<const uint4*>(sm + b_slot(r, c));
    }
  }
}

// D[t, h] = sum_d dO[t, h, d] O[t, h, d]
__global__ void __launch_bounds__(256) sparse_attn_bwd_prep_kernel(const __half* __restrict__ o, const __half* __restrict__ dout,
                                                                   long long n, float* __restrict__ dsum) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4* po = reinterpret_cast<const uint4*>(o + i * kD);
  const uint4* pg = reinterpret_cast<const uint4*>(dout + i * kD);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < kD / 8; ++j) {
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
  dsum[i] = s;
}

// Tile set-up shared by the two kernels.  PACKED = false: 64-row tile blockIdx.x of sequence blockIdx.z, the other operand
// ranges over that sequence.  PACKED = true (see csrc/sparse_attn.cu): 64 consecutive positions of the whole list, the other
// operand ranges over [start of the first, end of the last sequence the tile touches), rows masked by their own range.
template <bool PACKED>
__device__ __forceinline__ bool b_tile(const int* __restrict__ cu, const int* __restrict__ seq_of_pos, int M, int& p0, int& pend,
                                       int& obeg, int& oend) {
  if (PACKED) {
    p0 = blockIdx.x * 64;
    if (p0 >= M) return false;
    pend = M;
    obeg = __ldg(cu + __ldg(seq_of_pos + p0));
    oend = __ldg(cu + __ldg(seq_of_pos + min(p0 + 64, M) - 1) + 1);
  } else {
    obeg = __ldg(cu + blockIdx.z);
    oend = __ldg(cu + blockIdx.z + 1);
    p0 = obeg + blockIdx.x * 64;
    if (p0 >= oend) return false;
    pend = oend;
  }
  return true;
}
template <bool PACKED>
__device__ __forceinline__ void b_row_range(const int* __restrict__ cu, const int* __restrict__ seq_of_pos, int p, int pend, int obeg,
                                            int oend, int& lo, int& hi) {
  if (p >= pend) { lo = 0; hi = 0; }
  else if (PACKED) { const int sq = __ldg(seq_of_pos + p); lo = __ldg(cu + sq); hi = __ldg(cu + sq + 1); }
  else { lo = obeg; hi = oend; }
}

template <bool PACKED>
__global__ void __launch_bounds__(128) sparse_attn_bwd_dq_kernel(const __half* __restrict__ qkv, const __half* __restrict__ dout,
                                                                const float* __restrict__ lse2, const float* __restrict__ dsum,
                                                                __half* __restrict__ dqkv, const int* __restrict__ idx,
                                                                const int* __restrict__ cu, const int* __restrict__ seq_of_pos,
                                                                int M, int H, float scale, float scale_log2e) {
  extern __shared__ __align__(128) uint8_t sm[];                // Q | dO | K0 V0 | K1 V1   (48 KB, dynamic)
  const int h = blockIdx.y;
  int q0, len, kbeg, kend;                                      // positions: queries [q0, len), keys [kbeg, kend)
  if (!b_tile<PACKED>(cu, seq_of_pos, M, q0, len, kbeg, kend)) return;
  const int beg = 0;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
  int lo[2], hi[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) b_row_range<PACKED>(cu, seq_of_pos, q0 + 16 * warp + g + 8 * r, len, kbeg, kend, lo[r], hi[r]);
  const long long qrow = 3LL * H * kD, orow = (long long)H * kD;
  const __half* qb = qkv + (long long)h * kD;
  const __half* kb_ = qkv + ((long long)H + h) * kD;
  const __half* vb_ = qkv + (2LL * H + h) * kD;
  const uint32_t sQ = smem_u32(sm), sDO = sQ + 64 * 128, sKV = sDO + 64 * 128;
  const int nchunks = (kend - kbeg + 63) / 64;
  b_stage(sQ, qb, qrow, idx, beg, q0, len, tid);
  b_stage(sDO, dout + (long long)h * kD, orow, idx, beg, q0, len, tid);
  b_stage(sKV, kb_, qrow, idx, beg, kbeg, kend, tid);
  b_stage(sKV + 64 * 128, vb_, qrow, idx, beg, kbeg, kend, tid);
  asm volatile("cp.async.commit_group;" ::: "memory");
  // statistics of this thread's two rows
  float lse_r[2], d_r[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = q0 + 16 * warp + g + 8 * r;
    const bool ok = row < len;
    const long long grow = ok ? b_row(idx, beg + row) : 0;
    lse_r[r] = ok ? __ldg(lse2 + grow * H + h) : INFINITY;
    d_r[r] = ok ? __ldg(dsum + grow * H + h) : 0.f;
  }
  uint32_t qa[4][4], doa[4][4];
  float dq[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) dq[nt][e] = 0.f;
  for (int j = 0; j < nchunks; ++j) {
    const uint32_t bK = sKV + (uint32_t)(j & 1) * 2 * 64 * 128, bV = bK + 64 * 128;
    if (j + 1 < nchunks) {
      const uint32_t nK = sKV + (uint32_t)((j + 1) & 1) * 2 * 64 * 128;
      b_stage(nK, kb_, qrow, idx, beg, kbeg + (j + 1) * 64, kend, tid);
      b_stage(nK + 64 * 128, vb_, qrow, idx, beg, kbeg + (j + 1) * 64, kend, tid);
      asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (j == 0) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        b_ldsm4(qa[ks], sQ + b_slot(16 * warp + (lane & 15), 2 * ks + (lane >> 4)));
        b_ldsm4(doa[ks], sDO + b_slot(16 * warp + (lane & 15), 2 * ks + (lane >> 4)));
      }
    }
    float s[8][4], dp[8][4];
    b_mm_kmajor(s, qa, bK, lane);                  // S = Q K^T
    b_mm_kmajor(dp, doa, bV, lane);                // dP = dO V^T
    const int kp0 = kbeg + j * 64 + 2 * tg;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int kp = kp0 + 8 * nt + (e & 1);
        const float p = (kp >= lo[e >> 1] && kp < hi[e >> 1]) ? b_ex2(fmaf(s[nt][e], scale_log2e, -lse_r[e >> 1])) : 0.f;
        s[nt][e] = p * (dp[nt][e] - d_r[e >> 1]);  // dS
      }
    uint32_t dsa[4][4];
    b_pack(dsa, s);
    b_mm_rows(dq, dsa, bK, lane);                  // dQ += dS K
    __syncthreads();
  }
  b_store_rows(sm, dq, scale, dqkv + (long long)h * kD, qrow, idx, beg, q0, len, warp, lane);
}

template <bool PACKED>
__global__ void __launch_bounds__(128) sparse_attn_bwd_dkdv_kernel(const __half* __restrict__ qkv, const __half* __restrict__ dout,
                                                                  const float* __restrict__ lse2, const float* __restrict__ dsum,
                                                                  __half* __restrict__ dqkv, const int* __restrict__ idx,
                                                                  const int* __restrict__ cu, const int* __restrict__ seq_of_pos,
                                                                  int M, int H, float scale, float scale_log2e) {
  extern __shared__ __align__(128) uint8_t sm[];                // K | V | Q0 dO0 | Q1 dO1 (48 KB) | LSE2, D of two chunks
  float (*s_lse)[64] = reinterpret_cast<float (*)[64]>(sm + 6 * 64 * 128);
  float (*s_d)[64] = reinterpret_cast<float (*)[64]>(sm + 6 * 64 * 128 + 512);
  const int h = blockIdx.y;
  int k0, len, qbeg, qend;                                      // positions: keys [k0, len), queries [qbeg, qend)
  if (!b_tile<PACKED>(cu, seq_of_pos, M, k0, len, qbeg, qend)) return;
  const int beg = 0;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, tg = lane & 3;
  int lo[2], hi[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) b_row_range<PACKED>(cu, seq_of_pos, k0 + 16 * warp + (lane >> 2) + 8 * r, len, qbeg, qend, lo[r], hi[r]);
  const long long qrow = 3LL * H * kD, orow = (long long)H * kD;
  const __half* qb = qkv + (long long)h * kD;
  const __half* kb_ = qkv + ((long long)H + h) * kD;
  const __half* vb_ = qkv + (2LL * H + h) * kD;
  const __half* dob = dout + (long long)h * kD;
  const uint32_t sK = smem_u32(sm), sV = sK + 64 * 128, sQD = sV + 64 * 128;
  const int nchunks = (qend - qbeg + 63) / 64;
  auto stage_stats = [&](int buf, int r0) {
    if (tid < 64) {
      const bool ok = r0 + tid < qend;
      const long long grow = ok ? b_row(idx, beg + r0 + tid) : 0;
      s_lse[buf][tid] = ok ? __ldg(lse2 + grow * H + h) : INFINITY;
      s_d[buf][tid] = ok ? __ldg(dsum + grow * H + h) : 0.f;
    }
  };
  b_stage(sK, kb_, qrow, idx, beg, k0, len, tid);
  b_stage(sV, vb_, qrow, idx, beg, k0, len, tid);
  b_stage(sQD, qb, qrow, idx, beg, qbeg, qend, tid);
  b_stage(sQD + 64 * 128, dob, orow, idx, beg, qbeg, qend, tid);
  asm volatile("cp.async.commit_group;" ::: "memory");
  stage_stats(0, qbeg);
  uint32_t ka[4][4], va[4][4];
  float dk[8][4], dv[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) { dk[nt][e] = 0.f; dv[nt][e] = 0.f; }
  for (int j = 0; j < nchunks; ++j) {
    const uint32_t bQ = sQD + (uint32_t)(j & 1) * 2 * 64 * 128, bDO = bQ + 64 * 128;
    if (j + 1 < nchunks) {
      const uint32_t nQ = sQD + (uint32_t)((j + 1) & 1) * 2 * 64 * 128;
      b_stage(nQ, qb, qrow, idx, beg, qbeg + (j + 1) * 64, qend, tid);
      b_stage(nQ + 64 * 128, dob, orow, idx, beg, qbeg + (j + 1) * 64, qend, tid);
      stage_stats((j + 1) & 1, qbeg + (j + 1) * 64);
      asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (j == 0) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        b_ldsm4(ka[ks], sK + b_slot(16 * warp + (lane & 15), 2 * ks + (lane >> 4)));
        b_ldsm4(va[ks], sV + b_slot(16 * warp + (lane & 15), 2 * ks + (lane >> 4)));
      }
    }
    float s[8][4], dp[8][4];
    b_mm_kmajor(s, ka, bQ, lane);                  // S^T = K Q^T   (rows = keys, columns = queries of the chunk)
    b_mm_kmajor(dp, va, bDO, lane);                // dP^T = V dO^T
    const float* lse = s_lse[j & 1];
    const float* dd = s_d[j & 1];
    const int qp0 = qbeg + j * 64;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = 8 * nt + 2 * tg + (e & 1);
        const int qp = qp0 + col;                  // the query must sit in this key row's own sequence (packed tiles)
        const float p = (!PACKED || (qp >= lo[e >> 1] && qp < hi[e >> 1]))
                            ? b_ex2(fmaf(s[nt][e], scale_log2e, -lse[col])) : 0.f;   // +inf for padded queries -> 0
        s[nt][e] = p;
        dp[nt][e] = p * (dp[nt][e] - dd[col]);     // dS^T
      }
    uint32_t pa[4][4], dsa[4][4];
    b_pack(pa, s);
    b_pack(dsa, dp);
    b_mm_rows(dv, pa, bDO, lane);                  // dV += P^T dO
    b_mm_rows(dk, dsa, bQ, lane);                  // dK += dS^T Q
    __syncthreads();
  }
  b_store_rows(sm, dk, scale, dqkv + ((long long)H + h) * kD, qrow, idx, beg, k0, len, warp, lane);
  __syncwarp();
  b_store_rows(sm + 64 * 128, dv, 1.0f, dqkv + (2LL * H + h) * kD, qrow, idx, beg, k0, len, warp, lane);
}

}  // namespace
}  // namespace gvf

// Backward of gvf_sparse_varlen_attn_f16 for bijective lists (windowed / full attention: every voxel row appears exactly
// once; the serialized form with padded windows is forward-only).  qkv [T, 3, H, 64], dout [T, H, 64] (gradient of the
// attention output in voxel order), o [T, H, 64] the forward output, lse2 [T, H] from the forward, dsum [T, H] scratch,
// dqkv [T, 3, H, 64] receives dq | dk | dv in voxel order.  seq_of_pos != NULL selects the packed tiling (M positions).
static int sparse_attn_bwd_impl(const void* qkv, const void* o, const void* dout, const float* lse2, float* dsum, void* dqkv,
                                const int* gather_idx, const int* cu_seqlens, const int* seq_of_pos, int M, int num_seqs,
                                int max_seqlen, long long T, int H, int D, float scale, void* stream) {
  if (!qkv || !o || !dout || !lse2 || !dsum || !dqkv || !cu_seqlens || num_seqs < 0 || H <= 0 || T <= 0) return GVF_ERR_INVALID;
  if (D != gvf::kD) return GVF_ERR_UNSUPPORTED;
  if (((uintptr_t)qkv | (uintptr_t)o | (uintptr_t)dout | (uintptr_t)dqkv) & 15) return GVF_ERR_INVALID;
  const bool packed = seq_of_pos != nullptr;
  if (packed ? M <= 0 : (num_seqs == 0 || max_seqlen <= 0)) return GVF_OK;
  if (num_seqs > 65535 || H > 65535) return GVF_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = T * H;
  gvf::sparse_attn_bwd_prep_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const __half*)o, (const __half*)dout, n, dsum);
  if (cudaGetLastError() != cudaSuccess) return GVF_ERR_CUDA;
  const float sl2 = scale * 1.4426950408889634f;
  constexpr int SMEM_DQ = 6 * 64 * 128, SMEM_KV = 6 * 64 * 128 + 1024;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(gvf::sparse_attn_bwd_dq_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DQ) != cudaSuccess ||
        cudaFuncSetAttribute(gvf::sparse_attn_bwd_dkdv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_KV) != cudaSuccess ||
        cudaFuncSetAttribute(gvf::sparse_attn_bwd_dq_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DQ) != cudaSuccess ||
        cudaFuncSetAttribute(gvf::sparse_attn_bwd_dkdv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_KV) != cudaSuccess)
      return GVF_ERR_CUDA;
    configured = true;
  }
  if (packed) {
    const dim3 grid((M + 63) / 64, H, 1);
    gvf::sparse_attn_bwd_dq_kernel<true><<<grid, 128, SMEM_DQ, st>>>((const __half*)qkv, (const __half*)dout, lse2, dsum, (__half*)dqkv,
                                                                    gather_idx, cu_seqlens, seq_of_pos, M, H, scale, sl2);
    if (cudaGetLastError() != cudaSuccess) return GVF_ERR_CUDA;
    gvf::sparse_attn_bwd_dkdv_kernel<true><<<grid, 128, SMEM_KV, st>>>((const __half*)qkv, (const __half*)dout, lse2, dsum,
                                                                      (__half*)dqkv, gather_idx, cu_seqlens, seq_of_pos, M, H, scale, sl2);
    return cudaGetLastError() == cudaSuccess ? GVF_OK : GVF_ERR_CUDA;
  }
  const dim3 grid((max_seqlen + 63) / 64, H, num_seqs);
  gvf::sparse_attn_bwd_dq_kernel<false><<<grid, 128, SMEM_DQ, st>>>((const __half*)qkv, (const __half*)dout, lse2, dsum, (__half*)dqkv,
                                                                   gather_idx, cu_seqlens, nullptr, 0, H, scale, sl2);
  if (cudaGetLastError() != cudaSuccess) return GVF_ERR_CUDA;
  gvf::sparse_attn_bwd_dkdv_kernel<false><<<grid, 128, SMEM_KV, st>>>((const __half*)qkv, (const __half*)dout, lse2, dsum,
                                                                     (__half*)dqkv, gather_idx, cu_seqlens, nullptr, 0, H, scale, sl2);
  return cudaGetLastError() == cudaSuccess ? GVF_OK : GVF_ERR_CUDA;
}

extern "C" GVF_API int gvf_sparse_varlen_attn_bwd_f16(const void* qkv, const void* o, const void* dout, const float* lse2,
                                                      float* dsum, void* dqkv, const int* gather_idx, const int* cu_seqlens,
                                                      int num_seqs, int max_seqlen, long long T, int H, int D, float scale,
                                                      void* stream) {
  return sparse_attn_bwd_impl(qkv, o, dout, lse2, dsum, dqkv, gather_idx, cu_seqlens, nullptr, 0, num_seqs, max_seqlen, T, H, D,
                              scale, stream);
}

// Packed tiling of the same backward (csrc/sparse_attn.cu, PACKED): seq_of_pos [M] int32, M = positions in the list.
extern "C" GVF_API int gvf_sparse_packed_attn_bwd_f16(const void* qkv, const void* o, const void* dout, const float* lse2,
                                                      float* dsum, void* dqkv, const int* gather_idx, const int* cu_seqlens,
                                                      const int* seq_of_pos, int M, long long T, int H, int D, float scale,
                                                      void* stream) {
  if (!seq_of_pos) return GVF_ERR_INVALID;
  return sparse_attn_bwd_impl(qkv, o, dout, lse2, dsum, dqkv, gather_idx, cu_seqlens, seq_of_pos, M, 1, 1, T, H, D, scale, stream);
}
