// sparse_attn.cu -- windowed self-attention over sparse voxels with the window gather and the inverse
// permutation fused into the kernel's own staging loads and stores.
//
// Replaces `sparse_windowed_scaled_dot_product_self_attention` of the reference
// (sparse/attention/windowed_attn.py:61-135), as reached from `SparseMultiHeadAttention.forward`
// (sparse/attention/modules.py:193-196) in the swin blocks of the static VAE:
//     qkv_feats = qkv.feats[fwd_indices]                      # gather copy   [M, 3, H, C]
//     out = flash_attn_varlen_qkvpacked_func(qkv_feats, cu_seqlens, max(seq_lens))
//     out = out[bwd_indices]                                  # scatter copy  [T, H, C]
// Here the rows of a window are fetched straight from qkv.feats through `fwd_idx` by the cp.async that
// stages them into shared memory (whole 128 B head rows, so the gather costs nothing over a dense load),
// and every output row is written to out[fwd_idx[i]] -- which IS out[bwd_indices] -- so neither copy exists.
//
// Windows are short and ragged (1..512 voxels of an 8^3 window, typically tens): a 128-row tcgen05 tile
// would be mostly padding, so this is a warp-level kernel: one CTA = (window, head, 64 query rows), four
// warps x 16 rows, S = Q K^T and O = P V as mma.sync.m16n8k16 on ldmatrix fragments, flash-style running
// maximum over 64-key chunks, K/V chunks double buffered with cp.async.  Head dim 64 (768 / 12 heads).
// fp16 in / out, fp32 scores, statistics and accumulators; P rounded to fp16 for P V.
//
// PACKED form (round 2, the static VAE's trunk): on object surfaces a window holds ~16 of its 512 possible voxels, so one
// CTA per (window, head) is a 64-row tile with 48 rows of padding and 3 000 mostly idle CTAs per launch.  The sequences are
// contiguous ranges of ONE sorted position list, so a CTA can instead take 64 CONSECUTIVE positions -- several whole
// windows -- as its query tile and walk the key range [start of its first window, end of its last window): attention
// becomes block-diagonal over the sorted list, every row carries the [lo, hi) position range of its own window and a key
// column takes part iff its position falls inside it.  4x fewer CTAs, all rows busy: 33 -> ~10 us per launch at 4096 voxels.
#include "../../include/gvf_b200.h"
#include "tc_common.cuh"
#include "tma_host.h"

namespace gvf {
using namespace tc;

namespace {

constexpr int kD = 64, kQB = 64, kKB = 64;     // head dim, query rows per CTA, keys per chunk

__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 16 B piece `chunk` (0..7) of 128 B row `row`: XOR swizzle, conflict-free for ldmatrix and the row copies
__device__ __forceinline__ uint32_t slot(int row, int chunk) { return (uint32_t)(row * 8 + (chunk ^ (row & 7))) * 16u; }

// rows [r0, r0 + 64) of the window (clipped to `len`) of tensor `which` (0 q, 1 k, 2 v), head h -> smem tile
__device__ __forceinline__ void stage_rows(uint32_t dst, const __half* __restrict__ qkv, const int* __restrict__ idx,
                                           int beg, int r0, int len, int which, int h, int H, int tid) {
  const long long row_elems = 3LL * H * kD;
#pragma unroll
  for (int it = 0; it < (64 * 8) / 128; ++it) {
    const int e = it * 128 + tid, r = e >> 3, c = e & 7;
    const bool ok = r0 + r < len;
    const long long g = ok ? (idx ? (long long)__ldg(idx + beg + r0 + r) : (long long)(beg + r0 + r)) : 0;   // no list: identity
    const __half* src = qkv + g * row_elems + ((long long)which * H + h) * kD + c * 8;
    const int bytes = ok ? 16 : 0;                 // src-size 0: the 16 B are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + slot(r, c)), "l"(src), "r"(bytes) : "memory");
  }
}

// TMA staging (TMA = true): the 64 rows of a tile are fetched by 16 `cp.async.bulk.tensor ... tile::gather4` instructions
// (four gathered rows of 128 B each, out-of-range rows zero-filled) through one tensor map over qkv viewed as [T, 3 H 64],
// SWIZZLE_128B -- the same XOR pattern `slot` implements -- completing on an mbarrier, instead of 512 per-thread cp.async
// copies with their address arithmetic.  The gather stays fused into the staging load; only lanes 0-15 of warp 0 issue.
__device__ __forceinline__ void stage_rows_tma(uint32_t dst, const CUtensorMap* map, uint64_t* bar, const int* __restrict__ idx,
                                               int r0, int rend, int col, int tid) {
  if (tid < 16) {
    int rr[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int pos = r0 + 4 * tid + j;
      rr[j] = pos < rend ? (idx ? __ldg(idx + pos) : pos) : -1;          // -1: outside the tensor -> zeros
    }
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst + (uint32_t)tid * 512u),
        "l"(map), "r"(smem_u32(bar)), "r"(col), "r"(rr[0]), "r"(rr[1]), "r"(rr[2]), "r"(rr[3])
        : "memory");
  }
}

// PACKED = false: CTA = (64-row tile blockIdx.x of sequence blockIdx.z, head).  PACKED = true: CTA = (positions
// [64 blockIdx.x, +64) of the whole list, head); seq_of_pos[p] = sequence of position p, M = total positions.
template <bool PACKED, bool TMA>
__global__ void __launch_bounds__(128) sparse_window_attn_kernel(const __grid_constant__ CUtensorMap mapQKV,
                                                                const __half* __restrict__ qkv, __half* __restrict__ out,
                                                                const int* __restrict__ fwd_idx,
                                                                const int* __restrict__ out_idx,
                                                                const int* __restrict__ cu_seqlens,
                                                                const int* __restrict__ seq_of_pos, int M, int H,
                                                                float scale_log2e, float* __restrict__ lse2) {
  __shared__ __align__(1024) uint8_t sm[(1 + 4) * 64 * 128];    // Q | K0 V0 | K1 V1   (40 KB)
  __shared__ uint64_t bar_q, bar_kv[2];
  const int h = blockIdx.y;
  // positions: query tile [p0, pend) (at most 64), key range [kbeg, kend)
  int p0, pend, kbeg, kend;
  if (PACKED) {
    p0 = blockIdx.x * kQB;
    if (p0 >= M) return;
    pend = M;
    kbeg = __ldg(cu_seqlens + __ldg(seq_of_pos + p0));
    kend = __ldg(cu_seqlens + __ldg(seq_of_pos + min(p0 + kQB, M) - 1) + 1);
  } else {
    kbeg = __ldg(cu_seqlens + blockIdx.z);
    kend = __ldg(cu_seqlens + blockIdx.z + 1);
    p0 = kbeg + blockIdx.x * kQB;
    if (p0 >= kend) return;
    pend = kend;
  }
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
  const int* idx = fwd_idx;                      // gather list (nullptr: rows of qkv itself)
  const uint32_t sQ = smem_u32(sm), sKV = sQ + 64 * 128;
  const int nchunks = (kend - kbeg + kKB - 1) / kKB;
  // [lo, hi): positions of the keys this thread's two rows (g, g + 8 of the warp's 16) attend to
  int lo[2], hi[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int p = p0 + 16 * warp + g + 8 * r;
    if (p >= pend) { lo[r] = 0; hi[r] = 0; }
    else if (PACKED) { const int sq = __ldg(seq_of_pos + p); lo[r] = __ldg(cu_seqlens + sq); hi[r] = __ldg(cu_seqlens + sq + 1); }
    else { lo[r] = kbeg; hi[r] = kend; }
  }

  if (TMA) {
    if (tid == 0) {
      mbar_init(&bar_q, 1);
      mbar_init(&bar_kv[0], 1);
      mbar_init(&bar_kv[1], 1);
      fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
      mbar_arrive_expect_tx(&bar_q, 64 * 128);
      mbar_arrive_expect_tx(&bar_kv[0], 2 * 64 * 128);
    }
    __syncwarp();
    stage_rows_tma(sQ, &mapQKV, &bar_q, idx, p0, pend, h * kD, tid);
    stage_rows_tma(sKV, &mapQKV, &bar_kv[0], idx, kbeg, kend, (H + h) * kD, tid);
    stage_rows_tma(sKV + 64 * 128, &mapQKV, &bar_kv[0], idx, kbeg, kend, (2 * H + h) * kD, tid);
  } else {
    stage_rows(sQ, qkv, idx, 0, p0, pend, 0, h, H, tid);
    stage_rows(sKV, qkv, idx, 0, kbeg, kend, 1, h, H, tid);
    stage_rows(sKV + 64 * 128, qkv, idx, 0, kbeg, kend, 2, h, H, tid);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }

  uint32_t qa[4][4];
  float o[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[nt][e] = 0.f;
  float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};

  for (int j = 0; j < nchunks; ++j) {
    const uint32_t bK = sKV + (uint32_t)(j & 1) * 2 * 64 * 128, bV = bK + 64 * 128;
    if (TMA) {
      if (j + 1 < nchunks) {                       // the buffer was released by the __syncthreads that ended chunk j - 1
        const uint32_t nK = sKV + (uint32_t)((j + 1) & 1) * 2 * 64 * 128;
        if (tid == 0) mbar_arrive_expect_tx(&bar_kv[(j + 1) & 1], 2 * 64 * 128);
        __syncwarp();
        stage_rows_tma(nK, &mapQKV, &bar_kv[(j + 1) & 1], idx, kbeg + (j + 1) * kKB, kend, (H + h) * kD, tid);
        stage_rows_tma(nK + 64 * 128, &mapQKV, &bar_kv[(j + 1) & 1], idx, kbeg + (j + 1) * kKB, kend, (2 * H + h) * kD, tid);
      }
      if (j == 0) mbar_wait(&bar_q, 0);
      mbar_wait(&bar_kv[j & 1], (uint32_t)(j >> 1) & 1u);
    } else {
    if (j + 1 < nchunks) {
      const uint32_t nK = sKV + (uint32_t)((j + 1) & 1) * 2 * 64 * 128;
      stage_rows(nK, qkv, idx, 0, kbeg + (j + 1) * kKB, kend, 1, h, H, tid);
      stage_rows(nK + 64 * 128, qkv, idx, 0, kbeg + (j + 1) * kKB, kend, 2, h, H, tid);
      asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    }
    if (j == 0) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) ldsm4(qa[ks], sQ + slot(16 * warp + (lane & 15), 2 * ks + (lane >> 4)));
    }
    // ---- S = Q K^T for this warp's 16 rows x 64 keys
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) s[nt][e] = 0.f;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t kb[4];
        ldsm4(kb, bK + slot(8 * nt + (lane & 7), 4 * hf + (lane >> 3)));
        mma16816(s[nt], qa[2 * hf], kb[0], kb[1]);
        mma16816(s[nt], qa[2 * hf + 1], kb[2], kb[3]);
      }
    }
    // ---- running softmax; rows g (e = 0,1) and g + 8 (e = 2,3), key position = kbeg + j*64 + 8 nt + 2 tg + (e & 1)
    const int kp0 = kbeg + j * kKB + 2 * tg;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int kp = kp0 + 8 * nt + (e & 1);
        s[nt][e] = (kp >= lo[e >> 1] && kp < hi[e >> 1]) ? s[nt][e] * scale_log2e : -INFINITY;
        mx[e >> 1] = fmaxf(mx[e >> 1], s[nt][e]);
      }
    float alpha[2], msub[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float mn = fmaxf(mrow[r], mx[r]);
      // a packed tile's chunk may hold none of this row's keys yet (mn = -inf): subtract 0 then, every p is exp2(-inf) = 0
      msub[r] = (mn == -INFINITY) ? 0.f : mn;
      alpha[r] = ex2f(mrow[r] - msub[r]);
      mrow[r] = mn;
      lrow[r] *= alpha[r];
    }
    uint32_t pa[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      float p[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        p[e] = ex2f(s[nt][e] - msub[e >> 1]);
        lrow[e >> 1] += p[e];
      }
      const __half2 lo = __floats2half2_rn(p[0], p[1]), hi = __floats2half2_rn(p[2], p[3]);
      pa[nt >> 1][(nt & 1) * 2] = *reinterpret_cast<const uint32_t*>(&lo);
      pa[nt >> 1][(nt & 1) * 2 + 1] = *reinterpret_cast<const uint32_t*>(&hi);
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      o[nt][0] *= alpha[0]; o[nt][1] *= alpha[0];
      o[nt][2] *= alpha[1]; o[nt][3] *= alpha[1];
    }
    // ---- O += P V
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
      for (int c2 = 0; c2 < 4; ++c2) {
        uint32_t vb[4];
        ldsm4t(vb, bV + slot(16 * ks + (lane & 7) + 8 * ((lane >> 3) & 1), 2 * c2 + (lane >> 4)));
        mma16816(o[2 * c2], pa[ks], vb[0], vb[1]);
        mma16816(o[2 * c2 + 1], pa[ks], vb[2], vb[3]);
      }
    __syncthreads();                               // everyone is done with this buffer before it is refilled
  }
  // ---- normalise, park the 16 rows in this warp's slice of the Q tile, write whole rows to out[fwd_idx[i]]
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 1);
    lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 2);
  }
  const float inv[2] = {1.0f / lrow[0], 1.0f / lrow[1]};
  if (lse2 && tg == 0) {                           // training: LSE2[row, h] = log2 sum_k exp(scale s) for the backward kernels
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int pos = p0 + 16 * warp + g + 8 * r;
      if (pos < pend) {
        const long long grow = out_idx ? (long long)__ldg(out_idx + pos) : (idx ? (long long)__ldg(idx + pos) : (long long)pos);
        if (grow >= 0) lse2[grow * H + h] = mrow[r] + log2f(lrow[r]);
      }
    }
  }
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int row = 16 * warp + g + 8 * r;
      *reinterpret_cast<__half2*>(sm + slot(row, nt) + 4 * tg) =
          __floats2half2_rn(o[nt][2 * r] * inv[r], o[nt][2 * r + 1] * inv[r]);
    }
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int e = it * 32 + lane, r = 16 * warp + (e >> 3), c = e & 7;
    if (p0 + r < pend) {
      // destination row: the scatter list when given (negative = this position is padding of an overlapping window,
      // serialized attention), else the row the query came from
      const int pos = p0 + r;
      const long long grow = out_idx ? (long long)__ldg(out_idx + pos) : (idx ? (long long)__ldg(idx + pos) : (long long)pos);
      if (grow >= 0)
        *reinterpret_cast<uint4*>(out + (grow * H + h) * kD + c * 8) = *reinterpret_cast<const uint4*>(sm + slot(r, c));
    }
  }
}

}  // namespace
}  // namespace gvf

// qkv [T, 3, H, 64] fp16 (the to_qkv output layout, modules.py:165-166), out [T, H, 64] fp16.
// fwd_idx [M] int32: voxel rows ordered by window; cu_seqlens [W + 1] int32: window w owns
// fwd_idx[cu_seqlens[w] .. cu_seqlens[w + 1]); max_seqlen bounds the grid.  Rows outside every window are
// left untouched.
extern "C" GVF_API int gvf_sparse_window_attn_f16(const void* qkv, void* out, const int* fwd_idx,
                                                  const int* cu_seqlens, int num_windows, int max_seqlen,
                                                  int H, int D, float scale, void* stream) {
  if (!qkv || !out || !fwd_idx || !cu_seqlens || num_windows < 0 || H <= 0 || max_seqlen < 0) return GVF_ERR_INVALID;
  if (D != gvf::kD) return GVF_ERR_UNSUPPORTED;
  if (((uintptr_t)qkv | (uintptr_t)out) & 15) return GVF_ERR_INVALID;
  if (num_windows == 0 || max_seqlen == 0) return GVF_OK;
  if (num_windows > 65535 || H > 65535) return GVF_ERR_UNSUPPORTED;
  const dim3 grid((max_seqlen + gvf::kQB - 1) / gvf::kQB, H, num_windows);
  gvf::sparse_window_attn_kernel<false, false><<<grid, 128, 0, (cudaStream_t)stream>>>(
      CUtensorMap{}, (const __half*)qkv, (__half*)out, fwd_idx, nullptr, cu_seqlens, nullptr, 0, H, scale * 1.4426950408889634f, nullptr);
  return cudaGetLastError() == cudaSuccess ? GVF_OK : GVF_ERR_CUDA;
}

// General form: variable-length self-attention over rows of a packed qkv [T, 3, H, 64] tensor.
//   gather_idx  [M] int32 or NULL: position i of the sequence list reads qkv row gather_idx[i] (NULL: row i itself --
//               sparse full attention, sparse/attention/full_attn.py:90-215 with cu_seqlens = the batch layout);
//   scatter_idx [M] int32 or NULL: its result goes to out row scatter_idx[i], negative = dropped (serialized attention,
//               sparse/attention/serialized_attn.py:38-192: windows are padded to window_size with wrapped-around
//               neighbours whose results are discarded); NULL: the row the query was read from.
extern "C" GVF_API int gvf_sparse_varlen_attn_f16(const void* qkv, void* out, const int* gather_idx, const int* scatter_idx,
                                                  const int* cu_seqlens, int num_seqs, int max_seqlen, int H, int D,
                                                  float scale, void* stream) {
  if (!qkv || !out || !cu_seqlens || num_seqs < 0 || H <= 0 || max_seqlen < 0) return GVF_ERR_INVALID;
  if (D != gvf::kD) return GVF_ERR_UNSUPPORTED;
  if (((uintptr_t)qkv | (uintptr_t)out) & 15) return GVF_ERR_INVALID;
  if (num_seqs == 0 || max_seqlen == 0) return GVF_OK;
  if (num_seqs > 65535 || H > 65535) return GVF_ERR_UNSUPPORTED;
  const dim3 grid((max_seqlen + gvf::kQB - 1) / gvf::kQB, H, num_seqs);
  gvf::sparse_window_attn_kernel<false, false><<<grid, 128, 0, (cudaStream_t)stream>>>(
      CUtensorMap{}, (const __half*)qkv, (__half*)out, gather_idx, scatter_idx, cu_seqlens, nullptr, 0, H, scale * 1.4426950408889634f, nullptr);
  return cudaGetLastError() == cudaSuccess ? GVF_OK : GVF_ERR_CUDA;
}

// Training forward: as above, plus LSE2 [T, H] fp32 (per voxel row) for gvf_sparse_varlen_attn_bwd_f16.
extern "C" GVF_API int gvf_sparse_varlen_attn_lse_f16(const void* qkv, void* out, float* lse2, const int* gather_idx,
                                                      const int* scatter_idx, const int* cu_seqlens, int num_seqs,
                                                      int max_seqlen, int H, int D, float scale, void* stream) {
  if (!qkv || !out || !lse2 || !cu_seqlens || num_seqs < 0 || H <= 0 || max_seqlen < 0) return GVF_ERR_INVALID;
  if (D != gvf::kD) return GVF_ERR_UNSUPPORTED;
  if (((uintptr_t)qkv | (uintptr_t)out) & 15) return GVF_ERR_INVALID;
  if (num_seqs == 0 || max_seqlen == 0) return GVF_OK;
  if (num_seqs > 65535 || H > 65535) return GVF_ERR_UNSUPPORTED;
  const dim3 grid((max_seqlen + gvf::kQB - 1) / gvf::kQB, H, num_seqs);
  gvf::sparse_window_attn_kernel<false, false><<<grid, 128, 0, (cudaStream_t)stream>>>(
      CUtensorMap{}, (const __half*)qkv, (__half*)out, gather_idx, scatter_idx, cu_seqlens, nullptr, 0, H, scale * 1.4426950408889634f, lse2);
  return cudaGetLastError() == cudaSuccess ? GVF_OK : GVF_ERR_CUDA;
}

// Packed form of the three entry points above (see the file header): the sequences must be contiguous ranges of one
// position list of M entries, seq_of_pos [M] int32 = the sequence a position belongs to (non-decreasing).  One CTA per 64
// consecutive positions and head, whatever the sequence lengths.  lse2 may be NULL (inference).
// 0 (default): per-thread cp.async staging, 1: TMA tile::gather4 staging.  Measured on B200 at the static VAE's shape
// (tools/window_attn_bench.py, 2 x 2048 surface voxels, 12 heads): per-window tiling 30.7 us, packed + cp.async 17.7-18.3 us,
// packed + TMA 18.8-18.9 us -- the kernel is bound by its chain of dependent index loads, not by the staging instructions,
// so the bulk-tensor path buys nothing here and stays an option.
static int g_sparse_attn_tma = 0;
extern "C" GVF_API void gvf_sparse_attn_set_tma(int on) { g_sparse_attn_tma = on ? 1 : 0; }

extern "C" GVF_API int gvf_sparse_packed_attn_f16(const void* qkv, void* out, float* lse2, const int* gather_idx,
                                                  const int* scatter_idx, const int* cu_seqlens, const int* seq_of_pos, int M,
                                                  int H, int D, float scale, void* stream) {
  const int T_rows = M;               // windowed / full attention: the packed list is a permutation of the tensor's rows
  if (!qkv || !out || !cu_seqlens || !seq_of_pos || M < 0 || H <= 0) return GVF_ERR_INVALID;
  if (D != gvf::kD) return GVF_ERR_UNSUPPORTED;
  if (((uintptr_t)qkv | (uintptr_t)out) & 15) return GVF_ERR_INVALID;
  if (M == 0) return GVF_OK;
  if (H > 65535) return GVF_ERR_UNSUPPORTED;
  const dim3 grid((M + gvf::kQB - 1) / gvf::kQB, H, 1);
  if (g_sparse_attn_tma) {
    // qkv as a 2-D tensor [rows, 3 H 64]: the gather lists address its rows, so the map must cover the largest row index --
    // the caller's tensor has at least M rows (windowed / full attention: exactly M)
    CUtensorMap map;
    if (!gvf::make_tmap_2d(&map, qkv, 2, (uint64_t)3 * H * gvf::kD, (uint64_t)(T_rows > 0 ? T_rows : M), (uint64_t)3 * H * gvf::kD, 64, 1))
      return GVF_ERR_CUDA;
    gvf::sparse_window_attn_kernel<true, true><<<grid, 128, 0, (cudaStream_t)stream>>>(
        map, (const __half*)qkv, (__half*)out, gather_idx, scatter_idx, cu_seqlens, seq_of_pos, M, H, scale * 1.4426950408889634f, lse2);
  } else {
    gvf::sparse_window_attn_kernel<true, false><<<grid, 128, 0, (cudaStream_t)stream>>>(
        CUtensorMap{}, (const __half*)qkv, (__half*)out, gather_idx, scatter_idx, cu_seqlens, seq_of_pos, M, H,
        scale * 1.4426950408889634f, lse2);
  }
  return cudaGetLastError() == cudaSuccess ? GVF_OK : GVF_ERR_CUDA;
}
