// sparse_trunk.cu -- native driver of a stack of SparseTransformerBlocks (reference
// model/sparse_voxel_diffusion/sparse_transformer.py:126-192, stacked at sparse_transformer_vae.py:55-91): the launch
// sequence of the static VAE's encoder / decoder trunk -- inference forward, training forward with the activations kept in
// a caller-owned arena, and the hand-written backward -- issued from C++ instead of one Python call per kernel.
//
// Why: at the reference's per-GPU batch (2 objects x 2048 voxels = 4096 tokens) a block is ~7 (forward) / ~25 (backward)
// kernels of 5-25 us each; driven from Python (ctypes marshalling + a torch allocation per launch, ~13 us) the step was
// host-bound: 15.0 ms to enqueue 1100 launches against 16.1 ms until the GPU finished (a perf_counter probe around the un-synchronised step).  Here a
// launch costs the CUDA runtime's ~2 us plus a cuTensorMapEncode per GEMM operand, and no allocator call: every
// intermediate lives in the arena / scratch the caller passes in.
//
// No kernels of its own: every step is one of the library's C-ABI entry points (csrc/elementwise.cu LayerNorm,
// csrc/gemm.cu tcgen05 GEMMs incl. the TN wgrad and the GELU / GELU' epilogues, csrc/sparse_attn*.cu window attention,
// csrc/backward.cu LayerNorm backward and bias sums).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gvf_b200.h"

namespace {

constexpr size_t kAlign = 256;
inline size_t up(size_t n) { return (n + kAlign - 1) / kAlign * kAlign; }

// activations one block keeps for its backward (the block input is the previous block's output slot)
struct BlockSlots {
  uint8_t *x1, *A, *QKV, *AO, *lse, *A2, *H0, *Hg;
};

struct Layout {
  size_t X, A, QKV, LSE, Hf;       // bytes of one [T,C] residual, [T,C] fp16, [T,3C] fp16, [T,H] fp32, [T,F] fp16 tensor
  size_t per_block, total;
  Layout(int T, int C, int H, int F, int nb, int fp16_residual) {
    const size_t rs = fp16_residual ? 2 : 4;
    X = up((size_t)T * C * rs);
    A = up((size_t)T * C * 2);
    QKV = up((size_t)T * 3 * C * 2);
    LSE = up((size_t)T * H * 4);
    Hf = up((size_t)T * F * 2);
    per_block = X /*x1*/ + A + QKV + A /*AO*/ + LSE + A /*A2*/ + 2 * Hf;
    total = (size_t)(nb + 1) * X + (size_t)nb * per_block;
  }
  uint8_t* xslot(uint8_t* arena, int i) const { return arena + (size_t)i * X; }
  BlockSlots block(uint8_t* arena, int nb, int i) const {
    uint8_t* p = arena + (size_t)(nb + 1) * X + (size_t)i * per_block;
    BlockSlots s;
    s.x1 = p; p += X;
    s.A = p; p += A;
    s.QKV = p; p += QKV;
    s.AO = p; p += A;
    s.lse = p; p += LSE;
    s.A2 = p; p += A;
    s.H0 = p; p += Hf;
    s.Hg = p;
    return s;
  }
};

#define GVF_TRY(expr)              \
  do {                             \
    const int st__ = (expr);       \
    if (st__ != GVF_OK) return st__; \
  } while (0)

inline int copy_async(void* dst, const void* src, size_t bytes, void* stream) {
  return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream) == cudaSuccess ? GVF_OK : GVF_ERR_CUDA;
}

inline bool bad_shape(int T, int C, int H, int F, int nb) {
  return T <= 0 || nb <= 0 || H <= 0 || C != 64 * H || (C % 8) || (F % 8) || F <= 0;
}

}  // namespace

extern "C" {

GVF_API size_t gvf_sparse_trunk_arena_bytes(int T, int C, int H, int F, int num_blocks, int fp16_residual) {
  if (bad_shape(T, C, H, F, num_blocks)) return 0;
  return Layout(T, C, H, F, num_blocks, fp16_residual).total;
}

GVF_API size_t gvf_sparse_trunk_scratch_bytes(int T, int C, int H, int F) {
  if (bad_shape(T, C, H, F, 1)) return 0;
  const Layout L(T, C, H, F, 1, 1);
  // backward: dx x 3, dx1 x 2, dA2, dAO, dA (8 A), dQKV x 2, dH0 x 2, dsum; inference forward: A16, QKV, AO, H1 (smaller)
  return 8 * L.A + 2 * L.QKV + 2 * L.Hf + L.LSE;
}

GVF_API int gvf_sparse_trunk_forward(const gvf_sparse_block* blocks, int num_blocks, int T, int C, int H, int F,
                                     int fp16_residual, const gvf_window_partition* parts, const void* x_in, void* arena,
                                     size_t arena_bytes, void* scratch, size_t scratch_bytes, void* x_out, void* stream) {
  if (!blocks || !parts || !x_in || !x_out || bad_shape(T, C, H, F, num_blocks)) return GVF_ERR_INVALID;
  const Layout L(T, C, H, F, num_blocks, fp16_residual);
  const size_t xbytes = (size_t)T * C * (fp16_residual ? 2 : 4);
  const int epi = fp16_residual ? 3 : 2;
  const float scale = 0.125f;                         // 1 / sqrt(64)
  if (arena) {
    // ---------------- training forward: everything the backward reads stays in the arena
    if (arena_bytes < L.total) return GVF_ERR_WORKSPACE;
    uint8_t* ar = (uint8_t*)arena;
    GVF_TRY(copy_async(L.xslot(ar, 0), x_in, xbytes, stream));
    for (int i = 0; i < num_blocks; ++i) {
      const gvf_sparse_block& b = blocks[i];
      const gvf_window_partition& pt = parts[i & 1];
      const BlockSlots s = L.block(ar, num_blocks, i);
      uint8_t *x0 = L.xslot(ar, i), *x2 = L.xslot(ar, i + 1);
      GVF_TRY(gvf_ln_mod_f16(x0, fp16_residual, s.A, T, C, 1e-6f, nullptr, nullptr, nullptr, nullptr, 0, 0, stream));
      GVF_TRY(gvf_gemm_f16(s.A, C, b.w_qkv, C, T, 3 * C, C, 0, b.b_qkv, s.QKV, 3 * C, nullptr, 0, 0, stream));
      if (pt.seq_of_pos)
        GVF_TRY(gvf_sparse_packed_attn_f16(s.QKV, s.AO, (float*)s.lse, pt.fwd_idx, nullptr, pt.cu_seqlens, pt.seq_of_pos, T, H, 64,
                                           scale, stream));
      else
        GVF_TRY(gvf_sparse_varlen_attn_lse_f16(s.QKV, s.AO, (float*)s.lse, pt.fwd_idx, nullptr, pt.cu_seqlens, pt.num_windows,
                                               pt.max_seqlen, H, 64, scale, stream));
      GVF_TRY(copy_async(s.x1, x0, xbytes, stream));
      GVF_TRY(gvf_gemm_f16(s.AO, C, b.w_out, C, T, C, C, epi, b.b_out, s.x1, C, nullptr, 0, 0, stream));
      GVF_TRY(gvf_ln_mod_f16(s.x1, fp16_residual, s.A2, T, C, 1e-6f, nullptr, nullptr, nullptr, nullptr, 0, 0, stream));
      // fc1 + GELU(tanh) in the epilogue, fp16 pre-activation kept for GELU'
      GVF_TRY(gvf_gemm_f16(s.A2, C, b.w1, C, T, F, C, 1, b.b1, s.Hg, F, s.H0, F, 0, stream));
      GVF_TRY(copy_async(x2, s.x1, xbytes, stream));
      GVF_TRY(gvf_gemm_f16(s.Hg, F, b.w2, F, T, C, F, epi, b.b2, x2, C, nullptr, 0, 0, stream));
    }
    return copy_async(x_out, L.xslot(ar, num_blocks), xbytes, stream);
  }
  // ---------------- inference forward: the residual stream is updated in place in x_out
  if (!scratch || scratch_bytes < 2 * L.A + L.QKV + L.Hf) return GVF_ERR_WORKSPACE;
  uint8_t* sc = (uint8_t*)scratch;
  uint8_t *A16 = sc, *AO = sc + L.A, *QKV = sc + 2 * L.A, *H1 = sc + 2 * L.A + L.QKV;
  if (x_out != x_in) GVF_TRY(copy_async(x_out, x_in, xbytes, stream));
  for (int i = 0; i < num_blocks; ++i) {
    const gvf_sparse_block& b = blocks[i];
    const gvf_window_partition& pt = parts[i & 1];
    GVF_TRY(gvf_ln_mod_f16(x_out, fp16_residual, A16, T, C, 1e-6f, nullptr, nullptr, nullptr, nullptr, 0, 0, stream));
    GVF_TRY(gvf_gemm_f16(A16, C, b.w_qkv, C, T, 3 * C, C, 0, b.b_qkv, QKV, 3 * C, nullptr, 0, 0, stream));
    if (pt.seq_of_pos)
      GVF_TRY(gvf_sparse_packed_attn_f16(QKV, AO, nullptr, pt.fwd_idx, nullptr, pt.cu_seqlens, pt.seq_of_pos, T, H, 64, scale, stream));
    else
      GVF_TRY(gvf_sparse_window_attn_f16(QKV, AO, pt.fwd_idx, pt.cu_seqlens, pt.num_windows, pt.max_seqlen, H, 64, scale, stream));
    GVF_TRY(gvf_gemm_f16(AO, C, b.w_out, C, T, C, C, epi, b.b_out, x_out, C, nullptr, 0, 0, stream));
    GVF_TRY(gvf_ln_mod_f16(x_out, fp16_residual, A16, T, C, 1e-6f, nullptr, nullptr, nullptr, nullptr, 0, 0, stream));
    GVF_TRY(gvf_gemm_f16(A16, C, b.w1, C, T, F, C, 1, b.b1, H1, F, nullptr, 0, 0, stream));
    GVF_TRY(gvf_gemm_f16(H1, F, b.w2, F, T, C, F, epi, b.b2, x_out, C, nullptr, 0, 0, stream));
  }
  return GVF_OK;
}

GVF_API int gvf_sparse_trunk_backward(const gvf_sparse_block* blocks, int num_blocks, int T, int C, int H, int F,
                                      int fp16_residual, const gvf_window_partition* parts, const void* arena,
                                      size_t arena_bytes, const void* d_out, void* scratch, size_t scratch_bytes,
                                      float* reduce_ws, size_t reduce_ws_bytes, void* d_in, void* stream) {
  if (!blocks || !parts || !arena || !d_out || !scratch || !reduce_ws || !d_in || bad_shape(T, C, H, F, num_blocks))
    return GVF_ERR_INVALID;
  const Layout L(T, C, H, F, num_blocks, fp16_residual);
  if (arena_bytes < L.total || scratch_bytes < gvf_sparse_trunk_scratch_bytes(T, C, H, F)) return GVF_ERR_WORKSPACE;
  // Two streams.  The critical path -- dgrad GEMM -> LayerNorm backward -> dgrad -> attention backward -> dgrad ->
  // LayerNorm backward -- stays on the caller's stream; the four weight-gradient GEMMs and bias sums of a block hang off
  // it (each needs one activation gradient and a saved activation, nothing needs them back) and go to a side stream, where
  // they fill the SMs the 96-tile dgrad GEMMs, the LayerNorm passes and the small attention grids leave idle.  Buffers the
  // side stream reads are double- (dx1, dH0, dQKV) or triple- (dx) buffered by block; block i waits for the side work of
  // block i + 2 before it reuses them.
  static cudaStream_t side = nullptr;
  static cudaEvent_t* ev = nullptr;
  static int ev_cap = 0;
  if (!side && cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking) != cudaSuccess) return GVF_ERR_CUDA;
  if (ev_cap < 5 * num_blocks + 1) {
    cudaEvent_t* nev = new cudaEvent_t[5 * num_blocks + 1];
    for (int k = 0; k < 5 * num_blocks + 1; ++k) {
      if (k < ev_cap) nev[k] = ev[k];
      else if (cudaEventCreateWithFlags(&nev[k], cudaEventDisableTiming) != cudaSuccess) return GVF_ERR_CUDA;
    }
    delete[] ev;
    ev = nev;
    ev_cap = 5 * num_blocks + 1;
  }
  cudaStream_t ms = (cudaStream_t)stream;
#define GVF_CU(expr) do { if ((expr) != cudaSuccess) return GVF_ERR_CUDA; } while (0)
  uint8_t* ar = (uint8_t*)arena;
  uint8_t* sc = (uint8_t*)scratch;
  uint8_t* dxb[3];
  uint8_t *dx1b[2], *dH0b[2], *dQKVb[2];
  for (int k = 0; k < 3; ++k) { dxb[k] = sc; sc += L.A; }
  for (int k = 0; k < 2; ++k) { dx1b[k] = sc; sc += L.A; }
  uint8_t* dA2 = sc;  sc += L.A;
  uint8_t* dAO = sc;  sc += L.A;
  uint8_t* dA = sc;   sc += L.A;
  for (int k = 0; k < 2; ++k) { dQKVb[k] = sc; sc += L.QKV; }
  for (int k = 0; k < 2; ++k) { dH0b[k] = sc; sc += L.Hf; }
  float* dsum = (float*)sc;
  const float scale = 0.125f;
  const void* dx = d_out;                                  // fp16 [T, C] gradient of the block's output
  GVF_CU(cudaEventRecord(ev[5 * num_blocks], ms));          // d_out is ready (whatever produced it ran on `stream`)
  GVF_CU(cudaStreamWaitEvent(side, ev[5 * num_blocks], 0));
  for (int i = num_blocks - 1; i >= 0; --i) {
    const gvf_sparse_block& b = blocks[i];
    if (!b.g_w_qkv || !b.g_b_qkv || !b.g_w_out || !b.g_b_out || !b.g_w1 || !b.g_b1 || !b.g_w2 || !b.g_b2) return GVF_ERR_INVALID;
    const gvf_window_partition& pt = parts[i & 1];
    const BlockSlots s = L.block(ar, num_blocks, i);
    const uint8_t* x0 = L.xslot(ar, i);
    cudaEvent_t* e = ev + 5 * i;                           // [0] dH0, [1] dx1, [2] dQKV, [3] dx of the next block, [4] side done
    uint8_t *dx1 = dx1b[i & 1], *dH0 = dH0b[i & 1], *dQKV = dQKVb[i & 1];
    if (i + 2 < num_blocks) GVF_CU(cudaStreamWaitEvent(ms, ev[5 * (i + 2) + 4], 0));
    // x2 = x1 + fc2(GELU(fc1(LN x1)))
    // dgrad GEMMs read the [out, in] weights as MN-major B operands (gvf_gemm_nn_f16): no transposed copies
    GVF_TRY(gvf_gemm_nn_f16(dx, C, b.w2, F, T, F, C, 8, dH0, F, s.H0, F, ms));                   // dgrad x GELU'
    GVF_CU(cudaEventRecord(e[0], ms));
    GVF_TRY(gvf_gemm_tn_f16(dx, C, s.Hg, F, C, F, T, b.g_w2, F, side));
    GVF_TRY(gvf_colsum(dx, 1, T, C, C, reduce_ws, reduce_ws_bytes, b.g_b2, 0, side));
    GVF_TRY(gvf_gemm_nn_f16(dH0, F, b.w1, C, T, C, F, 0, dA2, C, nullptr, 0, ms));
    GVF_CU(cudaStreamWaitEvent(side, e[0], 0));
    GVF_TRY(gvf_gemm_tn_f16(dH0, F, s.A2, C, F, C, T, b.g_w1, C, side));
    GVF_TRY(gvf_colsum(dH0, 1, T, F, F, reduce_ws, reduce_ws_bytes, b.g_b1, 0, side));
    GVF_TRY(gvf_ln_bwd_f16(s.x1, fp16_residual, dA2, dx, dx1, T, C, 1e-6f, ms));
    GVF_CU(cudaEventRecord(e[1], ms));
    // x1 = x0 + to_out(window attention(to_qkv(LN x0)))
    GVF_TRY(gvf_gemm_nn_f16(dx1, C, b.w_out, C, T, C, C, 0, dAO, C, nullptr, 0, ms));
    GVF_CU(cudaStreamWaitEvent(side, e[1], 0));
    GVF_TRY(gvf_gemm_tn_f16(dx1, C, s.AO, C, C, C, T, b.g_w_out, C, side));
    GVF_TRY(gvf_colsum(dx1, 1, T, C, C, reduce_ws, reduce_ws_bytes, b.g_b_out, 0, side));
    if (pt.seq_of_pos)
      GVF_TRY(gvf_sparse_packed_attn_bwd_f16(s.QKV, s.AO, dAO, (const float*)s.lse, dsum, dQKV, pt.fwd_idx, pt.cu_seqlens,
                                             pt.seq_of_pos, T, T, H, 64, scale, ms));
    else
      GVF_TRY(gvf_sparse_varlen_attn_bwd_f16(s.QKV, s.AO, dAO, (const float*)s.lse, dsum, dQKV, pt.fwd_idx, pt.cu_seqlens,
                                             pt.num_windows, pt.max_seqlen, T, H, 64, scale, ms));
    GVF_CU(cudaEventRecord(e[2], ms));
    GVF_TRY(gvf_gemm_nn_f16(dQKV, 3 * C, b.w_qkv, C, T, C, 3 * C, 0, dA, C, nullptr, 0, ms));
    GVF_CU(cudaStreamWaitEvent(side, e[2], 0));
    GVF_TRY(gvf_gemm_tn_f16(dQKV, 3 * C, s.A, C, 3 * C, C, T, b.g_w_qkv, C, side));
    GVF_TRY(gvf_colsum(dQKV, 1, T, 3 * C, 3 * C, reduce_ws, reduce_ws_bytes, b.g_b_qkv, 0, side));
    GVF_CU(cudaEventRecord(e[4], side));
    void* dnext = (i == 0) ? d_in : (void*)dxb[i % 3];
    GVF_TRY(gvf_ln_bwd_f16(x0, fp16_residual, dA, dx1, dnext, T, C, 1e-6f, ms));
    GVF_CU(cudaEventRecord(e[3], ms));
    GVF_CU(cudaStreamWaitEvent(side, e[3], 0));            // the next block's side work reads dnext
    dx = dnext;
  }
  // join: every gradient is complete when the caller's stream passes this point
  GVF_CU(cudaStreamWaitEvent(ms, ev[4], 0));
  if (num_blocks > 1) GVF_CU(cudaStreamWaitEvent(ms, ev[5 + 4], 0));
#undef GVF_CU
  return GVF_OK;
}

}  // extern "C"
