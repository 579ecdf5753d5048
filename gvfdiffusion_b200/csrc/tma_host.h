// tma_host.h -- host-side CUtensorMap construction without linking libcuda (the driver
// entry point is fetched through the runtime so the library still loads on a CPU-only box).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace gvf {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

inline CUtensorMapSwizzle swizzle_for_bytes(int inner_bytes) {
  return inner_bytes >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                            : inner_bytes >= 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                : CU_TENSOR_MAP_SWIZZLE_32B;
}

// fp16 tensor, up to 4 dims (innermost first).  strides in ELEMENTS for dims 1..rank-1.
// returns false on failure.
inline bool make_tmap_f16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_elems, const uint32_t* box,
                          CUtensorMapSwizzle sw) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  cuuint64_t d[5], s[5];
  cuuint32_t b[5], e[5];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
    e[i] = 1;
    if (i > 0) s[i - 1] = strides_elems[i] * 2;
  }
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), d, s, b,
            e, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// row-major 2-D tensor of 2-byte (fp16) or 4-byte (fp32) elements, SWIZZLE_128B boxes (epilogue TMA stores)
inline bool make_tmap_2d(CUtensorMap* m, const void* base, int elem_bytes, uint64_t cols, uint64_t rows,
                         uint64_t ld_elems, uint32_t box_cols, uint32_t box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  cuuint64_t d[2] = {cols, rows}, s[1] = {ld_elems * (uint64_t)elem_bytes};
  cuuint32_t b[2] = {box_cols, box_rows}, e[2] = {1, 1};
  return fn(m, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
            const_cast<void*>(base), d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace gvf
