#include "tensormap.h"

#include <mutex>

namespace vt {

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  return fn;
}

static CUtensorMapSwizzle to_cu(TmapSwizzle s) {
  switch (s) {
    case TMAP_SW_32: return CU_TENSOR_MAP_SWIZZLE_32B;
    case TMAP_SW_64: return CU_TENSOR_MAP_SWIZZLE_64B;
    case TMAP_SW_128: return CU_TENSOR_MAP_SWIZZLE_128B;
    default: return CU_TENSOR_MAP_SWIZZLE_NONE;
  }
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows,
                      uint64_t row_stride_elems, uint32_t box_cols, uint32_t box_rows,
                      TmapSwizzle sw) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return -5;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_stride_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, to_cu(sw),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -5;
}

int make_tmap_u8_2d(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows,
                    uint64_t row_stride_bytes, uint32_t box_cols, uint32_t box_rows, TmapSwizzle sw) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return -5;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, to_cu(sw),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -5;
}

int make_tmap_bf16_3d(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows,
                      uint64_t batch, uint64_t row_stride_elems, uint64_t batch_stride_elems,
                      uint32_t box_cols, uint32_t box_rows, TmapSwizzle sw) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return -5;
  cuuint64_t dims[3] = {cols, rows, batch};
  cuuint64_t strides[2] = {row_stride_elems * 2, batch_stride_elems * 2};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, to_cu(sw),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -5;
}

}  // namespace vt
