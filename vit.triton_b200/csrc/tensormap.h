// Host-side CUtensorMap construction without linking libcuda: the driver entry point is fetched
// through the runtime (cudaGetDriverEntryPoint) once per process.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vt {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_tiled();

enum TmapSwizzle { TMAP_SW_NONE = 0, TMAP_SW_32 = 1, TMAP_SW_64 = 2, TMAP_SW_128 = 3 };

// 2-D bf16 tensor: `cols` contiguous elements per row, `rows` rows, `row_stride_elems` between
// rows.  Box = box_cols x box_rows.  Returns 0 on success.
int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows,
                      uint64_t row_stride_elems, uint32_t box_cols, uint32_t box_rows,
                      TmapSwizzle sw);

// 2-D tensor of single bytes (e4m3 operands / outputs): `cols` contiguous bytes per row.
int make_tmap_u8_2d(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows,
                    uint64_t row_stride_bytes, uint32_t box_cols, uint32_t box_rows, TmapSwizzle sw);

// 3-D bf16 tensor (cols, rows, batch) with element strides for rows and batches.
int make_tmap_bf16_3d(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows,
                      uint64_t batch, uint64_t row_stride_elems, uint64_t batch_stride_elems,
                      uint32_t box_cols, uint32_t box_rows, TmapSwizzle sw);

}  // namespace vt
