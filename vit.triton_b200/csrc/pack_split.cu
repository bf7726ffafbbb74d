// Operand packing for vt_bgemm (bgemm_sm100.cu): any strided fp32 / bf16 matrix batch -> dense, 16-byte
// aligned bf16 K-major rows the tensor maps can address, optionally SPLIT into bf16 pieces:
//
//   x = hi + lo + O(2^-17 |x|),  hi = bf16(x),  lo = bf16(x - hi)
//
// laid side by side along the row:  pattern 0 (A side)  [hi | hi | lo],  pattern 1 (B side)  [hi | lo | hi],
// so that ONE bf16 GEMM over K' = 3 * cpad accumulates  a_hi b_hi + a_hi b_lo + a_lo b_hi  in fp32 — the
// "3 x bf16" form of an fp32 product (4.7e-5 on the final hidden states, where one TF32
// pass — what the reference's tl.dot does, vit/kernels/matmul.py:92 — gives 3.1e-3).
// pieces == 6 is the three-way split x = x1 + x2 + x3 (24 mantissa bits, i.e. all of fp32) with the six
// products of weight >= 2^-16:  A side [a1|a1|a2|a1|a2|a3],  B side [b1|b2|b1|b3|b2|b1]  — exact to 2^-24 in
// exact arithmetic, but measured NO better than three pieces on B200 (5.3e-5 vs 3.2e-5 at K = 768): the
// tensor core's fp32 accumulation truncates, and twice the K steps cost more than lo*lo recovers.  The fp32
// model therefore uses three pieces (kernels/bgemm.py:split_pieces).
// With pieces == 1 it is a plain (transposing, zero-padding) conversion: odd row lengths (197 keys) and
// [K, N] operands of matmul3 (vit/kernels/matmul3.py:111-156) become K-major rows of a multiple of 8.
#include "common.cuh"

namespace vt {

namespace {

struct PackParams {
  const void* src;
  __nv_bfloat16* dst;
  int R, C, cpad, pieces, pattern;
  int batch_inner;
  long long total;                 // batch * R * cpad
  long long sSo, sSi, sSr, sSc;    // source element strides
  long long sDo, sDi, sDr;         // destination element strides (column stride 1)
};

template <typename T>
__global__ void __launch_bounds__(256)
pack_split_kernel(const PackParams p) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < p.total; idx += stride) {
    const int c = static_cast<int>(idx % p.cpad);
    const long long t = idx / p.cpad;
    const int r = static_cast<int>(t % p.R);
    const int z = static_cast<int>(t / p.R);
    const int zo = z / p.batch_inner, zi = z - zo * p.batch_inner;
    float x = 0.f;
    if (c < p.C) {
      const T* s = static_cast<const T*>(p.src) + zo * p.sSo + zi * p.sSi + r * p.sSr + c * p.sSc;
      if constexpr (sizeof(T) == 4) x = *s; else x = __bfloat162float(*s);
    }
    __nv_bfloat16* d = p.dst + zo * p.sDo + zi * p.sDi + r * p.sDr + c;
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    if (p.pieces == 1) {
      d[0] = hi;
    } else {
      const float r1 = x - __bfloat162float(hi);
      const __nv_bfloat16 lo = __float2bfloat16_rn(r1);
      d[0] = hi;
      d[p.cpad] = p.pattern == 0 ? hi : lo;
      d[2 * p.cpad] = p.pattern == 0 ? lo : hi;
      if (p.pieces == 6) {
        const __nv_bfloat16 lo2 = __float2bfloat16_rn(r1 - __bfloat162float(lo));
        d[3 * p.cpad] = p.pattern == 0 ? hi : lo2;
        d[4 * p.cpad] = lo;
        d[5 * p.cpad] = p.pattern == 0 ? lo2 : hi;
      }
    }
  }
}

}  // namespace

int pack_bf16(const void* src, int src_dtype, void* dst, int R, int C, int batch_outer, int batch_inner,
              const long long* sS, const long long* sD, int cpad, int pieces, int pattern, cudaStream_t stream) {
  if (!src || !dst || !sS || !sD || R <= 0 || C <= 0 || batch_outer <= 0 || batch_inner <= 0 || cpad < C)
    return VT_ERR_ARG;
  if (pieces != 1 && pieces != 3 && pieces != 6) return VT_ERR_ARG;
  if (pattern != 0 && pattern != 1) return VT_ERR_ARG;
  PackParams p;
  p.src = src;
  p.dst = static_cast<__nv_bfloat16*>(dst);
  p.R = R; p.C = C; p.cpad = cpad; p.pieces = pieces; p.pattern = pattern;
  p.batch_inner = batch_inner;
  p.total = static_cast<long long>(batch_outer) * batch_inner * R * cpad;
  p.sSo = sS[0]; p.sSi = sS[1]; p.sSr = sS[2]; p.sSc = sS[3];
  p.sDo = sD[0]; p.sDi = sD[1]; p.sDr = sD[2];
  long long blocks = (p.total + 255) / 256;
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  if (src_dtype == VT_F32)
    pack_split_kernel<float><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(p);
  else if (src_dtype == VT_BF16)
    pack_split_kernel<__nv_bfloat16><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(p);
  else
    return VT_ERR_DTYPE;
  return static_cast<int>(cudaGetLastError());
}

}  // namespace vt
