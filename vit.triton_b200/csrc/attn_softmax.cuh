// Softmax building blocks shared by the persistent attention kernels (attn5_sm100.cu; the superseded generations under tools/attn_generations/):
// row max and exp2 / bf16-P passes of one thread over a range of its row's score columns in TMEM.
#pragma once

#include "common.cuh"

namespace vt {

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Row max over this thread's score columns [c0, c1) of which [c0, min(c1, nvalid_end)) are valid.
__device__ __forceinline__ float row_max_part(uint32_t t_lane, int c0, int c1, int nvalid) {
  float mx0 = -INFINITY, mx1 = -INFINITY;
  int c = c0;
  const int full_end = c0 + (((nvalid < c1 ? nvalid : c1) - c0) & ~31);   // end of fully valid 32-col chunks
  for (; c + 64 <= full_end; c += 64) {
    uint32_t ra[32], rb[32];
    tmem_ld_32x32(t_lane + c, ra);
    tmem_ld_32x32(t_lane + c + 32, rb);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      mx0 = fmax3(mx0, __uint_as_float(ra[i]), __uint_as_float(ra[i + 1]));
      mx1 = fmax3(mx1, __uint_as_float(rb[i]), __uint_as_float(rb[i + 1]));
    }
  }
  for (; c < full_end; c += 32) {
    uint32_t r[32];
    tmem_ld_32x32(t_lane + c, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      mx0 = fmax3(mx0, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
      mx1 = fmax3(mx1, __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
    }
  }
  for (; c < c1; c += 16) {
    uint32_t r[16];
    tmem_ld_32x16(t_lane + c, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (c + i < nvalid) mx0 = fmaxf(mx0, __uint_as_float(r[i]));
  }
  return fmaxf(mx0, mx1);
}

// p = exp2(s * scale - m) over this thread's columns [c0, c1); P (bf16x2) written to TMEM columns
// p_col + (c - c0)/2; returns the partial row sum.  Next chunk's load is in flight during the math.
__device__ __forceinline__ float exp_part(uint32_t t_lane, int c0, int c1, int nvalid, uint32_t p_col,
                                          float scale_log2, float m) {
  float ps0 = 0.f, ps1 = 0.f, ps2 = 0.f, ps3 = 0.f;
  const int full_end = c0 + (((nvalid < c1 ? nvalid : c1) - c0) & ~31);
  int c = c0;
  for (; c < full_end; c += 32) {
    uint32_t r[32];
    tmem_ld_32x32(t_lane + c, r);
    tmem_ld_wait();
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float p0 = ex2_approx(fmaf(__uint_as_float(r[i + 0]), scale_log2, -m));
      const float p1 = ex2_approx(fmaf(__uint_as_float(r[i + 1]), scale_log2, -m));
      const float p2 = ex2_approx(fmaf(__uint_as_float(r[i + 2]), scale_log2, -m));
      const float p3 = ex2_approx(fmaf(__uint_as_float(r[i + 3]), scale_log2, -m));
      ps0 += p0; ps1 += p1; ps2 += p2; ps3 += p3;
      pk[(i >> 1) + 0] = pack_bf16x2(p0, p1);
      pk[(i >> 1) + 1] = pack_bf16x2(p2, p3);
    }
    tmem_st_32x16(p_col + ((c - c0) >> 1), pk);
  }
  for (; c < c1; c += 16) {
    uint32_t r[16];
    tmem_ld_32x16(t_lane + c, r);
    tmem_ld_wait();
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      float p0 = ex2_approx(fmaf(__uint_as_float(r[i]), scale_log2, -m));
      float p1 = ex2_approx(fmaf(__uint_as_float(r[i + 1]), scale_log2, -m));
      if (c + i >= nvalid) p0 = 0.f;
      if (c + i + 1 >= nvalid) p1 = 0.f;
      ps0 += p0;
      ps1 += p1;
      pk[i >> 1] = pack_bf16x2(p0, p1);
    }
    tmem_st_32x8(p_col + ((c - c0) >> 1), pk);
  }
  tmem_st_wait();
  return (ps0 + ps1) + (ps2 + ps3);
}

}  // namespace vt
