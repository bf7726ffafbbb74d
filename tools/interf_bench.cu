// Micro-benchmark: how much do the OTHER slot's phases slow down the attention exp pass?
// Warps 0-7 run the exp pass (TMEM S -> exp2 -> bf16 P -> TMEM) over 96 columns; warps 8-15 run an
// interference loop until warps 0-7 are done.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o interf_bench interf_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../vit.triton_b200/csrc/common.cuh"
using namespace vt;

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

template <int MODE>
__global__ void __launch_bounds__(512) bench(int iters, long long* out_cycles, float* sink, float scale) {
  __shared__ uint32_t slot;
  __shared__ volatile int done_flag;
  __shared__ uint64_t never_bar;
  __shared__ uint4 scratch[512];
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { done_flag = 0; mbar_init(smem_u32(&never_bar), 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc<512>(smem_u32(&slot)); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t t_lane = slot + ((uint32_t)((warp & 3) * 32) << 16) + ((warp >> 2) & 3) * 128;
  float ps0 = 0, ps1 = 0, ps2 = 0, ps3 = 0;
  const float m_new = 3.0f;
  __syncthreads();
  if (warp < 8) {
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int c = 0; c < 96; c += 32) {
        uint32_t r[32]; tmem_ld_32x32(t_lane + c, r); tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          float p0 = ex2_approx(fmaf(__uint_as_float(r[i]), scale, -m_new));
          float p1 = ex2_approx(fmaf(__uint_as_float(r[i+1]), scale, -m_new));
          float p2 = ex2_approx(fmaf(__uint_as_float(r[i+2]), scale, -m_new));
          float p3 = ex2_approx(fmaf(__uint_as_float(r[i+3]), scale, -m_new));
          ps0 += p0; ps1 += p1; ps2 += p2; ps3 += p3;
          pk[(i>>1)] = pack_bf16x2(p0, p1); pk[(i>>1)+1] = pack_bf16x2(p2, p3);
        }
        tmem_st_32x16(t_lane + (c >> 1), pk);
      }
      tmem_st_wait();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out_cycles[blockIdx.x] = t1 - t0; }
    __syncwarp();
    if ((threadIdx.x & 31) == 0) atomicAdd((int*)&done_flag, 1);
  } else {
    if (MODE == 1) {          // pass-1 like: TMEM loads + 3-input max
      while (done_flag < 8) {
        uint32_t ra[32];
        tmem_ld_32x32(t_lane, ra); tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 2) ps0 = fmax3(ps0, __uint_as_float(ra[i]), __uint_as_float(ra[i + 1]));
      }
    } else if (MODE == 2) {   // epilogue like: FMUL + pack + st.shared
      float a = threadIdx.x;
      while (done_flag < 8) {
        uint4 o;
        o.x = pack_bf16x2(a * 1.1f, a * 1.2f); o.y = pack_bf16x2(a * 1.3f, a * 1.4f);
        o.z = pack_bf16x2(a * 1.5f, a * 1.6f); o.w = pack_bf16x2(a * 1.7f, a * 1.8f);
        scratch[threadIdx.x] = o;
        a += 1.f;
      }
      ps0 = a;
    } else if (MODE == 3) {   // sleeping on an mbarrier that never completes
      while (done_flag < 8) { mbar_try_wait_hint(smem_u32(&never_bar), 0, 2000u); }
    } else if (MODE == 4) {   // un-hinted polling
      while (done_flag < 8) { mbar_try_wait(smem_u32(&never_bar), 0); }
    } else if (MODE == 5) {   // second exp pass at the same time (no turn taking)
      while (done_flag < 8) {
        uint32_t r[32]; tmem_ld_32x32(t_lane, r); tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) ps1 += ex2_approx(fmaf(__uint_as_float(r[i]), scale, -m_new));
      }
    } else if (MODE == 6) {   // integer ALU work only
      int a = threadIdx.x;
      while (done_flag < 8) {
#pragma unroll
        for (int i = 0; i < 32; ++i) a = a * 3 + (a >> 3);
      }
      ps0 = a;
    }
  }
  __syncthreads();
  if (ps0 + ps1 + ps2 + ps3 == 12345.f) sink[0] = ps0;
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(slot);
}

int main() {
  long long* d; float* s; cudaMalloc(&d, 148 * 8); cudaMalloc(&s, 4);
  const int iters = 300;
#define RUN(M, name) { bench<M><<<148, 512>>>(iters, d, s, 0.18f); cudaError_t e = cudaDeviceSynchronize(); long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); \
      printf("interference %-28s: %.0f cycles per 96-col pass (%.2f cyc/elem/warp, MUFU-bound = 16) %s\n", name, (double)h / iters, (double)h / iters / 96, cudaGetErrorString(e)); }
  RUN(0, "none") RUN(1, "LDTM + max3 (pass 1)") RUN(2, "FMUL + pack + STS (epilogue)") RUN(3, "mbarrier sleep (hinted)")
  RUN(4, "mbarrier poll (no hint)") RUN(5, "second exp pass") RUN(6, "integer ALU")
  return 0;
}
