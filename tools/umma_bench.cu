// Micro-benchmark: tensor-pipe time of the attention kernel's MMA sequences (one CTA per SM, M = 128, cta_group::1).
// How long do 13 dependent P.V steps (N = 80, K = 16, A from TMEM) take, against the same steps on two
// accumulators, and against the four Q.K^T steps (N = 208, K = 16, both operands from shared memory)?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../include -o umma_bench umma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../vit.triton_b200/csrc/common.cuh"
using namespace vt;

// mode: 0 = SS, one accumulator; 1 = SS, alternating between two accumulators; 2 = TS one acc; 3 = TS two acc
__global__ void __launch_bounds__(128, 1) bench(int n, int steps, int mode, int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar_mem;
  __shared__ uint32_t slot;
  const uint32_t bar = smem_u32(&bar_mem);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc<512>(smem_u32(&slot)); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1) {
    const uint32_t idesc_k = make_idesc_bf16(128, n, 0, 0);       // K-major B (scores)
    const uint32_t idesc_mn = make_idesc_bf16(128, n, 0, 1);      // MN-major B (P.V)
    const uint64_t adesc = make_desc_kmajor_sw128(base);
    const uint32_t b_smem = base + 16384;
    uint32_t ph = 0;
    long long best = 1LL << 60;
    for (int r = 0; r < reps; ++r) {
      const long long t0 = clock64();
      if (elect_one_sync()) {
        for (int k = 0; k < steps; ++k) {
          const uint32_t d = tmem + (mode < 2 ? 0 : 256) + ((mode & 1) ? (k & 1) * (mode < 2 ? 256 : 96) : 0);
          if (mode < 2) umma_ss(d, adesc + 2 * (k & 3), make_desc_kmajor_sw128(b_smem) + 2 * (k & 3), idesc_k, k > 1 ? 1u : 0u);
          else umma_ts(d, tmem + 8 * k, make_desc_mnmajor_sw128(b_smem + 2048u * k, 8192), idesc_mn, k > 1 ? 1u : 0u);
        }
        umma_commit(bar);
      }
      __syncwarp();
      while (!mbar_test_wait(bar, ph)) {}
      ph ^= 1u;
      const long long t1 = clock64();
      if (t1 - t0 < best) best = t1 - t0;
    }
    if ((threadIdx.x & 31) == 0) out[blockIdx.x] = best;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

int main() {
  long long* out;
  cudaMalloc(&out, 148 * sizeof(long long));
  const int smem = 1024 + 16384 + 98304;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  struct { int n, steps, mode; const char* what; } cases[] = {
      {208, 4, 0, "Q.K^T  4 x N=208 SS, one accumulator"},
      {80, 13, 0, "13 x N=80 SS, one accumulator"},
      {80, 13, 1, "13 x N=80 SS, two accumulators"},
      {80, 13, 2, "P.V   13 x N=80 TS, one accumulator"},
      {80, 13, 3, "P.V   13 x N=80 TS, two accumulators"},
      {64, 13, 2, "13 x N=64 TS, one accumulator"},
      {80, 1, 2, "1 x N=80 TS"},
      {80, 2, 2, "2 x N=80 TS, one accumulator"},
      {80, 26, 2, "26 x N=80 TS, one accumulator"},
      {208, 1, 0, "1 x N=208 SS"},
      {256, 8, 0, "8 x N=256 SS, one accumulator"},
      {256, 8, 1, "8 x N=256 SS, two accumulators"},
      {160, 7, 2, "7 x N=160 TS, one accumulator"},
  };
  for (auto& c : cases) {
    bench<<<148, 128, smem>>>(c.n, c.steps, c.mode, 50, out);
    cudaError_t rc = cudaDeviceSynchronize();
    if (rc != cudaSuccess) { printf("%s: %s\n", c.what, cudaGetErrorString(rc)); return 1; }
    long long h[148];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    long long mn = h[0], mx = h[0];
    for (int i = 1; i < 148; ++i) { mn = h[i] < mn ? h[i] : mn; mx = h[i] > mx ? h[i] : mx; }
    const double ideal = static_cast<double>(c.steps) * 128.0 * c.n * 16.0 / 4096.0;
    printf("%-72s best %5lld .. %5lld cycles (issue -> commit seen; ideal pipe time %.0f)\n", c.what, mn, mx, ideal);
  }
  return 0;
}
